// lvt_b200 :: bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA into 128B-swizzled shared memory, mbarrier pipelines).
//
// Persistent, warp-specialised kernel, one CTA per SM, 128 x BN output tiles (BN = 128 / 256):
//   warp 0    : TMA producer           (one elected lane, STAGES-deep smem ring across tiles)
//   warp 1    : TMEM owner + MMA issuer (one elected lane issues tcgen05.mma)
//   warps 2-5 : epilogue group 0  -> TMEM accumulator buffer 0 (even tiles of this CTA)
//   warps 6-9 : epilogue group 1  -> TMEM accumulator buffer 1 (odd tiles)
// The accumulator is double-buffered in TMEM (2 x BN columns), so the MMAs of tile i+1 overlap the
// epilogue of tile i.  Epilogue: TMEM -> registers (lane == row) -> per-warp smem transpose ->
// coalesced 16 B global accesses (bias / residual / mask reads and the stores).
// Replaces the cuBLAS calls under torch.bmm / nn.Linear / 1x1x1 Conv3d on the DSFVT path
// (reference: vidgen/modeling/autoregressive/vt_attention.py:63-80,120-128,138 and
// videotransformer.py:57,99,148-156) and their autograd backward.
#include <unordered_map>
#include <mutex>
#include <string>
#include <string.h>
#include <stdlib.h>

#include "../../include/lvt_b200.h"
#include "common.cuh"

extern void lvt_count_launch(int n);
extern int lvt_sm_limit();

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int NUM_THREADS = 320;           // 10 warps
constexpr int NUM_EPI_WARPS = 8;
// per-warp transpose buffer: 32 rows x 4 units of 16 B, unit k of row r stored at r*4 + (k ^ ((r>>1)&3))
// -> both the row-per-lane writes and the 4-lanes-per-row reads are bank-conflict free
constexpr int STG_BYTES_PER_WARP = 4096;  // generic path uses 2 KiB; TMA-store slab = 32 rows x 128 B
LVT_DEVICE_INLINE int stg_unit(int row, int k) { return row * 4 + (k ^ ((row >> 1) & 3)); }

// epilogue kinds (compile-time)
constexpr int EK_LINEAR = 0;
constexpr int EK_SOFTMAX_1x16x16 = 1;
constexpr int EK_SOFTMAX_4x8x8 = 2;
constexpr int EK_DS = 3;

struct GemmParams {
  int M, N, K, batch, splits;
  int tiles_m, tiles_n, total_tiles, kb_per;
  int a_cin, a_zdiv, b_cin, b_zdiv;
  int flags;
  float alpha;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  const float* res;
  const __nv_bfloat16* aux;
  const float* bias;
  int bias_mod;
  int o_cin, o_zdiv, o_shift;  // o_shift: log2(o_cin) for blocked outputs, -1 for plain rows
  long long o_ld, o_s_blk, o_s_zlo, o_s_zhi;
  float* lse;
  const float* delta;
  const float* bank_t;
  const float* bank_h;
  const float* bank_w;
  int heads;
  float* rowdot;  // LVT_GEMM_ROWDOT
  int rd_block, rd_L;
  // implicit-GEMM convolution (operand = NHWC activations read through per-tap shifted TMA boxes)
  int a_conv, b_conv, cv_C, cv_W, cv_H;
  // per-tap offsets packed 4 bits each (value + 8): no dynamically indexed parameter arrays -> no stack copy
  unsigned long long cv_dh_pk, cv_dw_pk, cv_ph_pk;
  long long* prof;  // fused attention kernels: optional clock64 timeline of CTA 0 (tools/attn_fwd_prof.py)
};
LVT_DEVICE_INLINE int cv_tap(unsigned long long pk, int tap) { return (int)((pk >> (4 * tap)) & 15ull) - 8; }

// store-path bits (template parameter ST); 0 = staged generic epilogue
constexpr int ST_TMA = 1;     // epilogue writes 128 B-row slabs to smem and stores them with TMA
constexpr int ST_F32 = 2;     // slab holds 32 fp32 columns (else 64 bf16 columns)
constexpr int ST_REDUCE = 4;  // cp.reduce.async.bulk (+=) instead of a plain store (split-K / grad accumulate)
constexpr int ST_CLOAD = 8;   // a second tensor with the output's addressing is TMA-loaded per slab:
                              // fp32 residual (added), bf16 ReLU-mask source, or bf16 P of the DS epilogue

template <int BN, int ST, int EK = EK_LINEAR, int CG = 1>
struct SmemLayout {
  static constexpr bool SOFTMAX = EK == EK_SOFTMAX_1x16x16 || EK == EK_SOFTMAX_4x8x8;
  static constexpr int B_TILE_BYTES = (BN / CG) * BK * 2;  // a CTA pair (CG == 2) splits the B tile
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  // (the attention-probability kernel has K = da = 128: two k-blocks per tile; three stages keep 1.5 tiles in flight)
  // CTA pairs double-buffer the TMA-store slab of every epilogue warp (a warp fills slab i+1 while the store of
  // slab i is still reading shared memory): measured, the single-slab epilogue -- not the main loop, which runs
  // at ~94 % of the cuBLAS rate with the epilogue switched off -- bounded the large GEMMs.
  static constexpr int NSLAB = (CG == 2 && (ST & ST_TMA)) ? 2 : 1;
  static constexpr int STAGES = SOFTMAX ? 3
                                : CG == 2 ? 6 - ((ST & ST_TMA) ? 1 : 0) - ((ST & ST_CLOAD) ? 1 : 0)
                                          : (BN == 256 ? 4 : 6) - ((ST & ST_CLOAD) ? 1 : 0);
  static constexpr int STG_PER_WARP = NSLAB * STG_BYTES_PER_WARP + ((ST & ST_CLOAD) ? STG_BYTES_PER_WARP : 0);
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int XCHG_OFFSET = STG_OFFSET + NUM_EPI_WARPS * STG_PER_WARP;  // [2][2][128] fp32 row max / row sum
  static constexpr int BAR_OFFSET = XCHG_OFFSET + (SOFTMAX ? 2048 : 0);
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024 /*alignment slack*/;
};

LVT_DEVICE_INLINE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

LVT_DEVICE_INLINE void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

struct TileCoord {
  int m0, n0, z, kb_begin, num_kb;
};

// (CTA pairs: a tile has cg * BM rows, this CTA owns the BM rows starting at rank * BM)
LVT_DEVICE_INLINE TileCoord decode_tile(const GemmParams& p, int tile, int bn, int cg = 1, int rank = 0) {
  TileCoord t;
  const int tn = tile % p.tiles_n;
  int r = tile / p.tiles_n;
  const int tm = r % p.tiles_m;
  r /= p.tiles_m;
  const int split = r % p.splits;
  t.z = r / p.splits;
  t.m0 = (tm * cg + rank) * BM;
  t.n0 = tn * bn;
  const int kb_total = (p.K + BK - 1) / BK;
  t.kb_begin = split * p.kb_per;
  t.num_kb = max(0, min(kb_total, t.kb_begin + p.kb_per) - t.kb_begin);
  return t;
}

// ST: 0 = staged generic epilogue (residual / mask / atomic / dual output / ragged N),
//     1 = TMA-store epilogue, bf16 output,  2 = TMA-store epilogue, fp32 output
// CG == 2: the kernel runs as CTA pairs (cluster of 2, tcgen05 cta_group::2) on 256 x BN tiles; see common.cuh.
template <int BN, bool A_MN, bool B_MN, int EK, int ST, int CG = 1>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_c,
                 const GemmParams p) {
  using L = SmemLayout<BN, ST, EK, CG>;
  constexpr int STAGES = L::STAGES;
  constexpr bool SOFTMAX = L::SOFTMAX;
  constexpr bool PAIR = CG == 2;
  static_assert(!PAIR || (EK == EK_LINEAR && BN == 256), "CTA pairs: linear epilogue, 256-wide tiles");
  const int rank = PAIR ? (int)cluster_ctarank() : 0;        // leader = 0
  const int cta0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // first tile of this CTA (pair)
  const int cta_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* c_bar = tmem_empty_bar + 2;           // [NUM_EPI_WARPS] C-slab loads (ST_CLOAD)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(c_bar + NUM_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar[0], 1);
    mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], SOFTMAX ? 8 : 4 * CG);  // one arrive per epilogue warp working on the buffer
    mbar_init(&tmem_empty_bar[1], SOFTMAX ? 8 : 4 * CG);  // (pair: the leader's barrier also counts the peer's warps)
#pragma unroll
    for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(&c_bar[w], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_2sm(tmem_ptr_smem, 2 * BN);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr_smem, 2 * BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync();  // the peer's barriers are initialised before any remote arrive / TMA credit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its results are visible from here on

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t kiter = 0;
      // pair: both CTAs load (own A rows, own half of the B tile); every byte is credited to the leader's barrier
      auto load5 = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
        if constexpr (PAIR) tma_load_5d_2sm(dst, m, bar, c0, c1, c2, c3, c4);
        else tma_load_5d(dst, m, bar, c0, c1, c2, c3, c4);
      };
      constexpr int BNL = BN / CG;  // B rows / columns staged by this CTA
      for (int tile = cta0; tile < p.total_tiles; tile += cta_stride) {
        TileCoord t = decode_tile(p, tile, BN, CG, rank);
        t.n0 += rank * BNL;
        const int a_zlo = t.z % p.a_zdiv, a_zhi = t.z / p.a_zdiv;
        const int b_zlo = t.z % p.b_zdiv, b_zhi = t.z / p.b_zdiv;
        if (!(p.a_conv | p.b_conv)) {  // plain operands: keep this loop free of the conv address math
          for (int it = 0; it < t.num_kb; ++it, ++kiter) {
            const int s = kiter % STAGES;
            const uint32_t ph = (kiter / STAGES) & 1;
            const int k0 = (t.kb_begin + it) * BK;
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], CG * L::STAGE_BYTES);
            uint8_t* a_dst = smem + s * L::STAGE_BYTES;
            uint8_t* b_dst = a_dst + A_TILE_BYTES;
            if (!A_MN) {
              load5(a_dst, &tm_a, &full_bar[s], k0 % p.a_cin, t.m0, k0 / p.a_cin, a_zlo, a_zhi);
            } else {
  #pragma unroll
              for (int j = 0; j < BM / 64; ++j) {
                const int c = t.m0 + 64 * j;
                load5(a_dst + j * (64 * BK * 2), &tm_a, &full_bar[s], c % p.a_cin, k0, c / p.a_cin,
                      a_zlo, a_zhi);
              }
            }
            if (!B_MN) {
              load5(b_dst, &tm_b, &full_bar[s], k0 % p.b_cin, t.n0, k0 / p.b_cin, b_zlo, b_zhi);
            } else {
  #pragma unroll
              for (int j = 0; j < BNL / 64; ++j) {
                const int c = t.n0 + 64 * j;
                load5(b_dst + j * (64 * BK * 2), &tm_b, &full_bar[s], c % p.b_cin, k0, c / p.b_cin,
                      b_zlo, b_zhi);
              }
            }
          }
          continue;
        }
        for (int it = 0; it < t.num_kb; ++it, ++kiter) {
          const int s = kiter % STAGES;
          const uint32_t ph = (kiter / STAGES) & 1;
          const int k0 = (t.kb_begin + it) * BK;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], CG * L::STAGE_BYTES);
          uint8_t* a_dst = smem + s * L::STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_TILE_BYTES;
          if (!A_MN) {
            if (p.a_conv) {
              // rows = 128 consecutive output pixels (n, h, w) of one image; k-block = 64 channels of one tap
              const int tap = k0 / p.cv_C, c0 = k0 - tap * p.cv_C;
              const int hw = p.cv_H * p.cv_W;
              const int img = t.m0 / hw, h0 = (t.m0 - img * hw) / p.cv_W;
              load5(a_dst, &tm_a, &full_bar[s], c0, cv_tap(p.cv_dw_pk, tap), h0 + cv_tap(p.cv_dh_pk, tap), img, cv_tap(p.cv_ph_pk, tap));
            } else {
              load5(a_dst, &tm_a, &full_bar[s], k0 % p.a_cin, t.m0, k0 / p.a_cin, a_zlo, a_zhi);
            }
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) {
              const int c = t.m0 + 64 * j;
              load5(a_dst + j * (64 * BK * 2), &tm_a, &full_bar[s], c % p.a_cin, k0, c / p.a_cin,
                    a_zlo, a_zhi);
            }
          }
          if (!B_MN) {
            load5(b_dst, &tm_b, &full_bar[s], k0 % p.b_cin, t.n0, k0 / p.b_cin, b_zlo, b_zhi);
          } else {
#pragma unroll
            for (int j = 0; j < BNL / 64; ++j) {
              const int c = t.n0 + 64 * j;
              if (p.b_conv) {
                // weight gradient: columns = (tap, channel), k-block = 64 consecutive pixels of one image
                const int tap = c / p.cv_C, c0 = c - tap * p.cv_C;
                const int hw = p.cv_H * p.cv_W;
                const int img = k0 / hw, h0 = (k0 - img * hw) / p.cv_W;
                load5(b_dst + j * (64 * BK * 2), &tm_b, &full_bar[s], c0, cv_tap(p.cv_dw_pk, tap),
                      h0 + cv_tap(p.cv_dh_pk, tap), img, cv_tap(p.cv_ph_pk, tap));
              } else {
                load5(b_dst + j * (64 * BK * 2), &tm_b, &full_bar[s], c % p.b_cin, k0, c / p.b_cin,
                      b_zlo, b_zhi);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (rank == 0 && elect_one()) {  // (pair: the leader issues for both CTAs)
      constexpr uint32_t idesc = umma_idesc(BM * CG, BN, /*bf16*/ 1, A_MN, B_MN);
      uint32_t kiter = 0, titer = 0;
      for (int tile = cta0; tile < p.total_tiles; tile += cta_stride, ++titer) {
        const TileCoord t = decode_tile(p, tile, BN, CG, rank);
        const uint32_t buf = titer & 1;
        if (tile + cta_stride >= p.total_tiles) pdl_launch_dependents();  // last tile of this CTA (PDL, see common.cuh)
        mbar_wait(&tmem_empty_bar[buf], ((titer >> 1) & 1) ^ 1);  // epilogue drained this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int it = 0; it < t.num_kb; ++it, ++kiter) {
          const int s = kiter % STAGES;
          const uint32_t ph = (kiter / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t b_base = a_base + A_TILE_BYTES;
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4) {
            // K-major: advance 16 bf16 = 32 B inside the 128 B swizzle row.
            // MN-major: advance 16 k-rows = 2 swizzle atoms of 1024 B.
            const uint64_t adesc = A_MN ? umma_smem_desc(a_base + k4 * 2048, 64 * BK * 2, 1024)
                                        : umma_smem_desc(a_base + k4 * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc(b_base + k4 * 2048, 64 * BK * 2, 1024)
                                        : umma_smem_desc(b_base + k4 * 32, 16, 1024);
            if constexpr (PAIR) umma_bf16_ss_2sm(tmem_d, adesc, bdesc, idesc, (it > 0 || k4 > 0) ? 1u : 0u);
            else umma_bf16_ss(tmem_d, adesc, bdesc, idesc, (it > 0 || k4 > 0) ? 1u : 0u);
          }
          // frees the smem slot (of both CTAs) when these MMAs retire
          if constexpr (PAIR) umma_commit_2sm(&empty_bar[s]);
          else umma_commit(&empty_bar[s]);
        }
        // accumulator of this tile complete (in both CTAs' tensor memory)
        if constexpr (PAIR) umma_commit_2sm(&tmem_full_bar[buf]);
        else umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;        // 0..7
    const int grp = ew >> 2;        // epilogue group == TMEM buffer
    const int q = warp & 3;         // TMEM lane quarter this warp may access
    float* stg = reinterpret_cast<float*>(smem + L::STG_OFFSET + ew * L::STG_PER_WARP);
    // coalesced mapping inside a 32-row x 16-col chunk: 4 lanes cover one row (4 x 16 B)
    const int c_row = lane >> 2;    // + 8 * it
    const int c_pc = lane & 3;      // 16-byte piece -> columns 4*c_pc .. 4*c_pc+3
    if constexpr (SOFTMAX) {
      // -------- attention probabilities: all 8 warps work on every tile; the two warps of a lane quarter own
      // one query row each per lane and 128 of the 256 keys.  ONE pass over TMEM: the 128 logits stay in
      // registers, the accumulator buffer is handed back at once, row max / row sum are exchanged through smem.
      // v = alpha*acc + B[head, i, j]; causal: j > i -> -1e4 (vt_attention.py:63-74); the relative-position bias
      // is separable, so each thread keeps its row's bank slices in registers:
      // B[i, j] = bt[tj] + bh[hj] + bw[wj]  (get_B, vt_attention.py:169-174).
      constexpr int BT = EK == EK_SOFTMAX_1x16x16 ? 1 : 4;
      constexpr int BH = EK == EK_SOFTMAX_1x16x16 ? 16 : 8;
      constexpr int BW = EK == EK_SOFTMAX_1x16x16 ? 16 : 8;
      static_assert(BN == 256 || !SOFTMAX, "softmax epilogue needs BN == 256");
      const int half = ew >> 2;    // which 128 keys
      float* const xchg = reinterpret_cast<float*>(smem + L::XCHG_OFFSET);
      float* const x_own_max = xchg + half * 128 + q * 32 + lane;
      float* const x_oth_max = xchg + (half ^ 1) * 128 + q * 32 + lane;
      float* const x_own_sum = x_own_max + 256;
      float* const x_oth_sum = x_oth_max + 256;
      const float kLog2e = 1.4426950408889634f;
      const bool causal = (p.flags & LVT_GEMM_CAUSAL) != 0;
      const float a2 = p.alpha * kLog2e;
      const float kMasked = -1e4f * kLog2e;
      uint4* const slab = reinterpret_cast<uint4*>(stg);
      uint32_t titer = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++titer) {
        const TileCoord t = decode_tile(p, tile, BN);
        const uint32_t buf = titer & 1;
        const uint32_t taddr = tmem_base + buf * BN + (static_cast<uint32_t>(q * 32) << 16) + half * 128;
        const int row_base = t.m0 + q * 32;
        const int row = row_base + lane;  // 0..255 inside the block (M == 256)
        const int head = t.z % p.heads;
        const int ti = row / (BH * BW), hi = (row / BW) % BH, wi = row % BW;
        // bank slices of this row for the 128 keys of this half, pre-scaled into the log2 domain:
        // key k = half*128 + kk -> tj = k / (BH*BW), hj = (k / BW) % BH, wj = k % BW
        constexpr int TJN = (BH * BW <= 128) ? 128 / (BH * BW) : 1;  // distinct tj inside a half
        constexpr int HJN = (128 / BW < BH) ? 128 / BW : BH;         // distinct hj inside a half
        float btl[TJN], bhl[HJN], bw[BW];
#pragma unroll
        for (int x = 0; x < TJN; ++x) {
          const int tj = (half * 128) / (BH * BW) + x;
          btl[x] = kLog2e * __ldg(p.bank_t + head * (2 * BT - 1) + (ti - tj + BT - 1));
        }
#pragma unroll
        for (int y = 0; y < HJN; ++y) {
          const int hj = ((half * 128) / BW + y) % BH;
          bhl[y] = kLog2e * __ldg(p.bank_h + head * (2 * BH - 1) + (hi - hj + BH - 1));
        }
#pragma unroll
        for (int x = 0; x < BW; ++x) bw[x] = kLog2e * __ldg(p.bank_w + head * (2 * BW - 1) + (wi - x + BW - 1));
        mbar_wait(&tmem_full_bar[buf], (titer >> 1) & 1);
        tc_fence_after();
        // all four accumulator loads in flight at once; the logits replace the raw accumulators in place
        uint32_t racc[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32(taddr + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&racc[32 * c]));
        tmem_ld_wait();
        float* const l = reinterpret_cast<float*>(racc);  // logits in the log2 domain: (alpha*acc + B) * log2(e)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int kk = 0; kk < 128; ++kk) {
          const float bias = (btl[(kk / (BH * BW)) % TJN] + bhl[(kk / BW) % HJN]) + bw[kk % BW];
          float v = __uint_as_float(racc[kk]) * a2 + bias;
          if (causal && half * 128 + kk > row) v = kMasked;
          racc[kk] = __float_as_uint(v);
          m4[kk & 3] = fmaxf(m4[kk & 3], v);
        }
        // the accumulator buffer may be overwritten by the MMAs of the tile after next
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        *x_own_max = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        mx = fmaxf(mx, *x_oth_max);
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 128; ++i) {
          l[i] = fast_exp2(l[i] - mx);
          s4[i & 3] += l[i];
        }
        float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        *x_own_sum = sum;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        sum += *x_oth_sum;
        const float inv = 1.f / sum;
        if (p.lse && half == 0) p.lse[(long long)t.z * p.M + row] = (mx + log2f(sum)) * 0.6931471805599453f;
        // P -> 128B-swizzled slab (32 rows x 64 bf16) -> TMA store
        const int o_zlo = t.z % p.o_zdiv, o_zhi = t.z / p.o_zdiv;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          if (lane == 0) bulk_wait_group_read<0>();  // the previous store of this warp has drained the slab
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float* e = l + 64 * sl + 8 * k;
            uint4 u;
            u.x = pack_bf16x2(e[0] * inv, e[1] * inv);
            u.y = pack_bf16x2(e[2] * inv, e[3] * inv);
            u.z = pack_bf16x2(e[4] * inv, e[5] * inv);
            u.w = pack_bf16x2(e[6] * inv, e[7] * inv);
            slab[lane * 8 + (k ^ (lane & 7))] = u;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&tm_o, slab, half * 128 + 64 * sl, row_base, 0, o_zlo, o_zhi);
            bulk_commit_group();
          }
        }
      }
      if (lane == 0) bulk_wait_group<0>();  // all TMA stores of this warp are complete
    } else {
    uint32_t titer = grp;
    uint32_t c_phase = 0;  // parity of this warp's C-slab barrier
    (void)c_phase;
    uint32_t slab_it = 0;  // running slab count of this warp (selects the slab buffer)
    (void)slab_it;
    // (pair: the accumulator-drained arrive goes to the leader's barrier, which the MMA issuer waits on)
    const uint32_t te_remote[2] = {PAIR ? mapa_u32(&tmem_empty_bar[0], 0) : 0u, PAIR ? mapa_u32(&tmem_empty_bar[1], 0) : 0u};
    for (int tile = cta0 + grp * cta_stride; tile < p.total_tiles; tile += 2 * cta_stride, titer += 2) {
      const TileCoord t = decode_tile(p, tile, BN, CG, rank);
      const uint32_t buf = grp;
      const uint32_t taddr = tmem_base + buf * BN + (static_cast<uint32_t>(q * 32) << 16);
      const long long o_zbase = (long long)(t.z / p.o_zdiv) * p.o_s_zhi + (long long)(t.z % p.o_zdiv) * p.o_s_zlo;
      const int row_base = t.m0 + q * 32;
      if ((ST & ST_TMA) == 0 && t.num_kb > 0) {  // (the TMA epilogue waits after prefetching its C slab)
        mbar_wait(&tmem_full_bar[buf], (titer >> 1) & 1);
        tc_fence_after();
      }

      if constexpr ((EK == EK_LINEAR || EK == EK_DS) && (ST & ST_TMA) != 0) {
        // -------- TMA epilogue: TMEM -> registers -> 128B-swizzled smem slab (32 rows x 128 B per warp)
        // -> one cp.async.bulk.tensor store (or reduce-add) per slab.  An optional second tensor
        // (residual / mask source / P) arrives the same way, prefetched one slab ahead.
        constexpr bool F32 = (ST & ST_F32) != 0;
        constexpr bool CLOAD = (ST & ST_CLOAD) != 0;
        constexpr int SLAB_COLS = F32 ? 32 : 64;  // 128 B per row
        constexpr int NSLAB = L::NSLAB;
        const uint4* const cbuf = reinterpret_cast<const uint4*>(reinterpret_cast<uint8_t*>(stg) + NSLAB * STG_BYTES_PER_WARP);
        const int o_zlo = t.z % p.o_zdiv, o_zhi = t.z / p.o_zdiv;
        const float alpha = p.alpha;
        const bool relu = (p.flags & LVT_GEMM_RELU) != 0;
        const int ncols = min(BN, p.N - t.n0);
        float rd_acc = 0.f;  // LVT_GEMM_ROWDOT: running sum over the current block of rd_block columns
        (void)rd_acc;
        float dl = 0.f;
        if constexpr (EK == EK_DS) dl = (row_base + lane < p.M) ? p.delta[(long long)t.z * p.M + row_base + lane] : 0.f;
        auto issue_c = [&](int c0) {
          const int col0 = t.n0 + c0;
          mbar_arrive_expect_tx(&c_bar[ew], STG_BYTES_PER_WARP);
          tma_load_5d(const_cast<uint4*>(cbuf), &tm_c, &c_bar[ew], col0 % p.o_cin, row_base, col0 / p.o_cin, o_zlo, o_zhi);
        };
        if constexpr (CLOAD) {
          if (lane == 0) issue_c(0);  // overlaps the wait for the accumulator
        }
        if (t.num_kb > 0) {
          mbar_wait(&tmem_full_bar[buf], (titer >> 1) & 1);
          tc_fence_after();
        }
#pragma unroll 1
        for (int c0 = 0; c0 < ncols; c0 += SLAB_COLS) {
          const int col0 = t.n0 + c0;
          uint4 cr[8];  // this lane's row of the C slab (128 B)
          if constexpr (CLOAD) {
            mbar_wait(&c_bar[ew], c_phase);
            c_phase ^= 1;
#pragma unroll
            for (int k = 0; k < 8; ++k) cr[k] = cbuf[lane * 8 + (k ^ (lane & 7))];
            __syncwarp();
            if (lane == 0 && c0 + SLAB_COLS < ncols) issue_c(c0 + SLAB_COLS);
          }
          // the store that last used this slab buffer has finished reading it
          uint4* const slab = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(stg) + (slab_it % NSLAB) * STG_BYTES_PER_WARP);
          ++slab_it;
          if (lane == 0) bulk_wait_group_read<NSLAB - 1>();
          __syncwarp();
#pragma unroll
          for (int h = 0; h < SLAB_COLS / 32; ++h) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c0 + 32 * h, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * alpha;
            if constexpr (EK == EK_DS) {
              // dS = P * (alpha*acc - delta[row]); P = 64 bf16 of this row in cr[]
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t w[4] = {cr[4 * h + k].x, cr[4 * h + k].y, cr[4 * h + k].z, cr[4 * h + k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
                  v[8 * k + 2 * j] = pf.x * (v[8 * k + 2 * j] - dl);
                  v[8 * k + 2 * j + 1] = pf.y * (v[8 * k + 2 * j + 1] - dl);
                }
              }
            } else {
              if (p.bias) {
                const float* bp = p.bias + col0 + 32 * h;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + i));
                  v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                }
              }
              if constexpr (CLOAD && F32) {  // fp32 residual
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  v[4 * k] += __uint_as_float(cr[k].x); v[4 * k + 1] += __uint_as_float(cr[k].y);
                  v[4 * k + 2] += __uint_as_float(cr[k].z); v[4 * k + 3] += __uint_as_float(cr[k].w);
                }
              }
              if constexpr (CLOAD && !F32) {
                if (p.flags & LVT_GEMM_ROWDOT) {  // delta = rowsum(dO * O): this lane's row of both
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint32_t w[4] = {cr[4 * h + k].x, cr[4 * h + k].y, cr[4 * h + k].z, cr[4 * h + k].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
                      rd_acc = fmaf(v[8 * k + 2 * j], pf.x, rd_acc);
                      rd_acc = fmaf(v[8 * k + 2 * j + 1], pf.y, rd_acc);
                    }
                  }
                }
                if (p.flags & LVT_GEMM_AUX_ADD) {  // bf16 residual (ResBlock skip connection)
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint32_t w[4] = {cr[4 * h + k].x, cr[4 * h + k].y, cr[4 * h + k].z, cr[4 * h + k].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
                      v[8 * k + 2 * j] += pf.x;
                      v[8 * k + 2 * j + 1] += pf.y;
                    }
                  }
                }
              }
              if (relu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              }
              if constexpr (CLOAD && !F32) {  // ReLU backward: keep where the bf16 mask source is > 0
                if (p.flags & LVT_GEMM_MASK) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t w[4] = {cr[4 * h + k].x, cr[4 * h + k].y, cr[4 * h + k].z, cr[4 * h + k].w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint32_t lo = w[j] & 0xFFFFu, hi = w[j] >> 16;
                    if (!(lo != 0 && lo < 0x8000u)) v[8 * k + 2 * j] = 0.f;
                    if (!(hi != 0 && hi < 0x8000u)) v[8 * k + 2 * j + 1] = 0.f;
                  }
                }
                }
              }
            }
            if constexpr (!F32) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // 4 units of 8 bf16
                uint4 u;
                u.x = pack_bf16x2(v[8 * k], v[8 * k + 1]);
                u.y = pack_bf16x2(v[8 * k + 2], v[8 * k + 3]);
                u.z = pack_bf16x2(v[8 * k + 4], v[8 * k + 5]);
                u.w = pack_bf16x2(v[8 * k + 6], v[8 * k + 7]);
                slab[lane * 8 + ((4 * h + k) ^ (lane & 7))] = u;
              }
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k)  // 8 units of 4 fp32
                slab[lane * 8 + (k ^ (lane & 7))] =
                    make_uint4(__float_as_uint(v[4 * k]), __float_as_uint(v[4 * k + 1]),
                               __float_as_uint(v[4 * k + 2]), __float_as_uint(v[4 * k + 3]));
            }
          }
          if constexpr (CLOAD && !F32 && EK == EK_LINEAR) {
            if ((p.flags & LVT_GEMM_ROWDOT) && (col0 + SLAB_COLS) % p.rd_block == 0) {  // block of columns complete
              const int row = row_base + lane;
              if (row < p.M) {
                const int seq = row / p.rd_L;
                p.rowdot[((long long)seq * (p.N / p.rd_block) + col0 / p.rd_block) * p.rd_L + (row - seq * p.rd_L)] = rd_acc;
              }
              rd_acc = 0.f;
            }
          }
          fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            if constexpr ((ST & ST_REDUCE) != 0)
              tma_reduce_add_5d(&tm_o, slab, col0 % p.o_cin, row_base, col0 / p.o_cin, o_zlo, o_zhi);
            else
              tma_store_5d(&tm_o, slab, col0 % p.o_cin, row_base, col0 / p.o_cin, o_zlo, o_zhi);
            bulk_commit_group();
          }
        }
      } else if constexpr (EK == EK_LINEAR || EK == EK_DS) {
        // per-tile invariants (kept out of the chunk loop: no divisions, 32-bit offsets inside the tile)
        const long long tile_off = o_zbase + (long long)row_base * p.o_ld;
        float* const of32 = p.out_f32 ? p.out_f32 + tile_off : nullptr;
        __nv_bfloat16* const obf = p.out_bf16 ? p.out_bf16 + tile_off : nullptr;
        const float* const resp = p.res ? p.res + tile_off : nullptr;
        const __nv_bfloat16* const auxp = p.aux ? p.aux + tile_off : nullptr;
        const int ld = (int)p.o_ld;
        const int rows_valid = p.M - row_base;  // rows of this warp's quarter that exist
        const float* const deltap = (EK == EK_DS) ? p.delta + (long long)t.z * p.M + row_base : nullptr;
        const int flags = p.flags;
        const float alpha = p.alpha;
        float4* const stg4 = reinterpret_cast<float4*>(stg);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          const int col0 = t.n0 + c0;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r[16];
          if (flags & (1 << 29)) break;  // DEBUG: no epilogue work at all
          if (t.num_kb > 0) {
            tmem_ld_32x16(taddr + c0, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = 0;
          }
          if (flags & (1 << 30)) continue;  // DEBUG: TMEM reads only
          // lane == row: stage alpha*acc (4 swizzled 16 B units per row)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            stg4[stg_unit(lane, k)] =
                make_float4(__uint_as_float(r[4 * k]) * alpha, __uint_as_float(r[4 * k + 1]) * alpha,
                            __uint_as_float(r[4 * k + 2]) * alpha, __uint_as_float(r[4 * k + 3]) * alpha);
          __syncwarp();
          const int col = col0 + 4 * c_pc;
          const bool vec_ok = col + 4 <= p.N;
          const int o_col = p.o_shift < 0 ? col
                                          : (int)((long long)(col >> p.o_shift) * p.o_s_blk) + (col & (p.o_cin - 1));
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (EK == EK_LINEAR && p.bias && p.bias_mod == 0 && vec_ok)
            bias4 = *reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rl = c_row + 8 * it;
            float4 v = stg4[stg_unit(rl, c_pc)];
            if (rl >= rows_valid || col >= p.N) continue;
            const int off = rl * ld + o_col;
            if (vec_ok) {
              if constexpr (EK == EK_DS) {
                const float dl = deltap[rl];
                const uint2 pu = *reinterpret_cast<const uint2*>(auxp + off);
                const float2 p0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pu.x));
                const float2 p1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pu.y));
                uint2 u;
                u.x = pack_bf16x2(p0.x * (v.x - dl), p0.y * (v.y - dl));
                u.y = pack_bf16x2(p1.x * (v.z - dl), p1.y * (v.w - dl));
                *reinterpret_cast<uint2*>(obf + off) = u;
              } else {
                if (p.bias) {
                  if (p.bias_mod > 0)
                    bias4 = *reinterpret_cast<const float4*>(p.bias + (long long)((row_base + rl) % p.bias_mod) * p.N + col);
                  v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
                }
                if (resp) {
                  const float4 r4 = *reinterpret_cast<const float4*>(resp + off);
                  v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
                }
                if (flags & LVT_GEMM_AUX_ADD) {
                  const uint2 au = *reinterpret_cast<const uint2*>(auxp + off);
                  const float2 q0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&au.x));
                  const float2 q1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&au.y));
                  v.x += q0.x; v.y += q0.y; v.z += q1.x; v.w += q1.y;
                }
                if (flags & LVT_GEMM_RELU) {
                  v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                }
                if (flags & LVT_GEMM_MASK) {
                  const uint2 mu = *reinterpret_cast<const uint2*>(auxp + off);
                  // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                  const uint32_t a0 = mu.x & 0xFFFFu, a1 = mu.x >> 16, a2 = mu.y & 0xFFFFu, a3 = mu.y >> 16;
                  if (!(a0 != 0 && a0 < 0x8000u)) v.x = 0.f;
                  if (!(a1 != 0 && a1 < 0x8000u)) v.y = 0.f;
                  if (!(a2 != 0 && a2 < 0x8000u)) v.z = 0.f;
                  if (!(a3 != 0 && a3 < 0x8000u)) v.w = 0.f;
                }
                if (of32) {
                  if (flags & LVT_GEMM_ATOMIC) red_add_v4(of32 + off, v.x, v.y, v.z, v.w);
                  else *reinterpret_cast<float4*>(of32 + off) = v;
                }
                if (obf) {
                  uint2 u;
                  u.x = pack_bf16x2(v.x, v.y);
                  u.y = pack_bf16x2(v.z, v.w);
                  *reinterpret_cast<uint2*>(obf + off) = u;
                }
              }
            } else {
              // ragged N tail: element-wise
              float ve[4] = {v.x, v.y, v.z, v.w};
              for (int e = 0; e < 4 && col + e < p.N; ++e) {
                float x = ve[e];
                if constexpr (EK == EK_DS) {
                  x = __bfloat162float(auxp[off + e]) * (x - deltap[rl]);
                  obf[off + e] = __float2bfloat16(x);
                } else {
                  if (p.bias)
                    x += p.bias[(p.bias_mod > 0 ? (long long)((row_base + rl) % p.bias_mod) * p.N : 0) + col + e];
                  if (resp) x += resp[off + e];
                  if (flags & LVT_GEMM_AUX_ADD) x += __bfloat162float(auxp[off + e]);
                  if (flags & LVT_GEMM_RELU) x = fmaxf(x, 0.f);
                  if ((flags & LVT_GEMM_MASK) && !(__bfloat162float(auxp[off + e]) > 0.f)) x = 0.f;
                  if (of32) {
                    if (flags & LVT_GEMM_ATOMIC) atomicAdd(of32 + off + e, x);
                    else of32[off + e] = x;
                  }
                  if (obf) obf[off + e] = __float2bfloat16(x);
                }
              }
            }
          }
          __syncwarp();
        }
      }
      // this warp no longer reads the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && rank != 0) mbar_arrive_cluster(te_remote[buf]);
        else mbar_arrive(&tmem_empty_bar[buf]);
      }
    }
    if ((ST & ST_TMA) != 0 && lane == 0) bulk_wait_group<0>();  // all TMA stores of this warp are complete
    }  // !SOFTMAX
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync();  // no CTA leaves (or frees tensor memory) while its peer may still use its smem / barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, 2 * BN);
    else tmem_dealloc(tmem_base, 2 * BN);
  }
}


// ----------------------------------------------------------------------------------------
// Fused attention forward for one 256-token block per (sequence, head):
//   S = Q K^T (tcgen05, 128 x 256 tile, K = da = 128)  ->  P = softmax(alpha*S + B [causal -1e4])  ->  O = P V
// (ScaledDotProductAttention.forward, vt_attention.py:61-81, with the bias of BlockLocalAttention.get_B).
// P never leaves the SM between the softmax and the second MMA: the epilogue warps write it as the
// 128B-swizzled K-major A operand (4 k-blocks of 128 rows x 64 keys) straight into shared memory, the same bytes
// are TMA-stored to HBM for the backward (optional).  Saves the P re-read of a separate P V GEMM and one launch.
//   warp 0: TMA producer (Q, K tile; V tile)        warp 1: MMA issuer (QK^T of tile i+1 is issued before PV of tile i)
//   warps 2-9: softmax + O epilogue; the two warps of a lane quarter own one query row per lane and 128 keys each
// TMEM: S at columns [0, 256), O at [256, 384).
// ----------------------------------------------------------------------------------------
struct AttnSmem {
  static constexpr int Q_OFF = 0;                  // 2 k-blocks x (128 rows x 128 B)
  static constexpr int K_OFF = 32768;              // 2 k-blocks x (256 rows x 128 B)
  static constexpr int V_OFF = 98304;              // 4 key-blocks x 2 x (64 keys x 128 B)   (MN-major B operand)
  static constexpr int P_OFF = 163840;             // 4 key-blocks x (128 rows x 128 B)      (K-major A operand)
  static constexpr int XCHG_OFF = 229376;          // [2][2][128] fp32
  static constexpr int BAR_OFF = XCHG_OFF + 2048;
  static constexpr int BINS_OFF = BAR_OFF + 256;   // [2][64] fp32 bank-gradient bins (backward)
  static constexpr int TOTAL = BINS_OFF + 512;
};
static_assert(AttnSmem::TOTAL <= 232448, "attention forward: shared memory budget");

// EK == EK_DS turns the same pipeline into the first half of the attention BACKWARD:
//   dP = dO V^T (first MMA)  ->  dS = P * (dP - delta) (P TMA-loaded into the slabs, overwritten in place)  ->
//   dQ = alpha * dS K (second MMA, K as the MN-major operand); dS is TMA-stored for the dK GEMM and the bank gradient.
template <int EK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_p,
                const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_c, const GemmParams p, int v_cin, int v_zdiv, int o2_cin,
                int o2_zdiv, int store_p, int bank_fused) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using A = AttnSmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A::BAR_OFF);
  uint64_t* qk_full = bars + 0;
  uint64_t* qk_empty = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* s_empty = bars + 3;
  uint64_t* v_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* pv_done = bars + 6;
  uint64_t* c_bar = bars + 16;  // [NUM_EPI_WARPS] P slab loads (EK_DS)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);
  constexpr bool BWD = EK == EK_DS;
  float* const bank_bins = reinterpret_cast<float*>(smem + AttnSmem::BINS_OFF);  // [2][64] (EK_DS, fused bank gradient)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("lvt_b200: attention forward needs 1024-byte aligned dynamic shared memory\n");
      __trap();
    }
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(qk_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, NUM_EPI_WARPS);
    mbar_init(v_full, 1);
    mbar_init(p_full, NUM_EPI_WARPS);
    mbar_init(pv_done, 1);
#pragma unroll
    for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(&c_bar[w], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 128) bank_bins[threadIdx.x] = 0.f;
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const TileCoord t = decode_tile(p, tile, 256);
        const int a_zlo = t.z % p.a_zdiv, a_zhi = t.z / p.a_zdiv;
        const int b_zlo = t.z % p.b_zdiv, b_zhi = t.z / p.b_zdiv;
        mbar_wait(qk_empty, (it & 1) ^ 1);  // QK^T of the previous tile has consumed the operands
        mbar_arrive_expect_tx(qk_full, 98304);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const int k0 = kb * BK;
          tma_load_5d(smem + A::Q_OFF + kb * 16384, &tm_q, qk_full, k0 % p.a_cin, t.m0, k0 / p.a_cin, a_zlo, a_zhi);
          tma_load_5d(smem + A::K_OFF + kb * 32768, &tm_k, qk_full, k0 % p.b_cin, 0, k0 / p.b_cin, b_zlo, b_zhi);
        }
        const int v_zlo = t.z % v_zdiv, v_zhi = t.z / v_zdiv;
        mbar_wait(pv_done, (it & 1) ^ 1);   // P V of the previous tile has consumed V
        mbar_arrive_expect_tx(v_full, 65536);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c = 64 * j;
            tma_load_5d(smem + A::V_OFF + kb * 16384 + j * 8192, &tm_v, v_full, c % v_cin, kb * BK, c / v_cin, v_zlo,
                        v_zhi);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc(BM, 256, /*bf16*/ 1, false, false);
      constexpr uint32_t idesc_pv = umma_idesc(BM, 128, /*bf16*/ 1, false, true);
      const uint32_t q_base = smem_u32(smem + A::Q_OFF), k_base = smem_u32(smem + A::K_OFF);
      const uint32_t v_base = smem_u32(smem + A::V_OFF), p_base = smem_u32(smem + A::P_OFF);
      auto issue_qk = [&](uint32_t it) {
        mbar_wait(qk_full, it & 1);
        mbar_wait(s_empty, (it & 1) ^ 1);  // the softmax warps hold the previous S in registers
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_bf16_ss(tmem_base, umma_smem_desc(q_base + kb * 16384 + k4 * 32, 16, 1024),
                         umma_smem_desc(k_base + kb * 32768 + k4 * 32, 16, 1024), idesc_qk, (kb | k4) ? 1u : 0u);
        }
        umma_commit(qk_empty);
        umma_commit(s_full);
      };
      const int my_tiles = blockIdx.x < p.total_tiles ? (p.total_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
      if (my_tiles > 0) issue_qk(0);
      for (int it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) issue_qk(it + 1);  // next tile's scores while this tile's softmax runs
        else pdl_launch_dependents();
        mbar_wait(v_full, it & 1);
        mbar_wait(p_full, it & 1);               // P of this tile is in shared memory (and O of the previous one was read)
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_bf16_ss(tmem_base + 256, umma_smem_desc(p_base + kb * 16384 + k4 * 32, 16, 1024),
                         umma_smem_desc(v_base + kb * 16384 + k4 * 2048, 8192, 1024), idesc_pv, (kb | k4) ? 1u : 0u);
        }
        umma_commit(pv_done);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + O epilogue
    constexpr int BT = EK == EK_SOFTMAX_4x8x8 ? 4 : 1;
    constexpr int BH = EK == EK_SOFTMAX_4x8x8 ? 8 : 16;
    constexpr int BW = EK == EK_SOFTMAX_4x8x8 ? 8 : 16;
    constexpr int TJN = (BH * BW <= 128) ? 128 / (BH * BW) : 1;
    constexpr int HJN = (128 / BW < BH) ? 128 / BW : BH;
    const int ew = warp - 2;
    const int q = warp & 3;      // TMEM lane quarter
    const int half = ew >> 2;    // which 128 keys (softmax) / which 64 output columns (O)
    float* const xchg = reinterpret_cast<float*>(smem + A::XCHG_OFF);
    float* const x_own_max = xchg + half * 128 + q * 32 + lane;
    float* const x_oth_max = xchg + (half ^ 1) * 128 + q * 32 + lane;
    float* const x_own_sum = x_own_max + 256;
    float* const x_oth_sum = x_oth_max + 256;
    const float kLog2e = 1.4426950408889634f;
    const bool causal = (p.flags & LVT_GEMM_CAUSAL) != 0;
    const float a2 = p.alpha * kLog2e;
    const float kMasked = -1e4f * kLog2e;
    // this warp's two P slabs: key-blocks 2*half and 2*half + 1, rows 32q .. 32q+31 of the A operand
    uint4* const slab0 = reinterpret_cast<uint4*>(smem + A::P_OFF + (2 * half) * 16384 + q * 4096);
    uint4* const slab1 = reinterpret_cast<uint4*>(smem + A::P_OFF + (2 * half + 1) * 16384 + q * 4096);
    uint32_t it = 0;
#define APROF(slot)                                                                                       \
  do {                                                                                                    \
    if (p.prof && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 32) p.prof[it * 8 + (slot)] = clock64(); \
  } while (0)
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const TileCoord t = decode_tile(p, tile, 256);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 128;
      const int row_base = t.m0 + q * 32;
      const int row = row_base + lane;
      APROF(0);
      float rs_h[8], cs_w[16];  // bank-gradient partial sums of this row (EK_DS)
      (void)rs_h;
      (void)cs_w;
      if constexpr (BWD) {
        // P slabs of this tile -> the slabs (the previous tile's second MMA is complete: this warp waited for pv_done
        // in its dQ epilogue; its own TMA stores out of the slabs must have finished reading)
        const int o_zlo = t.z % p.o_zdiv, o_zhi = t.z / p.o_zdiv;
        if (lane == 0) {
          bulk_wait_group_read<0>();
          mbar_arrive_expect_tx(&c_bar[ew], 8192);
          tma_load_5d(slab0, &tm_c, &c_bar[ew], half * 128, row_base, 0, o_zlo, o_zhi);
          tma_load_5d(slab1, &tm_c, &c_bar[ew], half * 128 + 64, row_base, 0, o_zlo, o_zhi);
        }
        const float dl = p.delta[(long long)t.z * p.M + row];
        // relative-position bank gradient of block (1,16,16) (get_B, vt_attention.py:169-174): this row's sums of
        // dS over the 16 key columns of each of its 8 key rows (-> dh_bank) and over the 8 key rows per column (-> dw_bank)
#pragma unroll
        for (int y = 0; y < 8; ++y) rs_h[y] = 0.f;
#pragma unroll
        for (int x = 0; x < 16; ++x) cs_w[x] = 0.f;
        mbar_wait(s_full, it & 1);
        tc_fence_after();
        mbar_wait(&c_bar[ew], it & 1);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          uint4* const slab = sl ? slab1 : slab0;
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + 64 * sl, r0);
          tmem_ld_32x32(taddr + 64 * sl + 32, r1);
          tmem_ld_wait();
          if (sl == 1) {  // the whole accumulator row is in registers: dP may be overwritten by the next tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t* acc = k < 4 ? &r0[8 * k] : &r1[8 * (k - 4)];
            uint4 pu = slab[lane * 8 + (k ^ (lane & 7))];
            uint32_t w[4] = {pu.x, pu.y, pu.z, pu.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
              const float v0 = pf.x * (__uint_as_float(acc[2 * j]) - dl), v1 = pf.y * (__uint_as_float(acc[2 * j + 1]) - dl);
              w[j] = pack_bf16x2(v0, v1);
              rs_h[4 * sl + (k >> 1)] += v0 + v1;        // key (half*128 + 64 sl + 8 k + 2 j + e) = (hj, wj)
              cs_w[8 * (k & 1) + 2 * j] += v0;
              cs_w[8 * (k & 1) + 2 * j + 1] += v1;
            }
            slab[lane * 8 + (k ^ (lane & 7))] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      } else {
      const int head = t.z % p.heads;
      const int ti = row / (BH * BW), hi = (row / BW) % BH, wi = row % BW;
      float btl[TJN], bhl[HJN], bw[BW];
#pragma unroll
      for (int x = 0; x < TJN; ++x) {
        const int tj = (half * 128) / (BH * BW) + x;
        btl[x] = kLog2e * __ldg(p.bank_t + head * (2 * BT - 1) + (ti - tj + BT - 1));
      }
#pragma unroll
      for (int y = 0; y < HJN; ++y) {
        const int hj = ((half * 128) / BW + y) % BH;
        bhl[y] = kLog2e * __ldg(p.bank_h + head * (2 * BH - 1) + (hi - hj + BH - 1));
      }
#pragma unroll
      for (int x = 0; x < BW; ++x) bw[x] = kLog2e * __ldg(p.bank_w + head * (2 * BW - 1) + (wi - x + BW - 1));
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      APROF(1);
      uint32_t racc[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(taddr + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&racc[32 * c]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);  // S may be overwritten by the next tile's QK^T
      APROF(2);
      float* const l = reinterpret_cast<float*>(racc);
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int kk = 0; kk < 128; ++kk) {
        const float bias = (btl[(kk / (BH * BW)) % TJN] + bhl[(kk / BW) % HJN]) + bw[kk % BW];
        float v = __uint_as_float(racc[kk]) * a2 + bias;
        if (causal && half * 128 + kk > row) v = kMasked;
        racc[kk] = __float_as_uint(v);
        m4[kk & 3] = fmaxf(m4[kk & 3], v);
      }
      float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      *x_own_max = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      mx = fmaxf(mx, *x_oth_max);
      APROF(3);
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 128; ++i) {
        l[i] = fast_exp2(l[i] - mx);
        s4[i & 3] += l[i];
      }
      float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      *x_own_sum = sum;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      sum += *x_oth_sum;
      APROF(4);
      const float inv = 1.f / sum;
      if (p.lse && half == 0) p.lse[(long long)t.z * p.M + row] = (mx + log2f(sum)) * 0.6931471805599453f;
      // P slabs (the previous tile's P V has completed: this warp waited for pv_done in its O epilogue; its own
      // TMA stores out of the slabs must have finished reading)
      if (lane == 0) bulk_wait_group_read<0>();
      __syncwarp();
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        uint4* const slab = sl ? slab1 : slab0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float* e = l + 64 * sl + 8 * k;
          uint4 u;
          u.x = pack_bf16x2(e[0] * inv, e[1] * inv);
          u.y = pack_bf16x2(e[2] * inv, e[3] * inv);
          u.z = pack_bf16x2(e[4] * inv, e[5] * inv);
          u.w = pack_bf16x2(e[6] * inv, e[7] * inv);
          slab[lane * 8 + (k ^ (lane & 7))] = u;
        }
      }
      }
      fence_proxy_async();  // generic-proxy writes -> visible to tcgen05.mma and to the TMA store
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(p_full);
        if (store_p) {
          const int o_zlo = t.z % p.o_zdiv, o_zhi = t.z / p.o_zdiv;
          tma_store_5d(&tm_p, slab0, half * 128, row_base, 0, o_zlo, o_zhi);
          tma_store_5d(&tm_p, slab1, half * 128 + 64, row_base, 0, o_zlo, o_zhi);
          bulk_commit_group();
        }
      }
      if constexpr (BWD) {
        if (bank_fused) {
          // fills the wait for the second MMA: warp / CTA reduction of the bank partial sums, one flush per tile
          float* const bins = bank_bins + (it & 1) * 64;  // [0,31) dh, [31,62) dw, [62] dt
          const int hi = (row >> 4) & 15, wi = row & 15;
          float tot = 0.f;
#pragma unroll
          for (int y = 0; y < 8; ++y) {
            float v = rs_h[y];
            tot += v;
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if ((lane & 15) == 0) atomicAdd(&bins[hi - (half * 8 + y) + 15], v);
          }
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const float v = cs_w[x] + __shfl_xor_sync(0xffffffffu, cs_w[x], 16);
            if (lane < 16) atomicAdd(&bins[31 + wi - x + 15], v);  // distinct bins across the 16 lanes
          }
          tot = warp_sum(tot);
          if (lane == 0) atomicAdd(&bins[62], tot);
          asm volatile("bar.sync 5, 256;" ::: "memory");  // the 8 epilogue warps
          if (ew == 0) {
            const int head = t.z % p.heads;
#pragma unroll
            for (int i = lane; i < 63; i += 32) {
              const float v = bins[i];
              bins[i] = 0.f;  // this buffer is added to again two tiles from now, after the next barrier
              float* dst = i < 31 ? const_cast<float*>(p.bank_h) + head * 31 + i
                                  : (i < 62 ? const_cast<float*>(p.bank_w) + head * 31 + (i - 31) : const_cast<float*>(p.bank_t) + head);
              atomicAdd(dst, v);
            }
          }
        }
      }
      // O = P V: 64 of the 128 output columns per warp
      APROF(5);
      mbar_wait(pv_done, it & 1);
      tc_fence_after();
      APROF(6);
      {
        uint32_t r0[32], r1[32];
        const uint32_t taddr_o = tmem_base + 256 + (static_cast<uint32_t>(q * 32) << 16) + half * 64;
        tmem_ld_32x32(taddr_o, r0);
        tmem_ld_32x32(taddr_o + 32, r1);
        tmem_ld_wait();
        if constexpr (BWD) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            r0[i] = __float_as_uint(__uint_as_float(r0[i]) * p.alpha);
            r1[i] = __float_as_uint(__uint_as_float(r1[i]) * p.alpha);
          }
        }
        if (lane == 0) bulk_wait_group_read<0>();  // the P stores have drained slab0
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r0[8 * k]), __uint_as_float(r0[8 * k + 1]));
          u.y = pack_bf16x2(__uint_as_float(r0[8 * k + 2]), __uint_as_float(r0[8 * k + 3]));
          u.z = pack_bf16x2(__uint_as_float(r0[8 * k + 4]), __uint_as_float(r0[8 * k + 5]));
          u.w = pack_bf16x2(__uint_as_float(r0[8 * k + 6]), __uint_as_float(r0[8 * k + 7]));
          slab0[lane * 8 + (k ^ (lane & 7))] = u;
          u.x = pack_bf16x2(__uint_as_float(r1[8 * k]), __uint_as_float(r1[8 * k + 1]));
          u.y = pack_bf16x2(__uint_as_float(r1[8 * k + 2]), __uint_as_float(r1[8 * k + 3]));
          u.z = pack_bf16x2(__uint_as_float(r1[8 * k + 4]), __uint_as_float(r1[8 * k + 5]));
          u.w = pack_bf16x2(__uint_as_float(r1[8 * k + 6]), __uint_as_float(r1[8 * k + 7]));
          slab0[lane * 8 + ((4 + k) ^ (lane & 7))] = u;
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          const int col = half * 64;
          tma_store_5d(&tm_o, slab0, col % o2_cin, row_base, col / o2_cin, t.z % o2_zdiv, t.z / o2_zdiv);
          bulk_commit_group();
        }
      }
      APROF(7);
    }
    if (lane == 0) bulk_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------------------
// host: TMA descriptor construction (driver entry point fetched at run time: no -lcuda)
// ----------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

struct MapKey {
  const void* base;
  long long v[12];
  bool operator==(const MapKey& o) const {
    return base == o.base && memcmp(v, o.v, sizeof(v)) == 0;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.base);
    for (int i = 0; i < 12; ++i) h = h * 1000003u ^ std::hash<long long>()(k.v[i]);
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
std::mutex g_map_mutex;

// Builds the 5-D bf16 tensor map of one operand:
//   dims    (cin, r_extent, n_blocks, zlo, zhi)
//   strides (    ld,        s_blk,    s_zlo, s_zhi)   in elements
//   box     (64, box_rows, 1, 1, 1), SWIZZLE_128B
int make_operand_map(CUtensorMap* out, const void* base, long long c_extent, long long r_extent,
                     int cin, long long ld, long long s_blk, int batch, int zdiv, long long s_zlo,
                     long long s_zhi, int box_rows, int esize = 2) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    lvt_set_error("cuTensorMapEncodeTiled driver entry point not available");
    return LVT_ERR_CUDA;
  }
  const long long nblk = (c_extent + cin - 1) / cin;
  const long long zlo = batch < zdiv ? batch : zdiv;
  const long long zhi = (batch + zdiv - 1) / zdiv;
  MapKey key;
  key.base = base;
  const long long kv[12] = {c_extent, r_extent, cin, ld, s_blk, batch, zdiv, s_zlo, s_zhi, box_rows, esize, 0};
  memcpy(key.v, kv, sizeof(kv));
  {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
      *out = it->second;
      return LVT_OK;
    }
  }
  cuuint64_t dims[5] = {(cuuint64_t)(cin < c_extent ? cin : c_extent), (cuuint64_t)r_extent,
                        (cuuint64_t)nblk, (cuuint64_t)zlo, (cuuint64_t)zhi};
  auto fix = [esize](long long s) -> cuuint64_t { return (cuuint64_t)((s > 0 ? s : 8) * esize); };
  cuuint64_t strides[4] = {fix(ld), fix(nblk > 1 ? s_blk : 8), fix(zlo > 1 ? s_zlo : 8),
                           fix(zhi > 1 ? s_zhi : 8)};
  for (int i = 0; i < 4; ++i) {
    if (strides[i] % 16 != 0) {
      lvt_set_error("GEMM operand stride %d (= %llu bytes) is not a multiple of 16 bytes", i,
                    (unsigned long long)strides[i]);
      return LVT_ERR_INVALID;
    }
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    lvt_set_error("GEMM operand base pointer is not 16-byte aligned");
    return LVT_ERR_INVALID;
  }
  cuuint32_t box[5] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(out, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                   const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    lvt_set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %llu,%llu,%llu,%llu,%llu "
                  "strides %llu,%llu,%llu,%llu box_rows %d)",
                  (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)dims[2], (unsigned long long)dims[3],
                  (unsigned long long)dims[4], (unsigned long long)strides[0],
                  (unsigned long long)strides[1], (unsigned long long)strides[2],
                  (unsigned long long)strides[3], box_rows);
    return LVT_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    if (g_map_cache.size() > 4096) g_map_cache.clear();
    g_map_cache.emplace(key, *out);
  }
  return LVT_OK;
}

// NHWC activation tensor (phase, n, h, w, c) for the implicit-GEMM convolution modes:
// box = (64 channels, W, box_h rows, 1, 1); out-of-range pixels are zero-filled by TMA (= conv padding).
int make_conv_map(CUtensorMap* out, const void* base, int C, int W, int H, int N, int P, long long pix_stride,
                  long long phase_stride, int box_h) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    lvt_set_error("cuTensorMapEncodeTiled driver entry point not available");
    return LVT_ERR_CUDA;
  }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)(P > 0 ? P : 1)};
  cuuint64_t strides[4] = {(cuuint64_t)pix_stride * 2, (cuuint64_t)pix_stride * W * 2,
                           (cuuint64_t)pix_stride * W * H * 2,
                           (cuuint64_t)(P > 1 ? phase_stride : pix_stride * W * H * (long long)N) * 2};
  for (int i = 0; i < 4; ++i)
    if (strides[i] % 16 != 0) {
      lvt_set_error("conv operand stride %d is not a multiple of 16 bytes", i);
      return LVT_ERR_INVALID;
    }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    lvt_set_error("conv operand base pointer is not 16-byte aligned");
    return LVT_ERR_INVALID;
  }
  cuuint32_t box[5] = {64, (cuuint32_t)W, (cuuint32_t)box_h, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    lvt_set_error("cuTensorMapEncodeTiled (conv) failed with CUresult %d", (int)r);
    return LVT_ERR_CUDA;
  }
  return LVT_OK;
}

struct Maps {
  CUtensorMap a, b, o, c;
};

template <int BN, bool A_MN, bool B_MN, int EK, int ST = 0, int CG = 1>
int launch_gemm(const Maps& m, const GemmParams& p, int grid, cudaStream_t stream) {
  using L = SmemLayout<BN, ST, EK, CG>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, EK, ST, CG>;
  static bool configured = false;
  if (!configured) {
    LVT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  LVT_CHECK_CUDA(lvt_launch_cluster(CG, kern, dim3(grid), dim3(NUM_THREADS), L::TOTAL, stream, m.a, m.b, m.o, m.c, p));
  lvt_count_launch(1);
  return LVT_OK;
}

template <int BN, int EK, int ST, int CG = 1>
int dispatch_major(const Maps& m, const GemmParams& p, int grid, bool a_mn, bool b_mn, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false, EK, ST, CG>(m, p, grid, stream);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true, EK, ST, CG>(m, p, grid, stream);
  if (a_mn && !b_mn) return launch_gemm<BN, true, false, EK, ST, CG>(m, p, grid, stream);
  return launch_gemm<BN, true, true, EK, ST, CG>(m, p, grid, stream);
}

template <int BN, int CG = 1>
int dispatch_store(const Maps& m, const GemmParams& p, int grid, bool a_mn, bool b_mn, int st,
                   cudaStream_t stream) {
  switch (st) {
    case ST_TMA: return dispatch_major<BN, EK_LINEAR, ST_TMA, CG>(m, p, grid, a_mn, b_mn, stream);
    case ST_TMA | ST_F32: return dispatch_major<BN, EK_LINEAR, ST_TMA | ST_F32, CG>(m, p, grid, a_mn, b_mn, stream);
    case ST_TMA | ST_F32 | ST_REDUCE:
      return dispatch_major<BN, EK_LINEAR, ST_TMA | ST_F32 | ST_REDUCE, CG>(m, p, grid, a_mn, b_mn, stream);
    case ST_TMA | ST_F32 | ST_CLOAD:
      return dispatch_major<BN, EK_LINEAR, ST_TMA | ST_F32 | ST_CLOAD, CG>(m, p, grid, a_mn, b_mn, stream);
    case ST_TMA | ST_CLOAD: return dispatch_major<BN, EK_LINEAR, ST_TMA | ST_CLOAD, CG>(m, p, grid, a_mn, b_mn, stream);
    default: return dispatch_major<BN, EK_LINEAR, 0, CG>(m, p, grid, a_mn, b_mn, stream);
  }
}

// CTA pairs (cta_group::2, 256 x 256 tiles) for the linear GEMMs: LVT_GEMM_CG=1 in the environment keeps every
// GEMM on single-CTA 128 x 256 tiles (A/B timing aid).
bool pairs_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LVT_GEMM_CG");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  const int lim = lvt_sm_limit();  // (lvt_set_sm_limit: SMs left to a concurrent collective)
  return (lim > 0 && lim < n) ? lim : n;
}

}  // namespace

// (shared with the other tcgen05 kernels of the library: attn_bwd.cu)
int lvt_make_operand_map(CUtensorMap* out, const void* base, long long c_extent, long long r_extent, int cin,
                         long long ld, long long s_blk, int batch, int zdiv, long long s_zlo, long long s_zhi,
                         int box_rows, int esize) {
  return make_operand_map(out, base, c_extent, r_extent, cin, ld, s_blk, batch, zdiv, s_zlo, s_zhi, box_rows, esize);
}

extern "C" int lvt_gemm_bf16(const LvtGemm* g, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LVT_CHECK_ARG(g != nullptr, "lvt_gemm_bf16: null descriptor");
  LVT_CHECK_ARG(g->M > 0 && g->N > 0 && g->K > 0 && g->batch > 0, "lvt_gemm_bf16: bad shape %dx%dx%d batch %d",
                g->M, g->N, g->K, g->batch);
  LVT_CHECK_ARG(g->a && g->b, "lvt_gemm_bf16: null operand");
  LVT_CHECK_ARG(g->out_f32 || g->out_bf16 || (g->mode == LVT_EPI_SOFTMAX && g->v), "lvt_gemm_bf16: no output");
  // splits < 0: split-K factor chosen here (fills the SMs / SM pairs once, >= 4 k-blocks per split)
  const bool auto_split = g->splits < 0;
  int splits = g->splits > 0 ? g->splits : 1;
  LVT_CHECK_ARG((splits == 1 && !auto_split) || ((g->flags & LVT_GEMM_ATOMIC) && !g->out_bf16 && g->mode == LVT_EPI_LINEAR),
                "lvt_gemm_bf16: split-K needs LVT_GEMM_ATOMIC fp32 output only");
  LVT_CHECK_ARG(!(g->flags & LVT_GEMM_ATOMIC) || (g->out_f32 && !g->res),
                "lvt_gemm_bf16: atomic output needs out_f32 and no residual");
  LVT_CHECK_ARG(g->a_cin > 0 && g->b_cin > 0 && g->o_cin > 0 && g->a_zdiv > 0 && g->b_zdiv > 0 && g->o_zdiv > 0,
                "lvt_gemm_bf16: cin/zdiv must be positive");
  LVT_CHECK_ARG(g->o_cin % 16 == 0 || g->o_cin >= g->N, "lvt_gemm_bf16: o_cin must be a multiple of 16");
  LVT_CHECK_ARG(g->o_ld % 4 == 0 && g->o_s_blk % 4 == 0 && g->o_s_zlo % 4 == 0 && g->o_s_zhi % 4 == 0,
                "lvt_gemm_bf16: output strides must be multiples of 4 elements");
  if (g->flags & (LVT_GEMM_MASK | LVT_GEMM_AUX_ADD)) LVT_CHECK_ARG(g->aux_bf16, "lvt_gemm_bf16: MASK / AUX_ADD need aux_bf16");
  if (g->flags & LVT_GEMM_ROWDOT)
    LVT_CHECK_ARG(g->aux_bf16 && g->rowdot && g->out_bf16 && !g->out_f32 && !g->res && g->batch == 1 && splits == 1 &&
                      g->mode == LVT_EPI_LINEAR && (g->rd_block == 64 || g->rd_block == 128) && g->N % g->rd_block == 0 &&
                      g->rd_L > 0 && g->M % g->rd_L == 0,
                  "lvt_gemm_bf16: ROWDOT needs aux_bf16, rowdot, a single bf16 output, batch 1, rd_block in {64, 128} dividing N, rd_L | M");

  int bn = 128;
  int ek = EK_LINEAR;
  if (g->mode == LVT_EPI_SOFTMAX) {
    LVT_CHECK_ARG(g->N == 256 && (g->out_bf16 || g->v) && g->bank_t && g->bank_h && g->bank_w && g->M == 256 && g->heads > 0,
                  "lvt_gemm_bf16: SOFTMAX mode needs M == N == 256, banks and out_bf16");
    if (g->v)
      LVT_CHECK_ARG(g->K == 128 && g->o2_bf16 && g->o2_n == 128 && g->v_cin > 0 && g->v_zdiv > 0 && g->o2_cin >= 64 &&
                        g->o2_cin % 64 == 0 && g->o2_zdiv > 0 && g->a_cin % 64 == 0 && g->b_cin % 64 == 0 &&
                        g->v_cin % 64 == 0,
                    "lvt_gemm_bf16: fused attention forward needs K == o2_n == 128, o2_bf16 and 64-aligned blockings");
    if (g->bt == 1 && g->bh == 16 && g->bw == 16) ek = EK_SOFTMAX_1x16x16;
    else if (g->bt == 4 && g->bh == 8 && g->bw == 8) ek = EK_SOFTMAX_4x8x8;
    else {
      lvt_set_error("lvt_gemm_bf16: SOFTMAX mode supports attention blocks (1,16,16) and (4,8,8), got (%d,%d,%d)",
                    g->bt, g->bh, g->bw);
      return LVT_ERR_INVALID;
    }
    LVT_CHECK_ARG(!g->a_mn_major && !g->b_mn_major, "lvt_gemm_bf16: SOFTMAX mode needs K-major Q and K");
    LVT_CHECK_ARG(g->o_cin >= g->N && g->o_ld % 8 == 0, "lvt_gemm_bf16: SOFTMAX output must be plain rows");
    bn = 256;
  } else if (g->mode == LVT_EPI_DS) {
    LVT_CHECK_ARG(g->out_bf16 && g->aux_bf16 && g->delta, "lvt_gemm_bf16: DS mode needs out_bf16, aux_bf16 (P) and delta");
    if (g->v)
      LVT_CHECK_ARG(g->v_cin > 0 && g->v_zdiv > 0 && g->o2_cin >= 64 && g->o2_cin % 64 == 0 && g->o2_zdiv > 0 &&
                        g->a_cin % 64 == 0 && g->b_cin % 64 == 0 && g->v_cin % 64 == 0,
                    "lvt_gemm_bf16: fused dS/dQ needs 64-aligned operand blockings");
    ek = EK_DS;
    bn = (g->N % 256 == 0) ? 256 : 128;
  } else {
    LVT_CHECK_ARG(g->mode == LVT_EPI_LINEAR, "lvt_gemm_bf16: unknown epilogue mode %d", g->mode);
    bn = (g->N % 256 == 0) ? 256 : 128;
    // latency-bound small problems (a few output tiles, e.g. the M = 256 GEMMs of the sampler): 128-wide tiles
    // give twice the CTAs and six instead of four k-blocks in flight per CTA
    if (bn == 256 && g->N % 128 == 0 && !auto_split &&
        (long long)((g->M + BM - 1) / BM) * (g->N / 256) * g->batch * (g->splits > 0 ? g->splits : 1) <= 37)
      bn = 128;
  }

  // CTA pairs on 256 x 256 tiles when the problem has at least one full pair tile per pair of SMs' worth of
  // rows (M >= 256) -- every linear-epilogue shape of the train step; the small M = 256 sampler GEMMs went to
  // 128-wide single-CTA tiles above
  const int cg = (ek == EK_LINEAR && bn == 256 && g->M >= 2 * BM && pairs_enabled()) ? 2 : 1;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = g->M; p.N = g->N; p.K = g->K; p.batch = g->batch;
  const int kb_total = (g->K + BK - 1) / BK;
  if (auto_split) {
    // never more tiles than CTAs (pairs): the kernel is persistent, one tile too many doubles the time of one SM
    const long long tiles = (long long)((g->M + BM * cg - 1) / (BM * cg)) * ((g->N + bn - 1) / bn) * g->batch;
    const long long slots = num_sms() / cg;
    long long sp = slots / (tiles > 0 ? tiles : 1);
    if (sp > g->K / 256) sp = g->K / 256;
    splits = (int)(sp < 1 ? 1 : sp);
  }
  if (splits > kb_total) splits = kb_total;
  p.kb_per = (kb_total + splits - 1) / splits;
  splits = (kb_total + p.kb_per - 1) / p.kb_per;  // no empty split
  p.splits = splits;
  p.tiles_m = (g->M + BM * cg - 1) / (BM * cg);
  p.tiles_n = (g->N + bn - 1) / bn;
  const long long total = (long long)p.tiles_m * p.tiles_n * splits * g->batch;
  LVT_CHECK_ARG(total < (1ll << 30), "lvt_gemm_bf16: too many tiles");
  p.total_tiles = (int)total;
  p.a_cin = g->a_cin; p.a_zdiv = g->a_zdiv; p.b_cin = g->b_cin; p.b_zdiv = g->b_zdiv;
  p.flags = g->flags; p.alpha = g->alpha;
  p.out_f32 = g->out_f32; p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(g->out_bf16);
  p.res = g->res; p.aux = reinterpret_cast<const __nv_bfloat16*>(g->aux_bf16);
  p.bias = g->bias; p.bias_mod = g->bias_mod;
  p.o_cin = g->o_cin; p.o_zdiv = g->o_zdiv;
  p.o_shift = -1;
  if (g->o_cin < g->N) {
    int sh = 0;
    while ((1 << sh) < g->o_cin) ++sh;
    LVT_CHECK_ARG((1 << sh) == g->o_cin, "lvt_gemm_bf16: a blocked output needs a power-of-two o_cin (got %d)", g->o_cin);
    p.o_shift = sh;
  }
  LVT_CHECK_ARG(g->o_ld < (1ll << 23), "lvt_gemm_bf16: o_ld too large");
  p.o_ld = g->o_ld; p.o_s_blk = g->o_s_blk; p.o_s_zlo = g->o_s_zlo; p.o_s_zhi = g->o_s_zhi;
  p.lse = g->lse; p.delta = g->delta;
  p.bank_t = g->bank_t; p.bank_h = g->bank_h; p.bank_w = g->bank_w;
  p.heads = g->heads > 0 ? g->heads : 1;
  p.rowdot = g->rowdot; p.rd_block = g->rd_block; p.rd_L = g->rd_L;
  p.prof = reinterpret_cast<long long*>(g->prof);

  Maps m;
  memset(&m, 0, sizeof(m));
  int rc;
  // contiguous / strided extents per major-ness
  if (g->a_conv || g->b_conv) {
    LVT_CHECK_ARG(g->batch == 1 && g->cv_ntaps > 0 && g->cv_ntaps <= 16 && g->cv_C % 64 == 0 && g->cv_W > 0 &&
                      128 % g->cv_W == 0 && (g->cv_H * g->cv_W) % 128 == 0 && g->cv_N > 0,
                  "lvt_gemm_bf16: conv mode needs batch 1, <= 16 taps, C %% 64 == 0, W | 128, H*W %% 128 == 0");
    LVT_CHECK_ARG(!(g->a_conv && g->b_conv), "lvt_gemm_bf16: only one conv operand");
    p.a_conv = g->a_conv; p.b_conv = g->b_conv; p.cv_C = g->cv_C; p.cv_W = g->cv_W; p.cv_H = g->cv_H;
    for (int i = 0; i < g->cv_ntaps; ++i) {
      LVT_CHECK_ARG(g->cv_dh[i] >= -8 && g->cv_dh[i] < 8 && g->cv_dw[i] >= -8 && g->cv_dw[i] < 8 && g->cv_ph[i] >= 0 &&
                        g->cv_ph[i] < 8, "lvt_gemm_bf16: conv tap offsets must lie in [-8, 8), phases in [0, 8)");
      p.cv_dh_pk |= (unsigned long long)(g->cv_dh[i] + 8) << (4 * i);
      p.cv_dw_pk |= (unsigned long long)(g->cv_dw[i] + 8) << (4 * i);
      p.cv_ph_pk |= (unsigned long long)(g->cv_ph[i] + 8) << (4 * i);
    }
    const long long pix = g->cv_pix_stride > 0 ? g->cv_pix_stride : g->cv_C;
    if (g->a_conv) {
      LVT_CHECK_ARG(!g->a_mn_major && g->K == g->cv_ntaps * g->cv_C && g->M == g->cv_N * g->cv_H * g->cv_W,
                    "lvt_gemm_bf16: conv A needs K == ntaps*C and M == N*H*W");
      rc = make_conv_map(&m.a, g->a, g->cv_C, g->cv_W, g->cv_H, g->cv_N, g->cv_P, pix, g->cv_s_phase, BM / g->cv_W);
    } else {
      LVT_CHECK_ARG(g->b_mn_major && g->N == g->cv_ntaps * g->cv_C && g->K == g->cv_N * g->cv_H * g->cv_W,
                    "lvt_gemm_bf16: conv B needs MN-major, N == ntaps*C, K == N*H*W");
      rc = make_conv_map(&m.b, g->b, g->cv_C, g->cv_W, g->cv_H, g->cv_N, g->cv_P, pix, g->cv_s_phase, BK / g->cv_W);
    }
    if (rc) return rc;
  }
  if (!g->a_conv) {
    rc = g->a_mn_major
             ? make_operand_map(&m.a, g->a, g->M, g->K, g->a_cin, g->a_ld, g->a_s_blk, g->batch, g->a_zdiv, g->a_s_zlo, g->a_s_zhi, BK)
             : make_operand_map(&m.a, g->a, g->K, g->M, g->a_cin, g->a_ld, g->a_s_blk, g->batch, g->a_zdiv, g->a_s_zlo, g->a_s_zhi, BM);
    if (rc) return rc;
  }
  if (!g->b_conv) {
    rc = g->b_mn_major
             ? make_operand_map(&m.b, g->b, g->N, g->K, g->b_cin, g->b_ld, g->b_s_blk, g->batch, g->b_zdiv, g->b_s_zlo, g->b_s_zhi, BK)
             : make_operand_map(&m.b, g->b, g->K, g->N, g->b_cin, g->b_ld, g->b_s_blk, g->batch, g->b_zdiv, g->b_s_zlo, g->b_s_zhi, bn / cg);
    if (rc) return rc;
  }

  // TMA epilogue when the output is ONE tensor of plain 128 B-aligned rows; an optional second tensor with
  // the same addressing (fp32 residual / bf16 mask source / bf16 P) is TMA-loaded per slab.
  int st = 0;
  const bool debug_flags = (g->flags & (3 << 29)) != 0;
  const bool single_out = (g->out_f32 == nullptr) != (g->out_bf16 == nullptr);
  if ((ek == EK_LINEAR || ek == EK_DS) && single_out && g->bias_mod == 0 && !debug_flags) {
    const int esize = g->out_bf16 ? 2 : 4;
    const int slab_cols = 128 / esize;
    const void* obase = g->out_bf16 ? g->out_bf16 : (void*)g->out_f32;
    const bool atomic = (g->flags & LVT_GEMM_ATOMIC) != 0,
               mask = (g->flags & (LVT_GEMM_MASK | LVT_GEMM_AUX_ADD | LVT_GEMM_ROWDOT)) != 0;
    const void* cbase = nullptr;
    int want = -1;
    if (ek == EK_DS) { want = ST_TMA | ST_CLOAD; cbase = g->aux_bf16; }
    else if (atomic) { if (!g->out_bf16 && !mask && !g->res) want = ST_TMA | ST_F32 | ST_REDUCE; }
    else if (g->res && !mask) { if (g->out_f32) { want = ST_TMA | ST_F32 | ST_CLOAD; cbase = g->res; } }
    else if (mask && !g->res) { if (g->out_bf16) { want = ST_TMA | ST_CLOAD; cbase = g->aux_bf16; } }
    else if (!g->res && !mask) want = g->out_bf16 ? ST_TMA : (ST_TMA | ST_F32);
    const bool aligned = g->N % slab_cols == 0 && g->o_cin % slab_cols == 0 && (g->o_ld * esize) % 16 == 0 &&
                         (g->o_s_blk * esize) % 16 == 0 && (g->o_s_zlo * esize) % 16 == 0 &&
                         (g->o_s_zhi * esize) % 16 == 0 && (reinterpret_cast<uintptr_t>(obase) & 15) == 0 &&
                         (!cbase || (reinterpret_cast<uintptr_t>(cbase) & 15) == 0) &&
                         (!g->bias || (reinterpret_cast<uintptr_t>(g->bias) & 15) == 0);
    if (want >= 0 && aligned) {
      rc = make_operand_map(&m.o, obase, g->N, g->M, g->o_cin, g->o_ld, g->o_s_blk, g->batch, g->o_zdiv,
                            g->o_s_zlo, g->o_s_zhi, 32, esize);
      if (rc) return rc;
      if (cbase) {
        rc = make_operand_map(&m.c, cbase, g->N, g->M, g->o_cin, g->o_ld, g->o_s_blk, g->batch, g->o_zdiv,
                              g->o_s_zlo, g->o_s_zhi, 32, esize);
        if (rc) return rc;
      }
      st = want;
    }
  }

  LVT_CHECK_ARG(!(g->flags & LVT_GEMM_ROWDOT) || st == (ST_TMA | ST_CLOAD),
                "lvt_gemm_bf16: ROWDOT needs 128-byte aligned bf16 output / aux rows (TMA epilogue)");
  const int grid = cg == 2 ? 2 * (p.total_tiles < num_sms() / 2 ? p.total_tiles : num_sms() / 2)
                           : (p.total_tiles < num_sms() ? p.total_tiles : num_sms());
  const bool amn = g->a_mn_major != 0, bmn = g->b_mn_major != 0;
  if ((ek == EK_SOFTMAX_1x16x16 || ek == EK_SOFTMAX_4x8x8) && g->out_bf16) {
    rc = make_operand_map(&m.o, g->out_bf16, g->N, g->M, g->o_cin, g->o_ld, g->o_s_blk, g->batch, g->o_zdiv,
                          g->o_s_zlo, g->o_s_zhi, 32, 2);
    if (rc) return rc;
  }
  if (g->v && g->mode != LVT_EPI_LINEAR) {
    // fused attention kernels: the second operand (V forward / K backward) as an MN-major B operand (64-key
    // boxes), the second output (O / dQ) through its own store map
    if (ek == EK_DS)
      LVT_CHECK_ARG(st == (ST_TMA | ST_CLOAD) && g->N == 256 && g->M == 256 && g->K == 128 && g->o2_bf16 && g->o2_n == 128 &&
                        !amn && !bmn,
                    "lvt_gemm_bf16: fused dS/dQ needs M == N == 256, K == 128, K-major dO and V, aligned P / dS rows");
    CUtensorMap tm_v, tm_o2;
    rc = make_operand_map(&tm_v, g->v, g->o2_n, g->N, g->v_cin, g->v_ld, 0, g->batch, g->v_zdiv, g->v_s_zlo, g->v_s_zhi, BK);
    if (rc) return rc;
    rc = make_operand_map(&tm_o2, g->o2_bf16, g->o2_n, g->M, g->o2_cin, g->o2_ld, 0, g->batch, g->o2_zdiv, g->o2_s_zlo,
                          g->o2_s_zhi, 32, 2);
    if (rc) return rc;
    if (!g->out_bf16) m.o = tm_o2;  // unused placeholder (store_p == 0)
    if (ek != EK_DS) m.c = m.o;     // unused placeholder
    // EK_DS with bank pointers: they are the dt/dh/dw_bank GRADIENTS, accumulated in the epilogue (block (1,16,16))
    const int bank_fused = (ek == EK_DS && g->bank_t && g->bank_h && g->bank_w) ? 1 : 0;
    if (bank_fused)
      LVT_CHECK_ARG(g->bt == 1 && g->bh == 16 && g->bw == 16 && g->heads > 0,
                    "lvt_gemm_bf16: the fused bank gradient covers attention block (1,16,16) only");
    auto launch = [&](auto kern, int which) -> int {
      static bool configured[3] = {false, false, false};
      if (!configured[which]) {
        LVT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::TOTAL));
        configured[which] = true;
      }
      LVT_CHECK_CUDA(lvt_launch(kern, dim3(grid), dim3(NUM_THREADS), AttnSmem::TOTAL, stream, m.a, m.b, tm_v, m.o, tm_o2,
                                m.c, p, g->v_cin, g->v_zdiv, g->o2_cin, g->o2_zdiv, g->out_bf16 ? 1 : 0, bank_fused));
      lvt_count_launch(1);
      return LVT_OK;
    };
    if (ek == EK_SOFTMAX_1x16x16) return launch(attn_fwd_kernel<EK_SOFTMAX_1x16x16>, 0);
    if (ek == EK_SOFTMAX_4x8x8) return launch(attn_fwd_kernel<EK_SOFTMAX_4x8x8>, 1);
    return launch(attn_fwd_kernel<EK_DS>, 2);
  }
  if (ek == EK_SOFTMAX_1x16x16) return launch_gemm<256, false, false, EK_SOFTMAX_1x16x16>(m, p, grid, stream);
  if (ek == EK_SOFTMAX_4x8x8) return launch_gemm<256, false, false, EK_SOFTMAX_4x8x8>(m, p, grid, stream);
  if (ek == EK_DS) {
    LVT_CHECK_ARG(!amn && !bmn, "lvt_gemm_bf16: DS mode needs K-major dO and V");
    if (st != 0) {
      if (bn == 256) return launch_gemm<256, false, false, EK_DS, ST_TMA | ST_CLOAD>(m, p, grid, stream);
      return launch_gemm<128, false, false, EK_DS, ST_TMA | ST_CLOAD>(m, p, grid, stream);
    }
    if (bn == 256) return launch_gemm<256, false, false, EK_DS>(m, p, grid, stream);
    return launch_gemm<128, false, false, EK_DS>(m, p, grid, stream);
  }
  if (cg == 2) return dispatch_store<256, 2>(m, p, grid, amn, bmn, st, stream);
  if (bn == 256) return dispatch_store<256>(m, p, grid, amn, bmn, st, stream);
  return dispatch_store<128>(m, p, grid, amn, bmn, st, stream);
}
