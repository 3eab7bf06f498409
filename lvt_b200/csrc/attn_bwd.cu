// lvt_b200 :: fused attention BACKWARD for one 256-token block per (sequence, head): no P / dS in HBM.
//
// Reference semantics: autograd of ScaledDotProductAttention.forward (vt_attention.py:61-81) with the
// relative-position bias of BlockLocalAttention.get_B (vt_attention.py:169-174):
//   S  = scale * Q K^T + B  [causal: keys after the query -> -1e4]      P  = softmax_row(S)
//   dP = dO V^T             dS = P * (dP - delta),  delta = rowsum(dO * O)
//   dV = P^T dO             dK = scale * dS^T Q     dQ = scale * dS K     dB = dS (-> dt/dh/dw_bank)
// The forward keeps only the row log-sum-exp; this kernel recomputes S and P = exp(S - lse) on the tensor
// cores, so the only HBM traffic is Q, K, V, dO in and dQ, dK, dV out.
//
// One CTA per (sequence, head) z, persistent over z.  The 256 x 256 score matrix is processed as 128 x 128
// blocks (query half i, key half c) in the order (0,0) (1,0) | (0,1) (1,1)  [causal: (0,0) (1,0) | (1,1), the
// block above the diagonal is exactly zero]:
//   warp 0    TMA producer: K_c, V_c when c changes, Q_i, dO_i per block (single-buffered, 32 KiB each)
//   warp 1    MMA issuer:   S = Q_i K_c^T, dP = dO_i V_c^T                      (tensor memory columns 0-127, 128-255)
//                           dV_c += P^T dO_i, dK_c += dS^T Q_i, dQ_i(c) = dS K_c (columns 256-383, 384-511, 0-127)
//   warps 2-9 epilogue:     S, dP -> P, dS (bf16, 128B-swizzled slabs in shared memory that serve as K-major AND
//                           MN-major A operands), bank gradients from the fp32 dS; accumulator drains
// The shared-memory operand buffers are laid out [64-column half][128 rows][128 B]; the same bytes are read as a
// K-major operand (S, dP: contraction over da) and as an MN-major operand (dV, dK, dQ: contraction over the rows)
// by changing only the UMMA descriptor (LBO = 16 KiB between the two 64-wide atoms).
// dV_c / dK_c accumulate over both query halves in tensor memory; dQ_i needs both key halves, which do not fit
// next to them (6 x 128 columns), so the c = 0 partial of dQ_i is parked in a per-CTA fp32 scratch (128 KiB,
// written and re-read by the same thread, L2-resident) and added when c = 1 completes it.
#include <string.h>

#include "../../include/lvt_b200.h"
#include "common.cuh"

extern void lvt_count_launch(int n);
extern int lvt_sm_limit();
int lvt_make_operand_map(CUtensorMap* out, const void* base, long long c_extent, long long r_extent, int cin,
                         long long ld, long long s_blk, int batch, int zdiv, long long s_zlo, long long s_zhi,
                         int box_rows, int esize);

namespace {

constexpr int NUM_THREADS = 320;
constexpr int NUM_EPI_WARPS = 8;
constexpr int L = 256;    // positions per attention block
constexpr int DA = 128;   // head dimension

struct Smem {
  static constexpr int Q_OFF = 0;        // 2 x [2 da-halves][128 queries][128 B]: Q of the next block lands while this one runs
  static constexpr int DO_OFF = 65536;   // [2 da-halves][128 queries][128 B]
  static constexpr int K_OFF = 98304;    // [2 da-halves][128 keys][128 B]
  static constexpr int V_OFF = 131072;
  static constexpr int P_OFF = 163840;   // [2 key-halves][128 queries][128 B]; also the store staging of dQ / dK
  static constexpr int DS_OFF = 196608;  // same for dS; staging of dV
  static constexpr int BAR_OFF = 229376;
  static constexpr int BINS_OFF = BAR_OFF + 256;  // [8 warps][64] fp32 bank-gradient sums
  static constexpr int TOTAL = BINS_OFF + 2048;
};
static_assert(Smem::TOTAL <= 232448, "attention backward: shared memory budget");

struct Params {
  int nz, heads, causal;
  float scale;
  const float* lse;
  const float* delta;
  const float* bank_t;
  const float* bank_h;
  const float* bank_w;
  float* dbank_t;
  float* dbank_h;
  float* dbank_w;
  float* scratch;
  long long* prof;  // optional timeline of CTA 0 (tools/attn_bwd_prof.py): 16 clock64 stamps per block
};

struct Blk {
  int i, c;
  bool first_c, last_c, dq_in, dq_final, new_q, new_kv;
};
// Block order inside one z: (0,0) (1,0) (1,1) (0,1); causal stops after three (the block above the diagonal is zero).
// first_c / last_c: first / last contribution to dV_c, dK_c (accumulate flag / drain); dq_in: a partial of dQ_i
// waits in the scratch; dq_final: this block completes dQ_i; new_q: Q_i and dO_i differ from the previous block's;
// new_kv: K_c and V_c do.
LVT_DEVICE_INLINE Blk blk_info(bool causal, int b) {
  Blk k;
  k.i = (b == 1 || b == 2) ? 1 : 0;
  k.c = b >> 1;
  k.first_c = (b & 1) == 0;
  k.last_c = (b & 1) != 0 || (causal && b == 2);
  k.new_q = b != 2;
  k.new_kv = (b & 1) == 0;
  if (!causal) {
    k.dq_in = b >= 2; k.dq_final = b >= 2;
  } else {
    k.dq_in = b == 2; k.dq_final = b != 1;
  }
  return k;
}

#define PROF(slot)                                                                              \
  do {                                                                                          \
    if (p.prof && blockIdx.x == 0 && n < 64) p.prof[n * 16 + (slot)] = clock64();               \
  } while (0)

LVT_DEVICE_INLINE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int BT, int BH, int BW>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_dk,
                const __grid_constant__ CUtensorMap tm_dv, const Params p) {
  static_assert(BT * BH * BW == L && 64 % BW == 0 && (BH * BW) % 64 == 0, "attention block shape");
  extern __shared__ __align__(1024) uint8_t smem[];
  using S = Smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* q_full = bars + 0;     // [2] Q_i landed in buffer 0 / 1
  uint64_t* do_full = bars + 2;    // dO_i landed
  uint64_t* k_full = bars + 3;     // K_c landed
  uint64_t* v_full = bars + 4;     // V_c landed
  uint64_t* s_full = bars + 5;     // S, dP complete in tensor memory (V_c no longer needed)
  uint64_t* ps_full = bars + 6;    // P, dS slabs written (8 warps)
  uint64_t* dv_done = bars + 7;    // dV MMAs complete: dO_i and the P slab are free
  uint64_t* dq_done = bars + 8;    // ... and dQ: K_c free, dQ readable
  uint64_t* mma2_done = bars + 9;  // ... and dK: Q_i and the dS slab free, dV / dK readable
  uint64_t* s_free = bars + 10;    // the epilogue has read dQ (columns 0-127): the next S may be issued (8 warps)
  uint64_t* acc_free = bars + 11;  // the epilogue has drained dV_c / dK_c: the next c may overwrite them (8 warps)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);
  float* const bank_bins = reinterpret_cast<float*>(smem + S::BINS_OFF);  // [8 warps][64]
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool causal = p.causal != 0;
  const int nblk = causal ? 3 : 4;
  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("lvt_b200: attention backward needs 1024-byte aligned dynamic shared memory\n");
      __trap();
    }
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    for (int i = 0; i < 6; ++i) mbar_init(bars + i, 1);  // q_full[2], do_full, k_full, v_full, s_full
    mbar_init(ps_full, NUM_EPI_WARPS);
    mbar_init(dv_done, 1);
    mbar_init(dq_done, 1);
    mbar_init(mma2_done, 1);
    mbar_init(s_free, NUM_EPI_WARPS);
    mbar_init(acc_free, NUM_EPI_WARPS);
    fence_barrier_init();
  }
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
    // Every operand is requested as soon as its buffer is free: V_c after the first MMA group of the previous
    // block, dO_i after its dV MMAs, K_c after its dQ MMAs, and Q_i -- double-buffered -- a whole block ahead.
    if (elect_one()) {
      auto load_q = [&](int buf, int z, int i) {
        mbar_arrive_expect_tx(&q_full[buf], 32768);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
          tma_load_5d(smem + S::Q_OFF + buf * 32768 + kb * 16384, &tm_q, &q_full[buf], kb * 64, i * 128, 0, z % p.heads,
                      z / p.heads);
      };
      uint32_t n = 0;
      int qs = 0;
      if ((int)blockIdx.x < p.nz) load_q(1, blockIdx.x, 0);  // the first block toggles to buffer 1
      for (int z = blockIdx.x; z < p.nz; z += gridDim.x) {
        const int zlo = z % p.heads, zhi = z / p.heads;
        const int zn = z + gridDim.x;  // this CTA's next z: pulled into L2 while the current one is processed
        const int znlo = zn % p.heads, znhi = zn / p.heads;
        for (int b = 0; b < nblk; ++b, ++n) {
          const Blk k = blk_info(causal, b);
          if (k.new_q) qs ^= 1;
          const uint32_t prev = (n & 1) ^ 1;  // parity of block n-1's phase (passes at once for n == 0)
          mbar_wait(s_full, prev);
          PROF(0);
          if (k.new_kv) {
            mbar_arrive_expect_tx(v_full, 32768);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
              tma_load_5d(smem + S::V_OFF + kb * 16384, &tm_v, v_full, kb * 64, k.c * 128, 0, zlo, zhi);
          }
          mbar_wait(dv_done, prev);
          if (k.new_q) {
            mbar_arrive_expect_tx(do_full, 32768);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
              tma_load_5d(smem + S::DO_OFF + kb * 16384, &tm_do, do_full, kb * 64, k.i * 128, 0, zlo, zhi);
          }
          mbar_wait(dq_done, prev);
          if (k.new_kv) {
            mbar_arrive_expect_tx(k_full, 32768);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
              tma_load_5d(smem + S::K_OFF + kb * 16384, &tm_k, k_full, kb * 64, k.c * 128, 0, zlo, zhi);
          }
          mbar_wait(mma2_done, prev);
          // Q of the NEXT block into the buffer this block does not use (last read by block n-1 at the latest)
          {
            const int b2 = b + 1 < nblk ? b + 1 : 0;
            const int z2 = b + 1 < nblk ? z : zn;
            if (z2 < p.nz) {
              const Blk k2 = blk_info(causal, b2);
              if (k2.new_q) load_q(qs ^ 1, z2, k2.i);
            }
          }
          if (zn < p.nz && b < 2) {  // block 0: rows 0-127 of the next z's Q, K, V, dO; block 1: rows 128-255
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              tma_prefetch_5d(&tm_q, kb * 64, b * 128, 0, znlo, znhi);
              tma_prefetch_5d(&tm_k, kb * 64, b * 128, 0, znlo, znhi);
              tma_prefetch_5d(&tm_v, kb * 64, b * 128, 0, znlo, znhi);
              tma_prefetch_5d(&tm_do, kb * 64, b * 128, 0, znlo, znhi);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t id_kk = umma_idesc(128, 128, /*bf16*/ 1, false, false);
      constexpr uint32_t id_mm = umma_idesc(128, 128, 1, true, true);
      constexpr uint32_t id_km = umma_idesc(128, 128, 1, false, true);
      const uint32_t do_base = smem_u32(smem + S::DO_OFF);
      const uint32_t k_base = smem_u32(smem + S::K_OFF), v_base = smem_u32(smem + S::V_OFF);
      const uint32_t p_base = smem_u32(smem + S::P_OFF), ds_base = smem_u32(smem + S::DS_OFF);
      // K-major view: 64 contraction elements per 128 B row; MN-major view of the SAME bytes: the rows are the
      // contraction index, the two 64-wide column halves are the M/N atoms, 16 KiB apart.  The descriptors are
      // built once; per MMA only the start-address field moves (a constant added to the low word), so the
      // single issuing thread spends a few instructions per tcgen05.mma instead of rebuilding 64-bit descriptors.
      const uint64_t kq0 = umma_smem_desc(smem_u32(smem + S::Q_OFF), 16, 1024), mq0 = umma_smem_desc(smem_u32(smem + S::Q_OFF), 16384, 1024);
      const uint64_t kdo = umma_smem_desc(do_base, 16, 1024), mdo = umma_smem_desc(do_base, 16384, 1024);
      const uint64_t kk = umma_smem_desc(k_base, 16, 1024), mk = umma_smem_desc(k_base, 16384, 1024);
      const uint64_t kv = umma_smem_desc(v_base, 16, 1024);
      const uint64_t mp = umma_smem_desc(p_base, 16384, 1024);
      const uint64_t kds = umma_smem_desc(ds_base, 16, 1024), mds = umma_smem_desc(ds_base, 16384, 1024);
      // byte offset of k-step (kb, k4) in the K-major view / of k-step k16 in the MN-major view, in 16 B units
      auto ko = [](int kb, int k4) { return (uint64_t)((kb * 16384 + k4 * 32) >> 4); };
      auto mo = [](int k16) { return (uint64_t)((k16 * 2048) >> 4); };
      uint32_t n = 0, nkv = 0, ndo = 0, nq0 = 0, nq1 = 0;
      int qs = 0;
      for (int z = blockIdx.x; z < p.nz; z += gridDim.x) {
        if (z + (int)gridDim.x >= p.nz) pdl_launch_dependents();
        for (int b = 0; b < nblk; ++b, ++n) {
          const Blk k = blk_info(causal, b);
          if (k.new_q) {
            qs ^= 1;
            if (qs) { mbar_wait(&q_full[1], nq1 & 1); ++nq1; }
            else { mbar_wait(&q_full[0], nq0 & 1); ++nq0; }
          }
          const uint64_t kq = kq0 + (uint64_t)(qs * (32768 >> 4)), mq = mq0 + (uint64_t)(qs * (32768 >> 4));
          if (k.new_kv) mbar_wait(k_full, nkv & 1);
          PROF(1);
          mbar_wait(s_free, (n & 1) ^ 1);  // dQ of the previous block (columns 0-127) has been read
          tc_fence_after();
          PROF(2);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16_ss(tmem_base, kq + ko(kb, k4), kk + ko(kb, k4), id_kk, (kb | k4) ? 1u : 0u);
          if (k.new_q) {
            mbar_wait(do_full, ndo & 1);
            ++ndo;
          }
          if (k.new_kv) mbar_wait(v_full, nkv & 1);
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16_ss(tmem_base + 128, kdo + ko(kb, k4), kv + ko(kb, k4), id_kk, (kb | k4) ? 1u : 0u);
          umma_commit(s_full);
          if (k.first_c) {  // dV_c / dK_c of the previous c have been drained (overlapped the two products above)
            mbar_wait(acc_free, (nkv & 1) ^ 1);
            ++nkv;
          }
          mbar_wait(ps_full, n & 1);  // P, dS are in shared memory; S and dP have been read
          tc_fence_after();
          PROF(3);
          const uint32_t acc0 = k.first_c ? 0u : 1u;
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16)  // dV_c[key, :] += sum_q P[q, key] dO[q, :]
            umma_bf16_ss(tmem_base + 256, mp + mo(k16), mdo + mo(k16), id_mm, k16 ? 1u : acc0);
          umma_commit(dv_done);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)     // dQ_i[q, :] = sum_key dS[q, key] K[key, :]
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16_ss(tmem_base, kds + ko(kb, k4), mk + mo(kb * 4 + k4), id_km, (kb | k4) ? 1u : 0u);
          umma_commit(dq_done);  // (in-order completion: dV no longer reads the P slab, which stages the dQ store)
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16)  // dK_c[key, :] += sum_q dS[q, key] Q[q, :]
            umma_bf16_ss(tmem_base + 384, mds + mo(k16), mq + mo(k16), id_mm, k16 ? 1u : acc0);
          umma_commit(mma2_done);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    constexpr int NY = 64 / BW;                  // key rows (of BW keys) inside this warp's 64 keys
    constexpr int NG = 32 / BW;                  // query rows (of BW queries) inside this warp's 32 queries
    constexpr int NBH = 2 * BH - 1, NBW = 2 * BW - 1, NBT = 2 * BT - 1;
    static_assert(NBH <= 32 && NBW <= 32 && NBT <= 32, "one bank bin per lane");
    const int ew = warp - 2;
    const int q = warp & 3;       // tensor-memory lane quarter: rows 32q .. 32q+31 of the block
    const int half = ew >> 2;     // which 64 keys (P, dS) / which 64 output columns (drains)
    const int row_l = q * 32 + lane;
    const float kLog2e = 1.4426950408889634f;
    const float a2 = p.scale * kLog2e;
    uint4* const pslab = reinterpret_cast<uint4*>(smem + S::P_OFF + half * 16384 + q * 4096);
    uint4* const dslab = reinterpret_cast<uint4*>(smem + S::DS_OFF + half * 16384 + q * 4096);
    const uint32_t t_own = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 64;
    // dQ partial sums: [query half][warp][16 column quads][32 lanes] float4 -> every access is one 512 B row per warp
    float4* const scratch = reinterpret_cast<float4*>(p.scratch) + (size_t)blockIdx.x * 8192 + ew * 512 + lane;
    auto stage32 = [&](uint4* slab, int ch, const float (&v)[32]) {  // 32 columns of this lane's row -> bf16 units
#pragma unroll
      for (int k8 = 0; k8 < 4; ++k8) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * k8], v[8 * k8 + 1]);
        u.y = pack_bf16x2(v[8 * k8 + 2], v[8 * k8 + 3]);
        u.z = pack_bf16x2(v[8 * k8 + 4], v[8 * k8 + 5]);
        u.w = pack_bf16x2(v[8 * k8 + 6], v[8 * k8 + 7]);
        slab[lane * 8 + ((4 * ch + k8) ^ (lane & 7))] = u;
      }
    };
    uint32_t n = 0;
    for (int z = blockIdx.x; z < p.nz; z += gridDim.x) {
      const int zlo = z % p.heads, zhi = z / p.heads;
      const int head = zlo;
      // bank gradients: lane b of every warp owns bin b of dh_bank, dw_bank and dt_bank for the whole z
      float acc_h = 0.f, acc_w = 0.f, acc_t = 0.f;
      for (int b = 0; b < nblk; ++b, ++n) {
        const Blk k = blk_info(causal, b);
        const int qi = k.i * 128 + row_l;  // query position inside the block of 256
        const int ti = qi / (BH * BW), hi = (qi / BW) % BH, wi = qi % BW;
        const int key0 = k.c * 128 + half * 64;
        const int tj = key0 / (BH * BW), hj0 = (key0 / BW) % BH;
        const float lse2 = p.lse[(size_t)z * L + qi] * kLog2e;
        const float dl = p.delta[(size_t)z * L + qi];
        // bias slices of this row for its 64 keys, log2 domain, minus the row's log-sum-exp
        // (get_B: B[i, j] = bt[ti-tj] + bh[hi-hj] + bw[wi-wj])
        float bwv[BW], bhv[NY];
        const float btv = kLog2e * __ldg(p.bank_t + head * NBT + (ti - tj + BT - 1)) - lse2;
#pragma unroll
        for (int y = 0; y < NY; ++y) bhv[y] = btv + kLog2e * __ldg(p.bank_h + head * NBH + (hi - (hj0 + y) + BH - 1));
#pragma unroll
        for (int x = 0; x < BW; ++x) bwv[x] = kLog2e * __ldg(p.bank_w + head * NBW + (wi - x + BW - 1));
        const bool diag = causal && k.i == k.c;
        float rs_h[NY], cs_w[BW];  // bank-gradient partial sums of this row: per key row / per key column
#pragma unroll
        for (int y = 0; y < NY; ++y) rs_h[y] = 0.f;
#pragma unroll
        for (int x = 0; x < BW; ++x) cs_w[x] = 0.f;

        mbar_wait(s_full, n & 1);
        tc_fence_after();
        if (ew == 0 && lane == 0) PROF(4);
        if (lane == 0) bulk_wait_group_read<0>();  // this warp's staged stores have drained its slab pieces
        __syncwarp();
        {
          // 16 keys at a time; the tensor-memory loads of the next 16 are in flight while these are processed
          uint32_t rsb[2][16], rpb[2][16];
          tmem_ld_32x16(t_own, rsb[0]);
          tmem_ld_32x16(t_own + 128, rpb[0]);
          tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            if (c4 < 3) {
              tmem_ld_32x16(t_own + 16 * (c4 + 1), rsb[(c4 + 1) & 1]);
              tmem_ld_32x16(t_own + 128 + 16 * (c4 + 1), rpb[(c4 + 1) & 1]);
            }
            const uint32_t(&rs)[16] = rsb[c4 & 1];
            const uint32_t(&rp)[16] = rpb[c4 & 1];
            uint32_t up[8], ud[8];
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const int kk = 16 * c4 + e;  // kk, kk + 1: the same key row (BW is even)
              float p0 = ex2(__uint_as_float(rs[e]) * a2 + (bhv[kk / BW] + bwv[kk % BW]));
              float p1 = ex2(__uint_as_float(rs[e + 1]) * a2 + (bhv[kk / BW] + bwv[kk % BW + 1]));
              if (diag) {  // keys after the query: the forward filled -1e4, exp underflows to exactly 0
                if (half * 64 + kk > row_l) p0 = 0.f;
                if (half * 64 + kk + 1 > row_l) p1 = 0.f;
              }
              // P rounded to bf16 BEFORE dS: delta = rowsum(dO * O) was formed from the bf16 P of the forward, and
              // dS = P * (dP - delta) only keeps its rows summing to zero (the cancellation that dominates peaked
              // rows) when both use the same P
              const uint32_t u = pack_bf16x2(p0, p1);
              up[e >> 1] = u;
              const float d0 = __uint_as_float(u << 16) * (__uint_as_float(rp[e]) - dl);
              const float d1 = __uint_as_float(u & 0xffff0000u) * (__uint_as_float(rp[e + 1]) - dl);
              ud[e >> 1] = pack_bf16x2(d0, d1);
              rs_h[kk / BW] += d0 + d1;
              cs_w[kk % BW] += d0;
              cs_w[kk % BW + 1] += d1;
            }
#pragma unroll
            for (int k8 = 0; k8 < 2; ++k8) {
              pslab[lane * 8 + ((2 * c4 + k8) ^ (lane & 7))] = make_uint4(up[4 * k8], up[4 * k8 + 1], up[4 * k8 + 2], up[4 * k8 + 3]);
              dslab[lane * 8 + ((2 * c4 + k8) ^ (lane & 7))] = make_uint4(ud[4 * k8], ud[4 * k8 + 1], ud[4 * k8 + 2], ud[4 * k8 + 3]);
            }
            if (c4 < 3) tmem_ld_wait();
          }
        }
        fence_proxy_async();  // generic-proxy writes -> visible to tcgen05.mma
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ps_full);
        if (ew == 0 && lane == 0) PROF(5);

        // fills the wait for the second MMA group: bank-gradient partial sums -> the lane that owns the bin
        // (shuffles only, no atomics)
        {
          const int hi0 = ((k.i * 128 + q * 32) / BW) % BH;  // hi of the warp's first query row; row g has hi0 + g
          float tot = 0.f;
#pragma unroll
          for (int y = 0; y < NY; ++y) {
            float v = rs_h[y];
            tot += v;
#pragma unroll
            for (int o = BW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);  // every lane of a query row
            const int g = lane - (hi0 - (hj0 + y) + BH - 1);                                // bin(row g) == lane
            const float w = __shfl_sync(0xffffffffu, v, (g & (NG - 1)) * BW);
            if (g >= 0 && g < NG) acc_h += w;
          }
#pragma unroll
          for (int x = 0; x < BW; ++x) {
            float v = cs_w[x];
#pragma unroll
            for (int o = 16; o >= BW; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);    // every lane with this wi
            const int src = lane - (BW - 1) + x;                                           // wi whose bin == lane
            const float w = __shfl_sync(0xffffffffu, v, src & (BW - 1));
            if (src >= 0 && src < BW) acc_w += w;
          }
          tot = warp_sum(tot);  // ti is uniform inside a warp (32 | BH*BW)
          if (lane == ti - tj + BT - 1) acc_t += tot;
        }
        if (ew == 0 && lane == 0) PROF(6);
        // the parked partial of dQ_i on its way to registers while the second MMA group finishes
        float4* const sc = scratch + (size_t)k.i * 4096;
        float4 part[16];
        if (k.dq_in) {
#pragma unroll
          for (int j = 0; j < 16; ++j) part[j] = sc[j * 32];
        }

        mbar_wait(dq_done, n & 1);
        tc_fence_after();
        if (ew == 0 && lane == 0) PROF(7);
        // dQ_i: scale, (+ the parked partial), then either park it or store it (staged in the P piece)
        uint32_t rq[2][32];
        tmem_ld_32x32(t_own, rq[0]);
        tmem_ld_32x32(t_own + 32, rq[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);  // dQ is in registers: the next block's S may overwrite it
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const uint32_t(&r)[32] = rq[ch];
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]) * p.scale;
          if (k.dq_in) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] += part[8 * ch + j].x; v[4 * j + 1] += part[8 * ch + j].y;
              v[4 * j + 2] += part[8 * ch + j].z; v[4 * j + 3] += part[8 * ch + j].w;
            }
          }
          if (k.dq_final) {
            stage32(pslab, ch, v);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) sc[(8 * ch + j) * 32] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        if (k.dq_final) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&tm_dq, pslab, half * 64, k.i * 128 + q * 32, 0, zlo, zhi);
            bulk_commit_group();
          }
        }
        if (ew == 0 && lane == 0) PROF(8);
        if (k.last_c) {
          mbar_wait(mma2_done, n & 1);
          tc_fence_after();
          // dV_c -> staged in this warp's dS piece
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t r[32];
            tmem_ld_32x32(t_own + 256 + 32 * ch, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
            stage32(dslab, ch, v);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&tm_dv, dslab, half * 64, k.c * 128 + q * 32, 0, zlo, zhi);
            bulk_commit_group();
            bulk_wait_group_read<1>();  // the dQ store (if any) has finished reading the P piece
          }
          __syncwarp();
          // dK_c -> staged in this warp's P piece
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t r[32];
            tmem_ld_32x32(t_own + 384 + 32 * ch, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]) * p.scale;
            stage32(pslab, ch, v);
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(acc_free);
            tma_store_5d(&tm_dk, pslab, half * 64, k.c * 128 + q * 32, 0, zlo, zhi);
            bulk_commit_group();
          }
        }
        if (ew == 0 && lane == 0) PROF(9);
      }
      // bank gradients of this z: every warp parks its lane-owned sums, the first epilogue warp adds the eight
      // rows and issues one atomic per bin
      asm volatile("bar.sync 6, 256;" ::: "memory");  // the previous z's sums have been read
      if (lane < NBH) bank_bins[ew * 64 + lane] = acc_h;
      if (lane < NBW) bank_bins[ew * 64 + NBH + lane] = acc_w;
      if (lane < NBT) bank_bins[ew * 64 + NBH + NBW + lane] = acc_t;
      asm volatile("bar.sync 5, 256;" ::: "memory");
      if (ew == 0) {
        for (int i = lane; i < NBH + NBW + NBT; i += 32) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < NUM_EPI_WARPS; ++w) v += bank_bins[w * 64 + i];
          float* dst = i < NBH ? p.dbank_h + head * NBH + i
                               : (i < NBH + NBW ? p.dbank_w + head * NBW + (i - NBH) : p.dbank_t + head * NBT + (i - NBH - NBW));
          atomicAdd(dst, v);
        }
      }
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

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

extern "C" long long lvt_attn_bwd_scratch_bytes(void) { return (long long)sm_count() * 32768 * 4; }

extern "C" int lvt_attn_bwd(const LvtAttnBwd* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LVT_CHECK_ARG(a != nullptr, "lvt_attn_bwd: null descriptor");
  LVT_CHECK_ARG(a->nseq > 0 && a->heads > 0, "lvt_attn_bwd: bad shape (nseq %d, heads %d)", a->nseq, a->heads);
  LVT_CHECK_ARG(a->qkv && a->dO && a->dqkv && a->lse && a->delta, "lvt_attn_bwd: null tensor");
  LVT_CHECK_ARG(a->bank_t && a->bank_h && a->bank_w && a->dbank_t && a->dbank_h && a->dbank_w,
                "lvt_attn_bwd: null bank / bank gradient");
  LVT_CHECK_ARG(a->scratch && a->scratch_bytes >= lvt_attn_bwd_scratch_bytes(),
                "lvt_attn_bwd: scratch must hold lvt_attn_bwd_scratch_bytes() = %lld bytes", lvt_attn_bwd_scratch_bytes());
  const int H = a->heads;
  LVT_CHECK_ARG(a->qkv_ld >= 3ll * H * DA && a->dqkv_ld >= 3ll * H * DA && a->do_ld >= (long long)H * DA,
                "lvt_attn_bwd: row strides too small for %d heads of %d", H, DA);
  const bool b1 = a->bt == 1 && a->bh == 16 && a->bw == 16, b4 = a->bt == 4 && a->bh == 8 && a->bw == 8;
  LVT_CHECK_ARG(b1 || b4, "lvt_attn_bwd: attention blocks (1,16,16) and (4,8,8) are supported, got (%d,%d,%d)", a->bt,
                a->bh, a->bw);
  const long long nz = (long long)a->nseq * H;
  LVT_CHECK_ARG(nz < (1ll << 30), "lvt_attn_bwd: too many (sequence, head) pairs");
  CUtensorMap mq, mk, mv, mdo, mdq, mdk, mdv;
  const uint16_t* qkv = reinterpret_cast<const uint16_t*>(a->qkv);
  uint16_t* dqkv = reinterpret_cast<uint16_t*>(a->dqkv);
  int rc;
  auto in_map = [&](CUtensorMap* m, const void* base, long long ld) {
    return lvt_make_operand_map(m, base, DA, L, DA, ld, 0, (int)nz, H, DA, (long long)L * ld, 128, 2);
  };
  auto out_map = [&](CUtensorMap* m, const void* base, long long ld) {
    return lvt_make_operand_map(m, base, DA, L, DA, ld, 0, (int)nz, H, DA, (long long)L * ld, 32, 2);
  };
  if ((rc = in_map(&mq, qkv, a->qkv_ld))) return rc;
  if ((rc = in_map(&mk, qkv + (size_t)H * DA, a->qkv_ld))) return rc;
  if ((rc = in_map(&mv, qkv + (size_t)2 * H * DA, a->qkv_ld))) return rc;
  if ((rc = in_map(&mdo, a->dO, a->do_ld))) return rc;
  if ((rc = out_map(&mdq, dqkv, a->dqkv_ld))) return rc;
  if ((rc = out_map(&mdk, dqkv + (size_t)H * DA, a->dqkv_ld))) return rc;
  if ((rc = out_map(&mdv, dqkv + (size_t)2 * H * DA, a->dqkv_ld))) return rc;
  Params p;
  memset(&p, 0, sizeof(p));
  p.nz = (int)nz; p.heads = H; p.causal = a->causal ? 1 : 0; p.scale = a->scale;
  p.lse = a->lse; p.delta = a->delta;
  p.bank_t = a->bank_t; p.bank_h = a->bank_h; p.bank_w = a->bank_w;
  p.dbank_t = a->dbank_t; p.dbank_h = a->dbank_h; p.dbank_w = a->dbank_w;
  p.scratch = a->scratch;
  p.prof = reinterpret_cast<long long*>(a->prof);
  const int lim = lvt_sm_limit();
  const int sms = (lim > 0 && lim < sm_count()) ? lim : sm_count();
  const int grid = (int)(nz < sms ? nz : sms);
  auto launch = [&](auto kern, int which) -> int {
    static bool configured[2] = {false, false};
    if (!configured[which]) {
      LVT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::TOTAL));
      configured[which] = true;
    }
    LVT_CHECK_CUDA(lvt_launch(kern, dim3(grid), dim3(NUM_THREADS), Smem::TOTAL, stream, mq, mk, mv, mdo, mdq, mdk, mdv, p));
    lvt_count_launch(1);
    return LVT_OK;
  };
  if (b1) return launch(attn_bwd_kernel<1, 16, 16>, 0);
  return launch(attn_bwd_kernel<4, 8, 8>, 1);
}
