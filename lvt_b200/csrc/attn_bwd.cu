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
int lvt_make_operand_map(CUtensorMap* out, const void* base, long long c_extent, long long r_extent, int cin,
                         long long ld, long long s_blk, int batch, int zdiv, long long s_zlo, long long s_zhi,
                         int box_rows, int esize);

namespace {

constexpr int NUM_THREADS = 320;
constexpr int NUM_EPI_WARPS = 8;
constexpr int L = 256;    // positions per attention block
constexpr int DA = 128;   // head dimension

struct Smem {
  static constexpr int Q_OFF = 0;        // [2 da-halves][128 queries][128 B]
  static constexpr int DO_OFF = 32768;
  static constexpr int K_OFF = 65536;    // [2 da-halves][128 keys][128 B]
  static constexpr int V_OFF = 98304;
  static constexpr int P_OFF = 131072;   // [2 key-halves][128 queries][128 B]; also the store staging of dQ / dK
  static constexpr int DS_OFF = 163840;  // same for dS; staging of dV
  static constexpr int BAR_OFF = 196608;
  static constexpr int BINS_OFF = BAR_OFF + 256;  // [2][64] fp32 bank-gradient bins
  static constexpr int TOTAL = BINS_OFF + 512;
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
};

struct Blk {
  int i, c;
  bool first_c, last_c, dq_in, dq_final;
};
// block order inside one z; first_c / last_c: first / last contribution to dV_c, dK_c (they also mark the K_c, V_c
// load and the drain); dq_in: a c = 0 partial of dQ_i waits in the scratch; dq_final: this block completes dQ_i
LVT_DEVICE_INLINE Blk blk_info(bool causal, int b) {
  Blk k;
  if (!causal) {
    k.i = b & 1; k.c = b >> 1;
    k.first_c = k.i == 0; k.last_c = k.i == 1;
    k.dq_in = k.c == 1; k.dq_final = k.c == 1;
  } else {
    k.i = b == 0 ? 0 : 1; k.c = b == 2 ? 1 : 0;
    k.first_c = b != 1; k.last_c = b != 0;
    k.dq_in = b == 2; k.dq_final = b != 1;
  }
  return k;
}

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
  uint64_t* qd_full = bars + 0;    // Q_i, dO_i landed
  uint64_t* kv_full = bars + 1;    // K_c, V_c landed
  uint64_t* s_full = bars + 2;     // S, dP complete in tensor memory
  uint64_t* ps_full = bars + 3;    // P, dS slabs written (8 warps)
  uint64_t* mma2_done = bars + 4;  // dV, dK, dQ MMAs complete: operands / slabs free, accumulators readable
  uint64_t* tmem_free = bars + 5;  // the epilogue has drained what the next block overwrites (8 warps)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);
  float* const bank_bins = reinterpret_cast<float*>(smem + S::BINS_OFF);
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
    mbar_init(qd_full, 1);
    mbar_init(kv_full, 1);
    mbar_init(s_full, 1);
    mbar_init(ps_full, NUM_EPI_WARPS);
    mbar_init(mma2_done, 1);
    mbar_init(tmem_free, NUM_EPI_WARPS);
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
      uint32_t n = 0;
      for (int z = blockIdx.x; z < p.nz; z += gridDim.x) {
        const int zlo = z % p.heads, zhi = z / p.heads;
        for (int b = 0; b < nblk; ++b, ++n) {
          const Blk k = blk_info(causal, b);
          mbar_wait(mma2_done, (n & 1) ^ 1);  // the previous block's second MMA group has consumed every operand
          if (k.first_c) {
            mbar_arrive_expect_tx(kv_full, 65536);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              tma_load_5d(smem + S::K_OFF + kb * 16384, &tm_k, kv_full, kb * 64, k.c * 128, 0, zlo, zhi);
              tma_load_5d(smem + S::V_OFF + kb * 16384, &tm_v, kv_full, kb * 64, k.c * 128, 0, zlo, zhi);
            }
          }
          mbar_arrive_expect_tx(qd_full, 65536);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            tma_load_5d(smem + S::Q_OFF + kb * 16384, &tm_q, qd_full, kb * 64, k.i * 128, 0, zlo, zhi);
            tma_load_5d(smem + S::DO_OFF + kb * 16384, &tm_do, qd_full, kb * 64, k.i * 128, 0, zlo, zhi);
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
      const uint32_t q_base = smem_u32(smem + S::Q_OFF), do_base = smem_u32(smem + S::DO_OFF);
      const uint32_t k_base = smem_u32(smem + S::K_OFF), v_base = smem_u32(smem + S::V_OFF);
      const uint32_t p_base = smem_u32(smem + S::P_OFF), ds_base = smem_u32(smem + S::DS_OFF);
      // K-major view: 64 contraction elements per 128 B row; MN-major view of the SAME bytes: the rows are the
      // contraction index, the two 64-wide column halves are the M/N atoms, 16 KiB apart
      auto kmaj = [](uint32_t base, int kb, int k4) { return umma_smem_desc(base + kb * 16384 + k4 * 32, 16, 1024); };
      auto mnmaj = [](uint32_t base, int k16) { return umma_smem_desc(base + k16 * 2048, 16384, 1024); };
      uint32_t n = 0, m = 0;
      for (int z = blockIdx.x; z < p.nz; z += gridDim.x) {
        if (z + (int)gridDim.x >= p.nz) pdl_launch_dependents();
        for (int b = 0; b < nblk; ++b, ++n) {
          const Blk k = blk_info(causal, b);
          if (k.first_c) {
            mbar_wait(kv_full, m & 1);
            ++m;
          }
          mbar_wait(qd_full, n & 1);
          mbar_wait(tmem_free, (n & 1) ^ 1);  // dQ (columns 0-127) and, at a change of c, dV / dK have been drained
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16_ss(tmem_base, kmaj(q_base, kb, k4), kmaj(k_base, kb, k4), id_kk, (kb | k4) ? 1u : 0u);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16_ss(tmem_base + 128, kmaj(do_base, kb, k4), kmaj(v_base, kb, k4), id_kk, (kb | k4) ? 1u : 0u);
          umma_commit(s_full);
          mbar_wait(ps_full, n & 1);  // P, dS are in shared memory; S and dP have been read
          tc_fence_after();
          const uint32_t acc0 = k.first_c ? 0u : 1u;
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16)  // dV_c[key, :] += sum_q P[q, key] dO[q, :]
            umma_bf16_ss(tmem_base + 256, mnmaj(p_base, k16), mnmaj(do_base, k16), id_mm, k16 ? 1u : acc0);
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16)  // dK_c[key, :] += sum_q dS[q, key] Q[q, :]
            umma_bf16_ss(tmem_base + 384, mnmaj(ds_base, k16), mnmaj(q_base, k16), id_mm, k16 ? 1u : acc0);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)     // dQ_i[q, :] = sum_key dS[q, key] K[key, :]
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16_ss(tmem_base, kmaj(ds_base, kb, k4), mnmaj(k_base, kb * 4 + k4), id_km, (kb | k4) ? 1u : 0u);
          umma_commit(mma2_done);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    constexpr int NY = 64 / BW;                  // key rows (of BW keys) inside this warp's 64 keys
    constexpr int NBH = 2 * BH - 1, NBW = 2 * BW - 1, NBT = 2 * BT - 1;
    const int ew = warp - 2;
    const int q = warp & 3;       // tensor-memory lane quarter: rows 32q .. 32q+31 of the block
    const int half = ew >> 2;     // which 64 keys (P, dS) / which 64 output columns (drains)
    const int row_l = q * 32 + lane;
    const float kLog2e = 1.4426950408889634f;
    const float a2 = p.scale * kLog2e;
    const float kMasked = -1e4f * kLog2e;
    uint4* const pslab = reinterpret_cast<uint4*>(smem + S::P_OFF + half * 16384 + q * 4096);
    uint4* const dslab = reinterpret_cast<uint4*>(smem + S::DS_OFF + half * 16384 + q * 4096);
    const uint32_t t_own = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 64;
    float* const scratch = p.scratch + (size_t)blockIdx.x * 32768 + (size_t)(half * 64) * 128 + row_l;
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
    uint32_t n = 0, zc = 0;
    for (int z = blockIdx.x; z < p.nz; z += gridDim.x, ++zc) {
      const int zlo = z % p.heads, zhi = z / p.heads;
      const int head = zlo;
      float* const bins = bank_bins + (zc & 1) * 64;  // [0, NBH) dh, [NBH, NBH+NBW) dw, then dt
      for (int b = 0; b < nblk; ++b, ++n) {
        const Blk k = blk_info(causal, b);
        const int qi = k.i * 128 + row_l;  // query position inside the block of 256
        const int ti = qi / (BH * BW), hi = (qi / BW) % BH, wi = qi % BW;
        const int key0 = k.c * 128 + half * 64;
        const int tj = key0 / (BH * BW), hj0 = (key0 / BW) % BH;
        const float lse2 = p.lse[(size_t)z * L + qi] * kLog2e;
        const float dl = p.delta[(size_t)z * L + qi];
        // bias slices of this row for its 64 keys, log2 domain (get_B: B[i, j] = bt[ti-tj] + bh[hi-hj] + bw[wi-wj])
        float bwv[BW], bhv[NY];
        const float btv = kLog2e * __ldg(p.bank_t + head * NBT + (ti - tj + BT - 1));
#pragma unroll
        for (int y = 0; y < NY; ++y) bhv[y] = kLog2e * __ldg(p.bank_h + head * NBH + (hi - (hj0 + y) + BH - 1));
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
        if (lane == 0) bulk_wait_group_read<0>();  // this warp's staged stores have drained its slab pieces
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t rs[32], rp[32];
          tmem_ld_32x32(t_own + 32 * ch, rs);
          tmem_ld_32x32(t_own + 128 + 32 * ch, rp);
          tmem_ld_wait();
          float pv[32], dv[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int kk = 32 * ch + e;
            const float bias = (btv + bhv[kk / BW]) + bwv[kk % BW];
            float s2 = __uint_as_float(rs[e]) * a2 + bias;
            if (diag && half * 64 + kk > row_l) s2 = kMasked;
            // P rounded to bf16 BEFORE dS: delta = rowsum(dO * O) was formed from the bf16 P of the forward, and
            // dS = P * (dP - delta) only keeps its rows summing to zero (the cancellation that dominates peaked
            // rows) when both use the same P
            pv[e] = __bfloat162float(__float2bfloat16_rn(ex2(s2 - lse2)));
            dv[e] = pv[e] * (__uint_as_float(rp[e]) - dl);
            rs_h[kk / BW] += dv[e];
            cs_w[kk % BW] += dv[e];
          }
          stage32(pslab, ch, pv);
          stage32(dslab, ch, dv);
        }
        fence_proxy_async();  // generic-proxy writes -> visible to tcgen05.mma
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ps_full);

        // fills the wait for the second MMA group: bank-gradient partial sums -> shared-memory bins
        {
          float tot = 0.f;
#pragma unroll
          for (int y = 0; y < NY; ++y) {
            float v = rs_h[y];
            tot += v;
#pragma unroll
            for (int o = BW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((lane & (BW - 1)) == 0) atomicAdd(&bins[hi - (hj0 + y) + BH - 1], v);
          }
#pragma unroll
          for (int x = 0; x < BW; ++x) {
            float v = cs_w[x];
#pragma unroll
            for (int o = 16; o >= BW; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane < BW) atomicAdd(&bins[NBH + wi - x + BW - 1], v);  // distinct bins across these lanes
          }
          tot = warp_sum(tot);  // ti is uniform inside a warp (32 | BH*BW)
          if (lane == 0) atomicAdd(&bins[NBH + NBW + ti - tj + BT - 1], tot);
        }

        mbar_wait(mma2_done, n & 1);
        tc_fence_after();
        // dQ_i: scale, (+ the parked c = 0 partial), then either park it or store it
        {
          float* const sc = scratch + (size_t)k.i * 16384;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t r[32];
            tmem_ld_32x32(t_own + 32 * ch, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]) * p.scale;
            if (k.dq_in) {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] += sc[(32 * ch + e) * 128];
            }
            if (k.dq_final) {
              stage32(pslab, ch, v);
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e) sc[(32 * ch + e) * 128] = v[e];
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
        }
        if (k.last_c) {
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
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&tm_dk, pslab, half * 64, k.c * 128 + q * 32, 0, zlo, zhi);
            bulk_commit_group();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_free);
      }
      // bank gradients of this z: one flush by the first epilogue warp; the other buffer serves the next z
      asm volatile("bar.sync 5, 256;" ::: "memory");
      if (ew == 0) {
        for (int i = lane; i < NBH + NBW + NBT; i += 32) {
          const float v = bins[i];
          bins[i] = 0.f;
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
  const int grid = (int)(nz < sm_count() ? nz : sm_count());
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
