// lvt_b200 :: VQ codebook search on the tensor cores with an EXACT fp32 re-rank.
//
// The reference result (vq_utils.py:7-24) is argmin_k of d_k = fl(fl(c2_k + x2) - 2*dot_k) with dot_k a
// sequential fp32 FMA chain (see vq.cu).  262 144 FLOP per latent position make the SIMT kernel
// FFMA-bound (1.8 % of the HBM roofline), so the 512-way scan is moved to tcgen05:
//   1. scores  s_k = c2_k - 2 * <tf32(x), tf32(c_k)>  for all 512 codes of a group: kind::tf32 MMA,
//      A = 128 positions x 64 dims read straight from the NCHW tensor (MN-major fp32 tile via TMA, no
//      conversion pass), B = -2*codebook resident in shared memory; an extra K-step with A = [1, 1, 0...],
//      B = [c2_hi, c2_lo, 0...] puts c2_k into the accumulator, so the epilogue never touches c2.
//   2. per position, the candidate set {k : s_k <= min_k s_k + 2E} with a rigorous error bound
//      E = 1.1 * 2^-8 * |x| * max_k|c_k| + 2^-18 * (|x|^2 + max|c|^2)   (tf32 truncation of both operands,
//      Cauchy-Schwarz; covers the tensor-core accumulation and the reference's own fp32 roundings).
//   3. the reference's exact arithmetic on the candidates only (typically 1-2 of 512), lowest index
//      wins ties  =>  bit-identical indices to the SIMT kernel / the oracle.
// One CTA = one codebook group (its 128 KiB of -2*c stay resident), persistent over 128-position tiles;
// warp 0 TMA producer, warp 1 MMA issuer, 16 scan warps.  Two kernels share this file:
//   vq_argmin_tc2_kernel (default): free-running pipeline -- two sets of warps on alternate tiles, one pass over
//     TMEM with a running threshold, pair-local exchanges only (see the comment above it);
//   vq_argmin_tc_kernel (round 1, LVT_VQ_TC1=1, kept for A/B runs): all 16 warps in lock step on one tile, two
//     passes over TMEM (row minimum, then threshold), CTA-wide barriers between the phases.
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "../../include/lvt_b200.h"
#include "common.cuh"

extern void lvt_count_launch(int n);

namespace {

constexpr int TK = 512;   // codes per group
constexpr int TD = 64;    // dims per group
constexpr int TM = 128;   // positions per tile
constexpr int NSLOT = 2;
constexpr int NGRP = 4;                     // scan groups: 128 codes (TMEM columns) each
constexpr int NEW = 4 * NGRP;               // scan warps
constexpr int WL_CAP = 512;                 // candidates kept per tile; more => exhaustive exact scan
constexpr int A_BYTES = TM * TD * 4;        // 32 KiB
constexpr int B_BYTES = 2 * TK * 128;       // two k-blocks of 32 fp32 (128 B rows)
constexpr int SM_A = B_BYTES;
constexpr int SM_BX = SM_A + NSLOT * A_BYTES;   // 512 x 32 B rows [c2_hi, c2_lo, 0...] (SWIZZLE_32B, K-major)
constexpr int SM_AX = SM_BX + TK * 32;          // 128 x 32 B rows [1, 1, 0...]
constexpr int SM_C2 = SM_AX + TM * 32;          // exact fp32 c2 table
constexpr int SM_WL = SM_C2 + TK * 4;           // [WL_CAP] u32 (row << 16 | code) candidates of the tile
constexpr int SM_XMIN = SM_WL + WL_CAP * 4;     // [NGRP][128] per-group minima
constexpr int SM_X2 = SM_XMIN + NGRP * TM * 4;      // [128] |x|^2, [8][128] its per-vector-lane partial sums
constexpr int SM_RB = SM_X2 + 9 * TM * 4;           // [2][128] u64 (ordered d << 32 | code) per-row winners
constexpr int SM_WCNT = SM_RB + 2 * TM * 8;         // candidate counter
constexpr int SM_RCNT = SM_WCNT + 64;               // [128] candidates per row
constexpr int SM_BAR = SM_RCNT + TM * 4;
constexpr int SM_TOTAL = SM_BAR + 256 + 1024;
static_assert(SM_TOTAL <= 232448, "shared memory budget");
constexpr int TC_THREADS = 64 + NEW * 32;   // TMA warp, MMA warp, 16 scan warps

LVT_DEVICE_INLINE void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
LVT_DEVICE_INLINE void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
LVT_DEVICE_INLINE void epi_bar(int id) { asm volatile("bar.sync %0, 512;" ::"r"(id) : "memory"); }
LVT_DEVICE_INLINE float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// w |= bit when v <= thr (two instructions per score)
#define VQ_TEST(w, v, thr, bit) \
  asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(w) : "f"(v), "f"(thr), "r"(bit))

// ATen-order squared norm (see vq.cu): 8-lane vectors into 4 accumulators, combined, lanes left to right
template <typename LoadFn>
LVT_DEVICE_INLINE float sqnorm64(LoadFn ld) {
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const float x = ld(i * 8 + l);
      const float sq = __fmul_rn(x, x);
      if (i < 4) acc[i][l] = sq;
      else acc[i - 4][l] = __fadd_rn(acc[i - 4][l], sq);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int l = 0; l < 8; ++l) {
    const float a = __fadd_rn(__fadd_rn(__fadd_rn(acc[0][l], acc[1][l]), acc[2][l]), acc[3][l]);
    s = (l == 0) ? a : __fadd_rn(s, a);
  }
  return s;
}

// Shared-memory matrix descriptor with an explicit layout type
// (1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B, 6 = SWIZZLE_32B).
LVT_DEVICE_INLINE uint64_t smem_desc_lt(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
LVT_DEVICE_INLINE uint32_t ordered_f32(float d) {  // order-preserving float -> uint map
  const uint32_t u = __float_as_uint(d);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// NHWC = true : z_e is channels-last [positions, num*64] (the VQ-VAE engine's layout): A is K-major,
//               two 128 x 32 fp32 k-blocks, SWIZZLE_128B.
// NHWC = false: z_e is NCHW (the reference's layout): A is MN-major; tf32 MN-major operands only exist in
//               the 32B-atom 128B swizzle (TMA SWIZZLE_128B_ATOM_32B <-> UMMA SWIZZLE_128B_BASE32B).
template <bool NHWC>
__global__ void __launch_bounds__(TC_THREADS, 1)
vq_argmin_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ codebook,
                    int64_t* __restrict__ idx_out, float* __restrict__ zq_out,
                    __nv_bfloat16* __restrict__ zq_bf16, float* __restrict__ counts,
                    float* __restrict__ sums, int num, int hw, int num_tiles, int ctas_per_group,
                    float* __restrict__ dbg, unsigned* __restrict__ clk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  float* c2s = reinterpret_cast<float*>(smem + SM_C2);
  uint32_t* wl_all = reinterpret_cast<uint32_t*>(smem + SM_WL);
  float* xmin = reinterpret_cast<float*>(smem + SM_XMIN);
  float* x2s = reinterpret_cast<float*>(smem + SM_X2);
  float* lsum = x2s + TM;
  unsigned long long* rowbest = reinterpret_cast<unsigned long long*>(smem + SM_RB);
  int* wcnt = reinterpret_cast<int*>(smem + SM_WCNT);
  int* rowcnt = reinterpret_cast<int*>(smem + SM_RCNT);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + SM_BAR);  // [NSLOT]
  uint64_t* a_empty = a_full + NSLOT;                              // [NSLOT]
  uint64_t* t_full = a_empty + NSLOT;                              // [2]
  uint64_t* t_empty = t_full + 2;                                  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* cmax2_s = reinterpret_cast<float*>(tmem_ptr_smem + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x % num;
  const int sub = blockIdx.x / num;
  const float* cbg = codebook + (size_t)g * TK * TD;
  const int tiles_per_frame = hw / TM;
  const int C = num * TD;
#define VQ_CLK(who, ev) do { if (clk && blockIdx.x == 0 && it < 24 && lane == 0) clk[(it * 4 + (who)) * 16 + (ev)] = (unsigned)clock(); } while (0)

  // ---- one-time setup: B = -2 * codebook[g] in the K-major 128B-swizzled UMMA layout; exact c2 table; the
  //      extra K-step operands that put c2 into the accumulator: A_x = [1, 1, 0...], B_x = [c2_hi, c2_lo, 0...]
  for (int i = threadIdx.x; i < TK * TD / 4; i += TC_THREADS) {
    const int k = i >> 4, jj = i & 15;  // code row, 16-byte chunk (4 dims) of the 64-dim row
    float4 v = __ldg(reinterpret_cast<const float4*>(cbg) + i);
    v.x *= -2.f; v.y *= -2.f; v.z *= -2.f; v.w *= -2.f;
    const int kb = jj >> 3, chunk = jj & 7;
    *reinterpret_cast<float4*>(smem + kb * (TK * 128) + k * 128 + ((chunk ^ (k & 7)) << 4)) = v;
  }
  float cm = 0.f;
  for (int k = threadIdx.x; k < TK; k += TC_THREADS) {
    const float* row = cbg + (size_t)k * TD;
    const float c2 = sqnorm64([&](int j) { return __ldg(row + j); });
    c2s[k] = c2;
    cm = fmaxf(cm, c2);
    const float hi = __uint_as_float(__float_as_uint(c2) & 0xFFFFE000u);  // tf32-exact part
    const float lo = c2 - hi;
    const int sw = (k >> 2) & 1;  // SWIZZLE_32B: 16 B chunk index ^= bit 2 of the row
    *reinterpret_cast<float4*>(smem + SM_BX + k * 32 + (sw << 4)) = make_float4(hi, lo, 0.f, 0.f);
    *reinterpret_cast<float4*>(smem + SM_BX + k * 32 + ((sw ^ 1) << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r = threadIdx.x; r < TM; r += TC_THREADS) {
    const int sw = (r >> 2) & 1;
    *reinterpret_cast<float4*>(smem + SM_AX + r * 32 + (sw << 4)) = make_float4(1.f, 1.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(smem + SM_AX + r * 32 + ((sw ^ 1) << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = threadIdx.x; i < 2 * TM; i += TC_THREADS) rowbest[i] = ~0ull;
  if (threadIdx.x == 0) *wcnt = 0;
  if (threadIdx.x < TM) rowcnt[threadIdx.x] = 0;
  cm = warp_max(cm);
  if (threadIdx.x == 0) {
    *cmax2_s = 0.f;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], NEW);
    }
    mbar_init(&t_full[0], 1);
    mbar_init(&t_full[1], 1);
    mbar_init(&t_empty[0], NEW / 2);
    mbar_init(&t_empty[1], NEW / 2);
    fence_barrier_init();
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  __syncthreads();
  if (lane == 0) atomicMax(reinterpret_cast<int*>(cmax2_s), __float_as_int(cm));  // cm >= 0: int order == float order
  fence_proxy_async();  // operands written through the generic proxy, read by tcgen05.mma (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const float cmax2 = *cmax2_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;
      for (int tile = sub; tile < num_tiles; tile += ctas_per_group, ++it) {
        const int slot = it % NSLOT;
        mbar_wait(&a_empty[slot], ((it / NSLOT) & 1) ^ 1);
        mbar_arrive_expect_tx(&a_full[slot], A_BYTES);
        uint8_t* dst = smem + SM_A + slot * A_BYTES;
        if (NHWC) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)  // 128 positions x 32 dims per box (128 B rows, one row per position)
            tma_load_2d(dst + kb * 16384, &tm_x, &a_full[slot], g * TD + kb * 32, tile * TM);
        } else {
          const int frame = tile / tiles_per_frame, s0 = (tile - frame * tiles_per_frame) * TM;
#pragma unroll
          for (int a = 0; a < 4; ++a)  // 32 positions x 64 dims per box (128 B rows, one row per dim)
            tma_load_2d(dst + a * 8192, &tm_x, &a_full[slot], s0 + 32 * a, frame * C + g * TD);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (kind::tf32)
    if (elect_one()) {
      // idesc: D f32, A/B tf32, A MN-major (NCHW) or K-major, B K-major, N = 256, M = 128
      constexpr uint32_t idesc = umma_idesc(TM, 256, /*tf32*/ 2, !NHWC, false);
      constexpr uint32_t idesc_x = umma_idesc(TM, 256, /*tf32*/ 2, false, false);
      const uint32_t b_base = smem_u32(smem);
      const uint64_t ax_desc = smem_desc_lt(smem_u32(smem + SM_AX), 16, 256, 6);
      int it = 0;
      for (int tile = sub; tile < num_tiles; tile += ctas_per_group, ++it) {
        const int slot = it % NSLOT;
        mbar_wait(&a_full[slot], (it / NSLOT) & 1);
        VQ_CLK(0, 0);
        const uint32_t a_base = smem_u32(smem + SM_A + slot * A_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&t_empty[h], (it & 1) ^ 1);  // scan of the previous tile's buffer h finished
          VQ_CLK(0, 1 + h);
          tc_fence_after();
          // K-step 0 writes c2 (= 1*c2_hi + 1*c2_lo) into the accumulator
          umma_tf32_ss(tmem_base + h * 256, ax_desc, smem_desc_lt(smem_u32(smem + SM_BX + h * 256 * 32), 16, 256, 6),
                       idesc_x, 0u);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // K = 8 per tf32 MMA
            // K-major: 32 B per K-step inside the 128 B row, next k-block after 4 steps.
            // MN-major (BASE32B): atoms of 4 k-rows x 128 B (SBO 512), 32-position atoms 8 KiB apart (LBO).
            const uint64_t adesc = NHWC ? smem_desc_lt(a_base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, 2)
                                        : smem_desc_lt(a_base + ks * 1024, 8192, 512, 1);
            const uint64_t bdesc = umma_smem_desc(b_base + (ks >> 2) * (TK * 128) + h * (256 * 128) + (ks & 3) * 32, 16, 1024);
            umma_tf32_ss(tmem_base + h * 256, adesc, bdesc, idesc, 1u);
          }
          umma_commit(&t_full[h]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ scan + exact re-rank
    // Four groups of four warps; group grp scans codes grp*128 .. grp*128+127 (TMEM buffer grp/2) of every tile.
    // A row whose window holds a single code needs no re-rank at all (that code is the exact argmin).
    const int ew = warp - 2;
    const int grp = ew >> 2;
    const int h = grp >> 1;
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;  // row (position) inside the tile
    const int et = ew * 32 + lane;  // scan thread id 0..511
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + grp * 128;
    uint32_t* wl = wl_all;
    int it = 0;
    for (int tile = sub; tile < num_tiles; tile += ctas_per_group, ++it) {
      const int slot = it % NSLOT;
      const uint32_t pos = (uint32_t)tile * TM + m;  // global position index (frame * hw + s) < 2^31
      const uint32_t frame = pos / (uint32_t)hw, s = pos - frame * (uint32_t)hw;
      const uint8_t* a_tile = smem + SM_A + slot * A_BYTES;
      // element j of row r of the (swizzled) tile
      auto xptr = [&](int r) { return NHWC ? a_tile + r * 128 : a_tile + (r >> 5) * 8192 + (r & 7) * 4; };
      auto xelem = [&](const uint8_t* xr, int rsw, int j) {
        return NHWC ? *reinterpret_cast<const float*>(xr + (j >> 5) * 16384 + ((((j >> 2) & 7) ^ rsw) << 4) + (j & 3) * 4)
                    : *reinterpret_cast<const float*>(xr + j * 128 + ((rsw ^ (j & 3)) << 5));
      };
      // exact reference distance (vq.cu): sequential fp32 FMA chain over the 64 dims, d = fl(fl(c2 + x2) - 2*dot).
      // The code row comes from global memory (L1 / L2 resident, 16 independent 16-byte loads issued up front):
      // while the re-rank runs the next tile's MMAs take most of the shared-memory bandwidth.
      auto exact_d = [&](int r, int k) {
        const float4* crow = reinterpret_cast<const float4*>(cbg + (size_t)k * TD);
        float4 cv[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) cv[jj] = __ldg(crow + jj);
        const uint8_t* xr = xptr(r);
        const int rsw = NHWC ? (r & 7) : ((r & 31) >> 3);
        float acc = 0.f;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          float4 xv;
          if (NHWC) {
            xv = *reinterpret_cast<const float4*>(xr + (jj >> 3) * 16384 + (((jj & 7) ^ rsw) << 4));
          } else {
            xv.x = *reinterpret_cast<const float*>(xr + (4 * jj) * 128 + ((rsw ^ 0) << 5));
            xv.y = *reinterpret_cast<const float*>(xr + (4 * jj + 1) * 128 + ((rsw ^ 1) << 5));
            xv.z = *reinterpret_cast<const float*>(xr + (4 * jj + 2) * 128 + ((rsw ^ 2) << 5));
            xv.w = *reinterpret_cast<const float*>(xr + (4 * jj + 3) * 128 + ((rsw ^ 3) << 5));
          }
          acc = __fmaf_rn(xv.x, cv[jj].x, acc);
          acc = __fmaf_rn(xv.y, cv[jj].y, acc);
          acc = __fmaf_rn(xv.z, cv[jj].z, acc);
          acc = __fmaf_rn(xv.w, cv[jj].w, acc);
        }
        return __fadd_rn(-2.f * acc, __fadd_rn(c2s[k], x2s[r]));  // -2 * acc is exact
      };
      const int who = ew == 0 ? 1 : (ew == 12 ? 2 : (ew == 5 ? 3 : -1));
#define VQ_ECLK(ev) do { if (who > 0) VQ_CLK(who, ev); } while (0)
      VQ_ECLK(0);
      mbar_wait(&a_full[slot], (it / NSLOT) & 1);
      VQ_ECLK(1);
      {
        // |x|^2 in ATen's order (vq.cu) = sum over the 8 vector lanes l (left to right) of
        // ((a0+a1)+a2)+a3, a_j = x[8j+l]^2 + x[8(j+4)+l]^2.  The four threads of a row take two lanes each.
        const uint8_t* xr = xptr(m);
        const int rsw = NHWC ? (m & 7) : ((m & 31) >> 3);
#pragma unroll
        for (int ll = 0; ll < 2; ++ll) {
          const int l = 2 * grp + ll;
          float a[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x0 = xelem(xr, rsw, 8 * j + l), x1 = xelem(xr, rsw, 8 * (j + 4) + l);
            a[j] = __fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1));
          }
          lsum[l * TM + m] = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[1]), a[2]), a[3]);
        }
      }
      VQ_ECLK(2);
      mbar_wait(&t_full[h], it & 1);
      tc_fence_after();
      VQ_ECLK(3);
      // pass 1: minimum tf32 score over this group's 128 codes
      {
        float m0 = INFINITY, m1 = INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(t_addr + c * 32, r0);
          tmem_ld_32x32(t_addr + c * 32 + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            m0 = fmin3(m0, __uint_as_float(r0[i]), __uint_as_float(r0[i + 1]));
            m1 = fmin3(m1, __uint_as_float(r1[i]), __uint_as_float(r1[i + 1]));
          }
        }
        xmin[grp * TM + m] = fminf(m0, m1);
      }
      VQ_ECLK(4);
      epi_bar(1);
      VQ_ECLK(5);
      float x2 = lsum[m];
#pragma unroll
      for (int l = 1; l < 8; ++l) x2 = __fadd_rn(x2, lsum[l * TM + m]);
      const float E = 1.1f * 0.00390625f * sqrtf(x2 * cmax2) + 3.8146973e-6f * (x2 + cmax2);
      if (grp == 0) x2s[m] = x2;
      const float gthr = fminf(fminf(xmin[m], xmin[TM + m]), fminf(xmin[2 * TM + m], xmin[3 * TM + m])) + 2.f * E;
      // pass 2: candidates {k : s_k <= min + 2E}
      uint32_t w[4];
#pragma unroll
      for (int c = 0; c < 4; c += 2) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_addr + c * 32, r0);
        tmem_ld_32x32(t_addr + c * 32 + 32, r1);
        tmem_ld_wait();
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          VQ_TEST(w0, __uint_as_float(r0[i]), gthr, 1u << i);
          VQ_TEST(w1, __uint_as_float(r1[i]), gthr, 1u << i);
        }
        w[c] = w0;
        w[c + 1] = w1;
        if (dbg && tile == 0 && g == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            dbg[(size_t)m * TK + grp * 128 + c * 32 + i] = __uint_as_float(r0[i]);
            dbg[(size_t)m * TK + grp * 128 + c * 32 + 32 + i] = __uint_as_float(r1[i]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[h]);  // TMEM buffer h may be overwritten by the next tile's MMAs
      {
        // one shared-memory atomic per warp: exclusive scan of the lanes' candidate counts
        const int c = __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        int base = 0;
        if (lane == 31 && incl > 0) base = atomicAdd(wcnt, incl);
        base = __shfl_sync(0xffffffffu, base, 31);
        if (c) {
          atomicAdd(&rowcnt[m], c);
          int p = base + incl - c;
          const uint32_t ent0 = ((uint32_t)m << 16) | (uint32_t)(grp * 128);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t ww = w[cc];
            while (ww) {
              const int b = __ffs(ww) - 1;
              ww &= ww - 1;
              if (p < WL_CAP) wl[p] = ent0 + cc * 32 + b;
              ++p;
            }
          }
        }
      }
      VQ_ECLK(6);
      // exact reference arithmetic on the candidates, spread over all scan threads; winner per row = min (d, code)
      epi_bar(2);
      VQ_ECLK(7);
      unsigned long long* rb = rowbest + (it & 1) * TM;
      const int n = *wcnt;  // candidates of the tile
      if (n <= WL_CAP) {
        for (int e = et; e < n; e += NEW * 32) {
          const uint32_t ent = wl[e];
          const int r = ent >> 16, k = ent & 0xFFFF;
          if (rowcnt[r] == 1) rb[r] = (unsigned long long)k;  // the only code inside the window IS the exact argmin
          else atomicMin(&rb[r], ((unsigned long long)ordered_f32(exact_d(r, k)) << 32) | (unsigned)k);
        }
      } else {  // pathological codebook (dozens of near-ties per row): every code, exactly
        for (int k = grp * 128; k < grp * 128 + 128; ++k)
          atomicMin(&rb[m], ((unsigned long long)ordered_f32(exact_d(m, k)) << 32) | (unsigned)k);
      }
      VQ_ECLK(8);
      epi_bar(3);
      VQ_ECLK(9);
      if (et == 0) *wcnt = 0;  // next appends come after the next tile's barrier 1
      const int besti = (int)(rb[m] & 0xFFFFFFFFull);
      if (grp == 0) {
        idx_out[((size_t)frame * num + g) * hw + s] = (int64_t)besti;
      } else if (grp == 1) {
        if (zq_out || zq_bf16) {
          const float* cr = cbg + (size_t)besti * TD;
          if (NHWC) {
            const size_t o = (size_t)pos * C + (size_t)g * TD;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(cr) + jj);
              if (zq_out) *reinterpret_cast<float4*>(zq_out + o + 4 * jj) = v;
              if (zq_bf16) {
                uint2 u;
                u.x = pack_bf16x2(v.x, v.y);
                u.y = pack_bf16x2(v.z, v.w);
                *reinterpret_cast<uint2*>(zq_bf16 + o + 4 * jj) = u;
              }
            }
          } else {
            float* zp = zq_out + ((size_t)frame * C + (size_t)g * TD) * hw + s;
#pragma unroll 8
            for (int j = 0; j < TD; ++j) zp[(size_t)j * hw] = __ldg(cr + j);
          }
        }
      } else if (grp == 2) {
        if (counts) atomicAdd(counts + (size_t)g * TK + besti, 1.f);
        if (sums) {
          float* sp = sums + ((size_t)g * TK + besti) * TD;
          const uint8_t* xr = xptr(m);
          const int rsw = NHWC ? (m & 7) : ((m & 31) >> 3);
#pragma unroll 8
          for (int j = 0; j < TD; ++j) atomicAdd(sp + j, xelem(xr, rsw, j));
        }
      } else {
        rowbest[((it + 1) & 1) * TM + m] = ~0ull;  // last read before this tile's first barrier
        rowcnt[m] = 0;                             // next increments come after the next tile's barrier 1
      }
      VQ_ECLK(10);
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_empty[slot]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------
// v2: the same algorithm as a free-running pipeline (no CTA-wide barrier inside the tile loop).
//   * the 16 scan warps form two SETS that take alternate tiles (set = tile parity = A slot); a set is four
//     PAIRS of warps, one pair per TMEM lane quarter: the pair owns rows q*32.. of the set's tiles, warp pw of
//     the pair scans its half of the codes of every accumulator buffer.  All exchanges (|x|^2 partial sums,
//     minima, candidate lists, winners) stay inside the pair: four 64-thread named barriers per tile.
//   * ONE pass over TMEM: each 32-column chunk is tested against the RUNNING minimum + 2E while it is in
//     registers (a superset of the final window: the running minimum only decreases), the buffer is released
//     to the MMA warp right after, and chunks whose own minimum is above the final threshold are dropped later.
//   * the threshold tests are split over the two fp32 pipes: NF of 32 scores as
//     flag = sat((thr - s) * 2^64) (FFMA.SAT with an immediate multiplier, exactly 0 or 1 by construction of thr)
//     accumulated into a float bit mask by a second FFMA, the rest as SETP + predicated LOP3 on the ALU pipe,
//     next to the FMNMX3 tree; the two kinds are interleaved in program order.
//   * candidates: up to three codes per lane are extracted branch-free into registers and placed into a
//     per-warp list by ballots (no atomics); the two warps of a pair share both lists evenly for the exact
//     re-rank, which reads the position from the A tile and the code row from the resident B operand
//     (-2c, halved exactly); a row whose window holds one code needs no re-rank at all.
//   * the slot's next tile is prefetched into L2 when a tile is loaded; the A slot is released after the re-rank.
// Measured on the way and not kept: re-rank operands from global memory with an early A release,
// one warp per TMEM lane quarter scanning all 512 codes (three tiles in flight, no pair barriers: a single
// scanning warp per scheduler issues one instruction every ~4 cycles), four 128-column accumulator buffers.
constexpr int WL2_CAP = 96;                         // list entries per warp and tile: three per lane
constexpr int S2_XP = SM_AX + TM * 32;              // [2 sets][5][128] f32: S0 of warp 0, L4..L7 of warp 1
constexpr int S2_X2 = S2_XP + 2 * 5 * TM * 4;       // [2][128] |x|^2
constexpr int S2_PMIN = S2_X2 + 2 * TM * 4;         // [2][2][128] minimum of each warp's 256 codes
constexpr int S2_RB = S2_PMIN + 2 * 2 * TM * 4;     // [2][128] u64 winners (ordered d << 32 | code)
constexpr int S2_WL = S2_RB + 2 * TM * 8;           // [16 warps][WL2_CAP] u16 (lane << 9 | code)
constexpr int S2_WN = S2_WL + 16 * WL2_CAP * 2;      // [16] entries in each warp's list
constexpr int S2_BAR = S2_WN + 64;
constexpr int S2_TOTAL = S2_BAR + 256 + 1024;
static_assert(S2_TOTAL <= 232448, "shared memory budget (v2)");
static_assert(S2_RB % 8 == 0 && S2_BAR % 8 == 0, "alignment");

// sat(thrH - s * 2^64): the multiplier is an immediate (FFMA imm-form issues at twice the rate of the 3-register form)
LVT_DEVICE_INLINE float flag_below(float s, float thrH) {
  float r;
  asm("fma.rn.sat.f32 %0, %1, 0fDF800000, %2;" : "=f"(r) : "f"(s), "f"(thrH));
  return r;
}
LVT_DEVICE_INLINE void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
#define VQ_TEST_LT(w, v, thr, bit) \
  asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(w) : "f"(v), "f"(thr), "r"(bit))

// One 32-column chunk in registers: its minimum, the running minimum, and the mask of scores below
// thr' = (running minimum + 2E) nudged up by >= 1 ulp and kept out of (-2^-40, 2^-40), so that for every fp32
// score s the product (thr' - s) * 2^64 is either <= 0 or >= 1: the saturated FFMA is an exact 0 / 1 flag.
template <int NF>
LVT_DEVICE_INLINE uint32_t scan_chunk(const uint32_t (&r)[32], float& rmin, const float twoE, float& submin) {
#define F(i) __uint_as_float(r[i])
  const float t0 = fmin3(F(0), F(1), F(2)), t1 = fmin3(F(3), F(4), F(5)), t2 = fmin3(F(6), F(7), F(8));
  const float t3 = fmin3(F(9), F(10), F(11)), t4 = fmin3(F(12), F(13), F(14)), t5 = fmin3(F(15), F(16), F(17));
  const float t6 = fmin3(F(18), F(19), F(20)), t7 = fmin3(F(21), F(22), F(23)), t8 = fmin3(F(24), F(25), F(26));
  const float t9 = fmin3(F(27), F(28), F(29)), t10 = fminf(F(30), F(31));
  const float u0 = fmin3(t0, t1, t2), u1 = fmin3(t3, t4, t5), u2 = fmin3(t6, t7, t8), u3 = fmin3(t9, t10, u0);
  submin = fmin3(u1, u2, u3);
  rmin = fminf(rmin, submin);
  float thr = rmin + twoE;
  thr = thr + fmaxf(fabsf(thr) * 2.4e-7f, 1e-18f);
  thr = fabsf(thr) < 1e-12f ? 1e-12f : thr;
  constexpr float H = 18446744073709551616.f;  // 2^64
  const float thrH = thr * H;
  (void)H;
  // the two kinds of test are interleaved in program order so that both pipes have work at any time; the float
  // masks hold 16 flags each (weights 2^0 .. 2^15: exact), two accumulators per mask for shorter chains
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
  uint32_t w = 0;
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    const int nfb = t * NF / 32;
    if ((t + 1) * NF / 32 > nfb) {  // scores 0 .. NF-1: fma pipe
      const float fl = flag_below(F(nfb), thrH);
      const float wt = (float)(1u << (nfb & 15));
      if (nfb < 16) {
        if (nfb & 1) a01 = __fmaf_rn(fl, wt, a01);
        else a00 = __fmaf_rn(fl, wt, a00);
      } else {
        if (nfb & 1) a11 = __fmaf_rn(fl, wt, a11);
        else a10 = __fmaf_rn(fl, wt, a10);
      }
    } else {                        // scores NF .. 31: alu pipe
      const int e = NF + t - nfb;
      VQ_TEST_LT(w, F(e), thr, 1u << e);
    }
  }
#undef F
  w |= (uint32_t)(a00 + a01);
  if (NF > 16) w |= (uint32_t)(a10 + a11) << 16;
  return w;
}

// <x, c> as the reference's sequential fp32 FMA chain over the 64 dims (vq.cu), both operands from shared memory:
// the position from the (swizzled) A tile, the code row from the resident B operand, which holds -2*c -- halving
// is exact, so c is recovered bit for bit (an L2 round trip per candidate was the long pole of the re-rank).
// Not inlined: its staging registers stay out of the register allocation of the scan loop.
template <bool NHWC>
__device__ __forceinline__ float exact_dot_smem(const uint8_t* xr, int rsw, const uint8_t* bsm, int k) {
  float4 cv[16];
#pragma unroll
  for (int jj = 0; jj < 16; ++jj)
    cv[jj] = *reinterpret_cast<const float4*>(bsm + (jj >> 3) * (TK * 128) + k * 128 + (((jj & 7) ^ (k & 7)) << 4));
  float acc = 0.f;
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) {
    float4 xv;
    if (NHWC) {
      xv = *reinterpret_cast<const float4*>(xr + (jj >> 3) * 16384 + (((jj & 7) ^ rsw) << 4));
    } else {
      xv.x = *reinterpret_cast<const float*>(xr + (4 * jj) * 128 + ((rsw ^ 0) << 5));
      xv.y = *reinterpret_cast<const float*>(xr + (4 * jj + 1) * 128 + ((rsw ^ 1) << 5));
      xv.z = *reinterpret_cast<const float*>(xr + (4 * jj + 2) * 128 + ((rsw ^ 2) << 5));
      xv.w = *reinterpret_cast<const float*>(xr + (4 * jj + 3) * 128 + ((rsw ^ 3) << 5));
    }
    acc = __fmaf_rn(xv.x, -0.5f * cv[jj].x, acc);
    acc = __fmaf_rn(xv.y, -0.5f * cv[jj].y, acc);
    acc = __fmaf_rn(xv.z, -0.5f * cv[jj].z, acc);
    acc = __fmaf_rn(xv.w, -0.5f * cv[jj].w, acc);
  }
  return acc;
}

template <bool NHWC, int NF, int NQ>
__global__ void __launch_bounds__(TC_THREADS, 1)
vq_argmin_tc2_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ z_e,
                     const float* __restrict__ codebook, int64_t* __restrict__ idx_out, float* __restrict__ zq_out,
                     __nv_bfloat16* __restrict__ zq_bf16, float* __restrict__ counts,
                     float* __restrict__ sums, int num, int hw, int num_tiles, int ctas_per_group,
                     int prefetch, unsigned* __restrict__ clk) {
  static_assert(NF % 2 == 0 && NF <= 32, "flags accumulated in two 16-bit float masks");
  static_assert(NQ == 2 || NQ == 4, "TMEM accumulator buffers per tile");
  constexpr int CB = TK / NQ;       // codes (TMEM columns) per accumulator buffer
  constexpr int CH = 8 / NQ;        // 32-column chunks of a buffer that one warp of a pair scans
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + S2_BAR);  // [2]  slot == set == tile parity
  uint64_t* a_empty = a_full + 2;                                  // [2]
  uint64_t* t_full = a_empty + 2;                                  // [NQ buffers][2 sets]
  uint64_t* t_empty = t_full + 8;                                  // [NQ buffers]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(t_empty + 4);
  float* cmax2_s = reinterpret_cast<float*>(tmem_ptr_smem + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x % num;
  const int sub = blockIdx.x / num;
  const float* cbg = codebook + (size_t)g * TK * TD;
  const int tiles_per_frame = hw / TM;
  const int C = num * TD;

  // ---- one-time setup (as v1): B = -2 * codebook[g] (K-major, 128B swizzle), B_x = [c2_hi, c2_lo, 0...], A_x = [1, 1, 0...]
  for (int i = threadIdx.x; i < TK * TD / 4; i += TC_THREADS) {
    const int k = i >> 4, jj = i & 15;
    float4 v = __ldg(reinterpret_cast<const float4*>(cbg) + i);
    v.x *= -2.f; v.y *= -2.f; v.z *= -2.f; v.w *= -2.f;
    const int kb = jj >> 3, chunk = jj & 7;
    *reinterpret_cast<float4*>(smem + kb * (TK * 128) + k * 128 + ((chunk ^ (k & 7)) << 4)) = v;
  }
  float cm = 0.f;
  for (int k = threadIdx.x; k < TK; k += TC_THREADS) {
    float4 rv[16];  // the whole row in 16 vector loads (64 scalar loads per thread throttle the load queue)
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) rv[jj] = __ldg(reinterpret_cast<const float4*>(cbg + (size_t)k * TD) + jj);
    const float c2 = sqnorm64([&](int j) {
      const float4 v = rv[j >> 2];
      return (j & 3) == 0 ? v.x : ((j & 3) == 1 ? v.y : ((j & 3) == 2 ? v.z : v.w));
    });
    cm = fmaxf(cm, c2);
    const float hi = __uint_as_float(__float_as_uint(c2) & 0xFFFFE000u);  // tf32-exact part; hi + lo == c2 exactly
    const float lo = c2 - hi;
    const int sw = (k >> 2) & 1;
    *reinterpret_cast<float4*>(smem + SM_BX + k * 32 + (sw << 4)) = make_float4(hi, lo, 0.f, 0.f);
    *reinterpret_cast<float4*>(smem + SM_BX + k * 32 + ((sw ^ 1) << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r = threadIdx.x; r < TM; r += TC_THREADS) {
    const int sw = (r >> 2) & 1;
    *reinterpret_cast<float4*>(smem + SM_AX + r * 32 + (sw << 4)) = make_float4(1.f, 1.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(smem + SM_AX + r * 32 + ((sw ^ 1) << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = threadIdx.x; i < 2 * TM; i += TC_THREADS) reinterpret_cast<unsigned long long*>(smem + S2_RB)[i] = ~0ull;
  cm = warp_max(cm);
  if (threadIdx.x == 0) {
    *cmax2_s = 0.f;
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 8);
    }
    for (int s = 0; s < NQ; ++s) mbar_init(&t_empty[s], 8);
    for (int s = 0; s < 2 * NQ; ++s) mbar_init(&t_full[s], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  __syncthreads();
  if (lane == 0) atomicMax(reinterpret_cast<int*>(cmax2_s), __float_as_int(cm));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const float cmax2 = *cmax2_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;
      for (int tile = sub; tile < num_tiles; tile += ctas_per_group, ++it) {
        const int slot = it & 1;
        mbar_wait(&a_empty[slot], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&a_full[slot], A_BYTES);
        uint8_t* dst = smem + SM_A + slot * A_BYTES;
        if (NHWC) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
            tma_load_2d(dst + kb * 16384, &tm_x, &a_full[slot], g * TD + kb * 32, tile * TM);
        } else {
          const int frame = tile / tiles_per_frame, s0 = (tile - frame * tiles_per_frame) * TM;
#pragma unroll
          for (int a = 0; a < 4; ++a)
            tma_load_2d(dst + a * 8192, &tm_x, &a_full[slot], s0 + 32 * a, frame * C + g * TD);
        }
        // the slot's next tile goes to L2 now: its load (issued when this tile's re-rank is over) then costs an
        // L2 hit instead of an HBM round trip
        const int nt = tile + 2 * ctas_per_group;
        if (prefetch && nt < num_tiles) {
          if (NHWC) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) tma_prefetch_2d(&tm_x, g * TD + kb * 32, nt * TM);
          } else {
            const int frame = nt / tiles_per_frame, s0 = (nt - frame * tiles_per_frame) * TM;
#pragma unroll
            for (int a = 0; a < 4; ++a) tma_prefetch_2d(&tm_x, s0 + 32 * a, frame * C + g * TD);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (kind::tf32)
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(TM, CB, /*tf32*/ 2, !NHWC, false);
      constexpr uint32_t idesc_x = umma_idesc(TM, CB, /*tf32*/ 2, false, false);
      const uint32_t b_base = smem_u32(smem);
      const uint64_t ax_desc = smem_desc_lt(smem_u32(smem + SM_AX), 16, 256, 6);
      int it = 0;
      for (int tile = sub; tile < num_tiles; tile += ctas_per_group, ++it) {
        const int slot = it & 1;
        mbar_wait(&a_full[slot], (it >> 1) & 1);
        if (clk && blockIdx.x == 0 && it < 24) clk[(it * 4) * 16 + 0] = (unsigned)clock();
        const uint32_t a_base = smem_u32(smem + SM_A + slot * A_BYTES);
#pragma unroll
        for (int h = 0; h < NQ; ++h) {
          mbar_wait(&t_empty[h], (it & 1) ^ 1);  // the other set has finished its pass over buffer h
          if (clk && blockIdx.x == 0 && it < 24 && h < 2) clk[(it * 4) * 16 + 1 + h] = (unsigned)clock();
          tc_fence_after();
          umma_tf32_ss(tmem_base + h * CB, ax_desc, smem_desc_lt(smem_u32(smem + SM_BX + h * CB * 32), 16, 256, 6),
                       idesc_x, 0u);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t adesc = NHWC ? smem_desc_lt(a_base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, 2)
                                        : smem_desc_lt(a_base + ks * 1024, 8192, 512, 1);
            const uint64_t bdesc = umma_smem_desc(b_base + (ks >> 2) * (TK * 128) + h * (CB * 128) + (ks & 3) * 32, 16, 1024);
            umma_tf32_ss(tmem_base + h * CB, adesc, bdesc, idesc, 1u);
          }
          umma_commit(&t_full[h * 2 + slot]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ scan + exact re-rank, pair-local
    const int sw_ = warp - 2;
    const int set = sw_ >> 3;
    const int pw = (sw_ >> 2) & 1;
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;  // row (position) inside the tile
    const int pair_bar = 1 + set * 4 + q;
    float* xp = reinterpret_cast<float*>(smem + S2_XP) + set * 5 * TM;
    float* x2s = reinterpret_cast<float*>(smem + S2_X2) + set * TM;
    float* pmin = reinterpret_cast<float*>(smem + S2_PMIN) + set * 2 * TM;
    unsigned long long* rb = reinterpret_cast<unsigned long long*>(smem + S2_RB) + set * TM;
    uint16_t* wl = reinterpret_cast<uint16_t*>(smem + S2_WL) + sw_ * WL2_CAP;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + pw * (CB / 2);
    const uint8_t* a_tile = smem + SM_A + set * A_BYTES;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int who = sw_ == 0 ? 1 : (sw_ == 8 ? 2 : (sw_ == 4 ? 3 : -1));
#define VQ2_CLK(ev) do { if (clk && who > 0 && blockIdx.x == 0 && it < 24 && lane == 0) clk[(it * 4 + who) * 16 + (ev)] = (unsigned)clock(); } while (0)
    auto xptr = [&](int r) { return NHWC ? a_tile + r * 128 : a_tile + (r >> 5) * 8192 + (r & 7) * 4; };
    auto xelem = [&](const uint8_t* xr, int rsw, int j) {
      return NHWC ? *reinterpret_cast<const float*>(xr + (j >> 5) * 16384 + ((((j >> 2) & 7) ^ rsw) << 4) + (j & 3) * 4)
                  : *reinterpret_cast<const float*>(xr + j * 128 + ((rsw ^ (j & 3)) << 5));
    };
    int i = 0;
    for (int it = set;; it += 2, ++i) {
      const int tile = sub + it * ctas_per_group;
      if (tile >= num_tiles) break;
      const uint32_t pos = (uint32_t)tile * TM + m;
      const uint32_t frame = pos / (uint32_t)hw, s = pos - frame * (uint32_t)hw;
      // exact reference distance (vq.cu): sequential fp32 FMA chain over the 64 dims, d = fl(fl(c2 + x2) - 2*dot)
      auto exact_d = [&](int r, int k) {
        const float acc = exact_dot_smem<NHWC>(xptr(r), NHWC ? (r & 7) : ((r & 31) >> 3), smem, k);
        const float2 c2p = *reinterpret_cast<const float2*>(smem + SM_BX + k * 32 + (((k >> 2) & 1) << 4));
        return __fadd_rn(-2.f * acc, __fadd_rn(__fadd_rn(c2p.x, c2p.y), x2s[r]));  // hi + lo == c2 and -2 * acc are exact
      };
      VQ2_CLK(0);
      mbar_wait(&a_full[set], i & 1);
      VQ2_CLK(1);
      {
        // |x|^2 in ATen's order (vq.cu): lanes l = 0..7 left to right, L_l = ((a0+a1)+a2)+a3,
        // a_j = x[8j+l]^2 + x[8(j+4)+l]^2.  Warp 0 of the pair takes l = 0..3 (and sums them: a prefix of the
        // chain), warp 1 takes l = 4..7.
        const uint8_t* xr = xptr(m);
        const int rsw = NHWC ? (m & 7) : ((m & 31) >> 3);
        float L[4];
#pragma unroll
        for (int ll = 0; ll < 4; ++ll) {
          const int l = 4 * pw + ll;
          float a[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x0 = xelem(xr, rsw, 8 * j + l), x1 = xelem(xr, rsw, 8 * (j + 4) + l);
            a[j] = __fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1));
          }
          L[ll] = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[1]), a[2]), a[3]);
        }
        if (pw == 0) {
          xp[m] = __fadd_rn(__fadd_rn(__fadd_rn(L[0], L[1]), L[2]), L[3]);
        } else {
#pragma unroll
          for (int ll = 0; ll < 4; ++ll) xp[(1 + ll) * TM + m] = L[ll];
        }
      }
      pair_sync(pair_bar);  // R1; both warps are also past their reads of the previous tile's winners
      float x2 = xp[m];
#pragma unroll
      for (int ll = 1; ll < 5; ++ll) x2 = __fadd_rn(x2, xp[ll * TM + m]);
      if (pw == 0) {
        x2s[m] = x2;
        rb[m] = ~0ull;
      }
      const float twoE = 2.f * (1.1f * 0.00390625f * sqrtf(x2 * cmax2) + 3.8146973e-6f * (x2 + cmax2));
      VQ2_CLK(2);
      float rmin = INFINITY;
      uint32_t w[8];
      float sub_[8];
#pragma unroll
      for (int h = 0; h < NQ; ++h) {
        mbar_wait(&t_full[h * 2 + set], i & 1);
        tc_fence_after();
        if (h == 0) VQ2_CLK(3);
        if (h == NQ / 2) VQ2_CLK(5);
        uint32_t b0[32], b1[32];
        tmem_ld_32x32(t_row + h * CB, b0);
#pragma unroll
        for (int c = 0; c < CH; c += 2) {
          tmem_ld_wait();
          tmem_ld_32x32(t_row + h * CB + (c + 1) * 32, b1);
          w[h * CH + c] = scan_chunk<NF>(b0, rmin, twoE, sub_[h * CH + c]);
          tmem_ld_wait();
          if (c + 2 < CH) {
            tmem_ld_32x32(t_row + h * CB + (c + 2) * 32, b0);
          } else {  // the last chunk is in registers: buffer h may be overwritten
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[h]);
          }
          w[h * CH + c + 1] = scan_chunk<NF>(b1, rmin, twoE, sub_[h * CH + c + 1]);
        }
        if (h == NQ / 2 - 1) VQ2_CLK(4);
        if (h == NQ - 1) VQ2_CLK(6);
      }
      pmin[pw * TM + m] = rmin;
      pair_sync(pair_bar);  // R2
      VQ2_CLK(11);
      const float omin = pmin[(pw ^ 1) * TM + m];
      const float gthr = fminf(rmin, omin) + twoE;
      const bool partner_has = omin <= gthr;
      // drop the chunks whose minimum is outside the final window; up to three surviving codes go into registers
      // (any three: the re-rank is order-free), a lane with more goes through all of its codes itself
      int cnt = 0, k1 = 0, k2 = 0, k3 = 0;
      bool over = false;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t ww = (sub_[c] <= gthr) ? w[c] : 0u;
        w[c] = ww;
        cnt += __popc(ww);
        // branch-free: most words are empty, but some lane of the warp almost always has one that is not
        const int base = (c / CH) * CB + pw * (CB / 2) + (c % CH) * 32 - 1;
        const uint32_t ww2 = ww & (ww - 1);
        const int t1 = base + __ffs(ww), t2 = base + __ffs(ww2);
        const bool h1 = ww != 0, h2 = ww2 != 0;
        k3 = h1 ? k2 : k3; k2 = h1 ? k1 : k2; k1 = h1 ? t1 : k1;
        k3 = h2 ? k2 : k3; k2 = h2 ? k1 : k2; k1 = h2 ? t2 : k1;
        over |= (ww2 & (ww2 - 1)) != 0;
      }
      over |= cnt > 3;
      const bool need_exact = !over && (cnt >= 2 || (cnt >= 1 && partner_has));
      VQ2_CLK(14);
      if (cnt == 1 && !partner_has) rb[m] = (unsigned long long)k1;  // the only code inside the window IS the exact argmin
      {
        // per-warp list of (lane, code): ballots give every entry its slot, no atomics
        const uint32_t b1_ = __ballot_sync(0xffffffffu, need_exact);
        const uint32_t b2_ = __ballot_sync(0xffffffffu, need_exact && cnt >= 2);
        const uint32_t b3_ = __ballot_sync(0xffffffffu, need_exact && cnt >= 3);
        const int n1 = __popc(b1_), n2 = n1 + __popc(b2_), n = n2 + __popc(b3_);  // <= 96 = WL2_CAP
        if (need_exact) wl[__popc(b1_ & lt_mask)] = (uint16_t)((lane << 9) | k1);
        if (need_exact && cnt >= 2) wl[n1 + __popc(b2_ & lt_mask)] = (uint16_t)((lane << 9) | k2);
        if (need_exact && cnt >= 3) wl[n2 + __popc(b3_ & lt_mask)] = (uint16_t)((lane << 9) | k3);
        if (__ballot_sync(0xffffffffu, over)) {  // rows with many near-ties (adversarial codebooks)
          if (over) {
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
              uint32_t ww = c == 0 ? w[0] : c == 1 ? w[1] : c == 2 ? w[2] : c == 3 ? w[3] : c == 4 ? w[4] : c == 5 ? w[5] : c == 6 ? w[6] : w[7];
              const int base = (c / CH) * CB + pw * (CB / 2) + (c % CH) * 32 - 1;
              while (ww) {
                const int k = base + __ffs(ww);
                ww &= ww - 1;
                atomicMin(&rb[m], ((unsigned long long)ordered_f32(exact_d(m, k)) << 32) | (unsigned)k);
              }
            }
          }
        }
        int* wn = reinterpret_cast<int*>(smem + S2_WN);
        if (lane == 0) wn[sw_] = n;
        VQ2_CLK(7);
        pair_sync(pair_bar);  // R3: both lists of the pair are complete; the two warps share the re-rank evenly
        const int nA = wn[sw_ & ~4], nT = nA + wn[sw_ | 4];
        const uint16_t* wlA = reinterpret_cast<const uint16_t*>(smem + S2_WL) + (sw_ & ~4) * WL2_CAP;
        const uint16_t* wlB = reinterpret_cast<const uint16_t*>(smem + S2_WL) + (sw_ | 4) * WL2_CAP;
        VQ2_CLK(15);
        for (int e = pw * 32 + lane; e < nT; e += 64) {
          const uint32_t ent = e < nA ? wlA[e] : wlB[e - nA];
          const int r = q * 32 + (int)(ent >> 9), k = (int)(ent & 511u);
          atomicMin(&rb[r], ((unsigned long long)ordered_f32(exact_d(r, k)) << 32) | (unsigned)k);
        }
      }
      VQ2_CLK(8);
      if (!sums) {  // the re-rank was the last reader of the A tile (the MMAs completed before the scan)
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_empty[set]);
      }
      pair_sync(pair_bar);  // R4: winners final
      VQ2_CLK(9);
      const int besti = (int)(rb[m] & 511ull);
      if (pw == 0) idx_out[((size_t)frame * num + g) * hw + s] = (int64_t)besti;
      if (NHWC && (zq_out || zq_bf16)) {
        // z_q rows: each warp of the pair writes half of the quarter's rows; two rows per instruction, 16 lanes per
        // row, so whole 128 / 256-byte lines are read and written (one row per lane would touch 32 lines each time)
        const int hrow = lane >> 4, l16 = lane & 15;
#pragma unroll 4
        for (int rr = pw * 16; rr < pw * 16 + 16; rr += 2) {
          const int r = rr + hrow;
          const int kb = __shfl_sync(0xffffffffu, besti, r);
          const float4 v = __ldg(reinterpret_cast<const float4*>(cbg + (size_t)kb * TD) + l16);
          const size_t o = ((size_t)tile * TM + q * 32 + r) * C + (size_t)g * TD + 4 * l16;
          if (zq_out) *reinterpret_cast<float4*>(zq_out + o) = v;
          if (zq_bf16) {
            uint2 u;
            u.x = pack_bf16x2(v.x, v.y);
            u.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(zq_bf16 + o) = u;
          }
        }
      }
      if (!NHWC && zq_out) {  // NCHW: lanes are consecutive positions, each warp of the pair writes 32 of the 64 dims
        const float* cr = cbg + (size_t)besti * TD + pw * 32;
        float* zp = zq_out + ((size_t)frame * C + (size_t)g * TD + pw * 32) * hw + s;
#pragma unroll 8
        for (int j = 0; j < TD / 2; ++j) zp[(size_t)j * hw] = __ldg(cr + j);
      }
      if (pw == 0 && counts) atomicAdd(counts + (size_t)g * TK + besti, 1.f);
      if (sums) {  // each warp of the pair adds half of the row
        float* sp = sums + ((size_t)g * TK + besti) * TD;
        const uint8_t* xr = xptr(m);
        const int rsw = NHWC ? (m & 7) : ((m & 31) >> 3);
#pragma unroll 8
        for (int j = 0; j < TD / 2; ++j) atomicAdd(sp + pw * 32 + j, xelem(xr, rsw, pw * 32 + j));
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_empty[set]);
      }
      VQ2_CLK(10);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

}  // namespace

static float* g_dbg_scores = nullptr;
static unsigned* g_dbg_clk = nullptr;
extern "C" void lvt_dbg_vq_clock(unsigned* p) { g_dbg_clk = p; }  // debugging aid: per-tile timeline of CTA 0
extern "C" void lvt_dbg_vq_scores(float* p) { g_dbg_scores = p; }  // debugging aid: dump tile 0 / group 0 scores

// Returns LVT_OK after launching, or 1 when the shape is not covered (caller falls back to the SIMT kernel).
int lvt_vq_argmin_tc_try(const float* z_e, const float* codebook, int64_t* idx_out, float* zq_out, void* zq_bf16,
                         float* counts, float* sums, int n, int num, int K, int D, int hw, bool nhwc,
                         cudaStream_t stream) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("LVT_VQ_SIMT");
    disabled = (e && e[0] == '1') ? 1 : 0;
  }
  const long long positions = (long long)n * hw;
  if (disabled || K != TK || D != TD || n <= 0 || positions >= (1ll << 31)) return 1;
  if (nhwc ? (positions % TM != 0) : (hw % TM != 0 || zq_bf16 != nullptr)) return 1;
  if ((reinterpret_cast<uintptr_t>(z_e) & 15) != 0 || (reinterpret_cast<uintptr_t>(codebook) & 15) != 0) return 1;
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return 1;
  CUtensorMap tm;
  cuuint32_t estr[2] = {1, 1};
  CUresult r;
  if (nhwc) {
    cuuint64_t dims[2] = {(cuuint64_t)num * D, (cuuint64_t)positions};
    cuuint64_t strides[1] = {(cuuint64_t)num * D * 4};
    cuuint32_t box[2] = {32, TM};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(z_e), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)hw, (cuuint64_t)n * num * D};
    cuuint64_t strides[1] = {(cuuint64_t)hw * 4};
    cuuint32_t box[2] = {32, 64};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(z_e), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return 1;
  static int version = -1, nf = 16, pf = 1, nq = 2;
  if (version < 0) {
    const char* e = getenv("LVT_VQ_TC1");
    const char* f = getenv("LVT_VQ_NF");
    if (f) nf = atoi(f);
    const char* qe = getenv("LVT_VQ_NQ");
    if (qe) nq = atoi(qe);
    const char* pe = getenv("LVT_VQ_NOPF");
    if (pe && pe[0] == '1') pf = 0;
    version = (e && e[0] == '1') ? 1 : 2;
  }
  const int num_tiles = (int)(positions / TM);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int per_group = sms / num;
  if (per_group < 1) per_group = 1;
  if (per_group > num_tiles) per_group = num_tiles;
  __nv_bfloat16* zb = reinterpret_cast<__nv_bfloat16*>(zq_bf16);
  if (version == 1) {
    static bool configured = false;
    if (!configured) {
      LVT_CHECK_CUDA(cudaFuncSetAttribute(vq_argmin_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
      LVT_CHECK_CUDA(cudaFuncSetAttribute(vq_argmin_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
      configured = true;
    }
    if (nhwc)
      vq_argmin_tc_kernel<true><<<per_group * num, TC_THREADS, SM_TOTAL, stream>>>(
          tm, codebook, idx_out, zq_out, zb, counts, sums, num, hw, num_tiles, per_group, g_dbg_scores, g_dbg_clk);
    else
      vq_argmin_tc_kernel<false><<<per_group * num, TC_THREADS, SM_TOTAL, stream>>>(
          tm, codebook, idx_out, zq_out, nullptr, counts, sums, num, hw, num_tiles, per_group, g_dbg_scores, g_dbg_clk);
  } else {
#define VQ2_LAUNCH(L, F, Q)                                                                                            \
  do {                                                                                                                \
    static bool cfg = false;                                                                                          \
    if (!cfg) {                                                                                                       \
      LVT_CHECK_CUDA(cudaFuncSetAttribute(vq_argmin_tc2_kernel<L, F, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_TOTAL)); \
      cfg = true;                                                                                                     \
    }                                                                                                                 \
    vq_argmin_tc2_kernel<L, F, Q><<<per_group * num, TC_THREADS, S2_TOTAL, stream>>>(                                 \
        tm, z_e, codebook, idx_out, zq_out, L ? zb : nullptr, counts, sums, num, hw, num_tiles, per_group, pf, g_dbg_clk); \
  } while (0)
    if (nhwc) {
      if (nf == 32) VQ2_LAUNCH(true, 32, 2);
      else if (nq == 4) VQ2_LAUNCH(true, 16, 4);
      else VQ2_LAUNCH(true, 16, 2);
    } else {
      if (nf == 32) VQ2_LAUNCH(false, 32, 2);
      else if (nq == 4) VQ2_LAUNCH(false, 16, 4);
      else VQ2_LAUNCH(false, 16, 2);
    }
#undef VQ2_LAUNCH
  }
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}
