// lvt_b200 :: VQ codebook search on the tensor cores with an EXACT fp32 re-rank.
//
// The reference result (vq_utils.py:7-24) is argmin_k of d_k = fl(fl(c2_k + x2) - 2*dot_k) with dot_k a
// sequential fp32 FMA chain (see vq.cu).  262 144 FLOP per latent position make the SIMT kernel
// FFMA-bound (1.8 % of the HBM roofline), so the 512-way scan is moved to tcgen05:
//   1. scores  s_k = c2_k - 2 * <tf32(x), tf32(c_k)>  for all 512 codes of a group: kind::tf32 MMA,
//      A = 128 positions x 64 dims read straight from the NCHW tensor (MN-major fp32 tile via TMA, no
//      conversion pass), B = -2*codebook resident in shared memory, accumulator PRE-LOADED with the
//      exact fp32 c2_k (tcgen05.st), so the epilogue never touches c2.
//   2. per position, the candidate set {k : s_k <= min_k s_k + 2E} with a rigorous error bound
//      E = 1.1 * 2^-8 * |x| * max_k|c_k| + 2^-18 * (|x|^2 + max|c|^2)   (tf32 truncation of both operands,
//      Cauchy-Schwarz; covers the tensor-core accumulation and the reference's own fp32 roundings).
//   3. the reference's exact arithmetic on the candidates only (typically 1-2 of 512), lowest index
//      wins ties  =>  bit-identical indices to the SIMT kernel / the oracle.
// One CTA = one codebook group (its 128 KiB of -2*c stay resident), persistent over 128-position tiles;
// warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 scan + re-rank TMEM buffer 0 (codes 0-255), warps 6-9 buffer 1 (codes 256-511).
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
constexpr int A_BYTES = TM * TD * 4;        // 32 KiB
constexpr int B_BYTES = 2 * TK * 128;       // two k-blocks of 32 fp32 (128 B rows)
constexpr int SM_A = B_BYTES;
constexpr int SM_C2 = SM_A + NSLOT * A_BYTES;
constexpr int LST_CAP = 16;                 // candidates kept per (position, half); more => exhaustive exact scan
constexpr int SM_LST = SM_C2 + TK * 4;      // [LST_CAP][256] u16 candidate lists
constexpr int SM_XMIN = SM_LST + LST_CAP * 256 * 2;  // [2][128] f32 per-half minima
constexpr int SM_XBEST = SM_XMIN + 2 * TM * 4;       // [2][128] (d, idx) per-half winners
constexpr int SM_BAR = SM_XBEST + 2 * TM * 8;
constexpr int SM_TOTAL = SM_BAR + 256 + 1024;
constexpr int TC_THREADS = 320;             // TMA warp, MMA warp, 2 x 4 scan warps

LVT_DEVICE_INLINE void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
LVT_DEVICE_INLINE void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
LVT_DEVICE_INLINE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
LVT_DEVICE_INLINE void epi_bar(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
LVT_DEVICE_INLINE float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

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

// Shared-memory matrix descriptor with an explicit layout type (1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B).
LVT_DEVICE_INLINE uint64_t smem_desc_lt(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
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
                    float* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* c2s = reinterpret_cast<float*>(smem + SM_C2);
  uint16_t* lst = reinterpret_cast<uint16_t*>(smem + SM_LST);
  float* xmin = reinterpret_cast<float*>(smem + SM_XMIN);
  float2* xbest = reinterpret_cast<float2*>(smem + SM_XBEST);
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

  // ---- one-time setup: B = -2 * codebook[g] in the K-major 128B-swizzled UMMA layout, exact c2 table
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
  }
  cm = warp_max(cm);
  if (threadIdx.x == 0) {
    *cmax2_s = 0.f;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 8);
    }
    mbar_init(&t_full[0], 1);
    mbar_init(&t_full[1], 1);
    mbar_init(&t_empty[0], 4);
    mbar_init(&t_empty[1], 4);
    fence_barrier_init();
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  __syncthreads();
  if (lane == 0) atomicMax(reinterpret_cast<int*>(cmax2_s), __float_as_int(cm));  // cm >= 0: int order == float order
  fence_proxy_async();  // B written through the generic proxy, read by tcgen05.mma (async proxy)
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
      // idesc: D f32, A/B tf32, A MN-major, B K-major, N = 256, M = 128
      constexpr uint32_t idesc = umma_idesc(TM, 256, /*tf32*/ 2, !NHWC, false);
      const uint32_t b_base = smem_u32(smem);
      int it = 0;
      for (int tile = sub; tile < num_tiles; tile += ctas_per_group, ++it) {
        const int slot = it % NSLOT;
        mbar_wait(&a_full[slot], (it / NSLOT) & 1);
        const uint32_t a_base = smem_u32(smem + SM_A + slot * A_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&t_empty[h], it & 1);  // scan of the previous tile done AND c2 pre-loaded
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // K = 8 per tf32 MMA
            // K-major: 32 B per K-step inside the 128 B row, next k-block after 4 steps.
            // MN-major (BASE32B): atoms of 4 k-rows x 128 B (SBO 512), 32-position atoms 8 KiB apart (LBO).
            const uint64_t adesc = NHWC ? smem_desc_lt(a_base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, 2)
                                        : smem_desc_lt(a_base + ks * 1024, 8192, 512, 1);
            const uint64_t bdesc = umma_smem_desc(b_base + (ks >> 2) * (TK * 128) + h * (256 * 128) + (ks & 3) * 32, 16, 1024);
            umma_tf32_ss(tmem_base + h * 256, adesc, bdesc, idesc, 1u);  // accumulate onto c2
          }
          umma_commit(&t_full[h]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ scan + exact re-rank
    // Two groups of four warps; group h owns TMEM buffer h (codes h*256 .. h*256+255) of every tile.
    const int h = (warp - 2) >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;  // row (position) inside the tile
    const int et = h * TM + m;    // epilogue thread id 0..255
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * 256;
    auto init_c2 = [&]() {
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        const float4* src = reinterpret_cast<const float4*>(c2s + h * 256 + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = src[i];
          r[4 * i] = __float_as_uint(v.x); r[4 * i + 1] = __float_as_uint(v.y);
          r[4 * i + 2] = __float_as_uint(v.z); r[4 * i + 3] = __float_as_uint(v.w);
        }
        tmem_st_32x32(t_addr + c * 32, r);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[h]);
    };
    // exact reference distance of code k (sequential fp32 FMA chain over the 64 dims, vq.cu)
    auto exact_d = [&](const float (&x)[TD], float x2, int k) {
      const uint8_t* brow = smem + k * 128;
      float acc = 0.f;
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const float4 v = *reinterpret_cast<const float4*>(brow + (jj >> 3) * (TK * 128) + (((jj & 7) ^ (k & 7)) << 4));
        acc = __fmaf_rn(x[4 * jj], -0.5f * v.x, acc);
        acc = __fmaf_rn(x[4 * jj + 1], -0.5f * v.y, acc);
        acc = __fmaf_rn(x[4 * jj + 2], -0.5f * v.z, acc);
        acc = __fmaf_rn(x[4 * jj + 3], -0.5f * v.w, acc);
      }
      return __fmaf_rn(-2.f, acc, __fadd_rn(c2s[k], x2));
    };
    init_c2();
    int it = 0;
    for (int tile = sub; tile < num_tiles; tile += ctas_per_group, ++it) {
      const int slot = it % NSLOT;
      const long long pos = (long long)tile * TM + m;  // global position index (frame * hw + s)
      const int frame = (int)(pos / hw), s = (int)(pos - (long long)frame * hw);
      mbar_wait(&a_full[slot], (it / NSLOT) & 1);
      // this position's 64-dim vector from the swizzled tile
      const uint8_t* at = NHWC ? smem + SM_A + slot * A_BYTES + m * 128
                               : smem + SM_A + slot * A_BYTES + (m >> 5) * 8192 + (m & 7) * 4;
      const int mchunk = (m & 31) >> 3;  // NCHW: 32 B chunk of the 128 B row
      auto load_x = [&](float (&x)[TD]) {
        if (NHWC) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const float4 v = *reinterpret_cast<const float4*>(at + (jj >> 3) * 16384 + (((jj & 7) ^ (m & 7)) << 4));
            x[4 * jj] = v.x; x[4 * jj + 1] = v.y; x[4 * jj + 2] = v.z; x[4 * jj + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < TD; ++j) x[j] = *reinterpret_cast<const float*>(at + j * 128 + ((mchunk ^ (j & 3)) << 5));
        }
      };
      float x2;
      {
        float x[TD];
        load_x(x);
        x2 = sqnorm64([&](int j) { return x[j]; });
      }
      const float E = 1.1f * 0.00390625f * sqrtf(x2 * cmax2) + 3.8146973e-6f * (x2 + cmax2);
      const float W = 2.f * E;
      mbar_wait(&t_full[h], it & 1);
      tc_fence_after();
      // pass 1: minimum tf32 score of this half (two TMEM loads in flight per wait)
      {
        float m0 = INFINITY, m1 = INFINITY;
#pragma unroll 1
        for (int c = 0; c < 8; c += 2) {
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
        xmin[et] = fminf(m0, m1);
      }
      epi_bar(1);
      const float gthr = fminf(xmin[m], xmin[TM + m]) + W;
      // pass 2: candidate list {k : s_k <= min + 2E} (increasing k), kept in shared memory
      int cnt = 0;
#pragma unroll 1
      for (int c = 0; c < 8; c += 2) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_addr + c * 32, r0);
        tmem_ld_32x32(t_addr + c * 32 + 32, r1);
        tmem_ld_wait();
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          w0 |= (__uint_as_float(r0[i]) <= gthr) ? (1u << i) : 0u;
          w1 |= (__uint_as_float(r1[i]) <= gthr) ? (1u << i) : 0u;
        }
        if (dbg && tile == 0 && g == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            dbg[(size_t)m * TK + h * 256 + c * 32 + i] = __uint_as_float(r0[i]);
            dbg[(size_t)m * TK + h * 256 + c * 32 + 32 + i] = __uint_as_float(r1[i]);
          }
        }
        while (w0) {
          const int b = __ffs(w0) - 1;
          w0 &= w0 - 1;
          if (cnt < LST_CAP) lst[cnt * 256 + et] = (uint16_t)(h * 256 + c * 32 + b);
          ++cnt;
        }
        while (w1) {
          const int b = __ffs(w1) - 1;
          w1 &= w1 - 1;
          if (cnt < LST_CAP) lst[cnt * 256 + et] = (uint16_t)(h * 256 + c * 32 + 32 + b);
          ++cnt;
        }
      }
      init_c2();  // buffer h is free again: pre-load c2 for the next tile and hand it to the MMA warp
      float x[TD];
      load_x(x);
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_empty[slot]);  // this half's MMAs retired (t_full) and x is in registers
      // exact reference arithmetic on the candidates, increasing code index, strict < (first minimum)
      float best = INFINITY;
      int besti = 0;
      if (cnt <= LST_CAP) {
        for (int i = 0; i < cnt; ++i) {
          const int k = lst[i * 256 + et];
          const float d = exact_d(x, x2, k);
          if (d < best) { best = d; besti = k; }
        }
      } else {  // pathological codebook (dozens of near-ties): every code of this half, exactly
        for (int k = h * 256; k < h * 256 + 256; ++k) {
          const float d = exact_d(x, x2, k);
          if (d < best) { best = d; besti = k; }
        }
      }
      xbest[et] = make_float2(best, __int_as_float(besti));
      epi_bar(2);
      {
        const float2 o = xbest[(h ^ 1) * TM + m];
        const float od = o.x;
        const int oi = __float_as_int(o.y);
        // lower half wins exact ties (first minimum)
        if (h == 0 ? (od < best) : !(best < od)) { best = od; besti = oi; }
      }
      if (h == 0) {
        idx_out[((size_t)frame * num + g) * hw + s] = (int64_t)besti;
        if (counts) atomicAdd(counts + (size_t)g * TK + besti, 1.f);
        if (sums) {
          float* sp = sums + ((size_t)g * TK + besti) * TD;
#pragma unroll
          for (int j = 0; j < TD; ++j) atomicAdd(sp + j, x[j]);
        }
      } else if (zq_out || zq_bf16) {
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
  if (disabled || K != TK || D != TD || n <= 0) return 1;
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
  static bool configured = false;
  if (!configured) {
    LVT_CHECK_CUDA(cudaFuncSetAttribute(vq_argmin_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    LVT_CHECK_CUDA(cudaFuncSetAttribute(vq_argmin_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    configured = true;
  }
  const int num_tiles = (int)(positions / TM);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int per_group = sms / num;
  if (per_group < 1) per_group = 1;
  if (per_group > num_tiles) per_group = num_tiles;
  if (nhwc)
    vq_argmin_tc_kernel<true><<<per_group * num, TC_THREADS, SM_TOTAL, stream>>>(
        tm, codebook, idx_out, zq_out, reinterpret_cast<__nv_bfloat16*>(zq_bf16), counts, sums, num, hw, num_tiles,
        per_group, g_dbg_scores);
  else
    vq_argmin_tc_kernel<false><<<per_group * num, TC_THREADS, SM_TOTAL, stream>>>(
        tm, codebook, idx_out, zq_out, nullptr, counts, sums, num, hw, num_tiles, per_group, g_dbg_scores);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}
