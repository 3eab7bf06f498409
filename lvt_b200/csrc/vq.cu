// lvt_b200 :: VQ codebook kernels (nearest-entry search, gather, EMA update).
//
// Reference semantics (vidgen/modeling/vq/vq_utils.py:7-24):
//     codebook_sqr = sum(codebook**2, dim=1); inputs_sqr = sum(x**2, dim=1, keepdim=True)
//     distances    = addmm(codebook_sqr + inputs_sqr, x, codebook.t(), alpha=-2, beta=1)   (fp32)
//     indices      = min(distances, dim=1)[1]                                      (first minimum)
// The integer result must be bit-exact, so the fp32 arithmetic is reproduced operation by
// operation (verified against ATen/MKL on CPU, see oracle/vq_oracle.c):
//   * dot(x, c_k)   : one sequential FMA chain over j = 0..D-1 starting from 0
//   * |v|^2         : products rounded individually, then ATen's vectorised row sum: vector i
//                     (8 lanes) is added into accumulator i % 4, accumulators are combined
//                     ((a0+a1)+a2)+a3 lane-wise, then the 8 lanes are summed left to right
//   * distance      : fl( fl(|c_k|^2 + |x|^2) - 2*dot )
// The kernel reads z_e in its native NCHW layout (no NHWC permute copy, no distance matrix in
// HBM; vq_embedding.py:25,36 + vq_utils.py:17-20 fused).
#include "../../include/lvt_b200.h"
#include <stdlib.h>
#include "common.cuh"

extern void lvt_count_launch(int n);
// tensor-core scan + exact re-rank (vq_tc.cu); returns 1 when the shape is not covered
int lvt_vq_argmin_tc_try(const float* z_e, const float* codebook, int64_t* idx_out, float* zq_out, void* zq_bf16,
                         float* counts, float* sums, int n, int num, int K, int D, int hw, bool nhwc,
                         cudaStream_t stream);

namespace {

// ATen-order squared norm of v[0..D) (D % 8 == 0); see header comment.
template <int D, typename LoadFn>
LVT_DEVICE_INLINE float sqnorm_aten_order(LoadFn ld) {
  constexpr int NV = D / 8;
  constexpr int NA = NV < 4 ? NV : 4;
  float acc[NA][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const float x = ld(i * 8 + l);
      const float sq = __fmul_rn(x, x);
      if (i < 4) acc[i][l] = sq;
      else acc[i % 4][l] = __fadd_rn(acc[i % 4][l], sq);
    }
  }
  float lanes[8];
#pragma unroll
  for (int l = 0; l < 8; ++l) {
    float a = acc[0][l];
#pragma unroll
    for (int j = 1; j < NA; ++j) a = __fadd_rn(a, acc[j][l]);
    lanes[l] = a;
  }
  float s = lanes[0];
#pragma unroll
  for (int l = 1; l < 8; ++l) s = __fadd_rn(s, lanes[l]);
  return s;
}

// Runtime-D variant (slow path).
LVT_DEVICE_INLINE float sqnorm_aten_order_rt(const float* v, long long stride, int D) {
  float acc[4][8];
  const int nv = D / 8;
  for (int i = 0; i < nv; ++i)
    for (int l = 0; l < 8; ++l) {
      const float x = v[(long long)(i * 8 + l) * stride];
      const float sq = __fmul_rn(x, x);
      if (i < 4) acc[i][l] = sq;
      else acc[i & 3][l] = __fadd_rn(acc[i & 3][l], sq);
    }
  const int na = nv < 4 ? nv : 4;
  float s = 0.f;
  for (int l = 0; l < 8; ++l) {
    float a = acc[0][l];
    for (int j = 1; j < na; ++j) a = __fadd_rn(a, acc[j][l]);
    s = (l == 0) ? a : __fadd_rn(s, a);
  }
  return s;
}

constexpr int VQ_THREADS = 512;

// Fast path: D == 64, codebook of group g resident in shared memory (K*D*4 <= 128 KiB).
// One thread = one (frame, position) vector of group g; x lives in registers.
template <int D>
__global__ void __launch_bounds__(VQ_THREADS, 1)
vq_argmin_smem_kernel(const float* __restrict__ z_e, const float* __restrict__ codebook,
                      int64_t* __restrict__ idx_out, float* __restrict__ zq_out,
                      __nv_bfloat16* __restrict__ zq_bf16, float* __restrict__ counts,
                      float* __restrict__ sums, long long total_pos, int num, int K, int hw,
                      long long pos_stride, long long ch_stride) {
  extern __shared__ float4 vq_smem4[];
  float* cb = reinterpret_cast<float*>(vq_smem4);  // [K][D]
  float* csq = cb + (size_t)K * D;                 // [K]
  const int g = blockIdx.y;
  const float* cbg = codebook + (size_t)g * K * D;

  for (int i = threadIdx.x; i < K * D / 4; i += VQ_THREADS)
    reinterpret_cast<float4*>(cb)[i] = __ldg(reinterpret_cast<const float4*>(cbg) + i);
  for (int k = threadIdx.x; k < K; k += VQ_THREADS) {
    const float* row = cbg + (size_t)k * D;
    csq[k] = sqnorm_aten_order<D>([&](int j) { return __ldg(row + j); });
  }
  __syncthreads();

  const long long p = (long long)blockIdx.x * VQ_THREADS + threadIdx.x;
  if (p >= total_pos) return;
  const long long frame = p / hw;
  const int s = (int)(p - frame * hw);
  const long long C = (long long)num * D;
  // NCHW: pos_stride 1, ch_stride hw;  NHWC: pos_stride C, ch_stride 1 (frame stride C*hw in both)
  const long long base = frame * C * hw + (long long)s * pos_stride + (long long)g * D * ch_stride;
  const float* xp = z_e + base;

  float x[D];
#pragma unroll
  for (int j = 0; j < D; ++j) x[j] = __ldg(xp + (long long)j * ch_stride);
  const float xsq = sqnorm_aten_order<D>([&](int j) { return x[j]; });

  float best = INFINITY;
  int besti = 0;
#pragma unroll 1
  for (int k = 0; k < K; k += 4) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const float4* c0 = reinterpret_cast<const float4*>(cb + (size_t)(k + 0) * D);
    const float4* c1 = reinterpret_cast<const float4*>(cb + (size_t)(k + 1) * D);
    const float4* c2 = reinterpret_cast<const float4*>(cb + (size_t)(k + 2) * D);
    const float4* c3 = reinterpret_cast<const float4*>(cb + (size_t)(k + 3) * D);
#pragma unroll
    for (int j = 0; j < D / 4; ++j) {
      const float4 a = c0[j], b = c1[j], c = c2[j], d = c3[j];
      acc0 = __fmaf_rn(x[4 * j], a.x, acc0); acc0 = __fmaf_rn(x[4 * j + 1], a.y, acc0);
      acc0 = __fmaf_rn(x[4 * j + 2], a.z, acc0); acc0 = __fmaf_rn(x[4 * j + 3], a.w, acc0);
      acc1 = __fmaf_rn(x[4 * j], b.x, acc1); acc1 = __fmaf_rn(x[4 * j + 1], b.y, acc1);
      acc1 = __fmaf_rn(x[4 * j + 2], b.z, acc1); acc1 = __fmaf_rn(x[4 * j + 3], b.w, acc1);
      acc2 = __fmaf_rn(x[4 * j], c.x, acc2); acc2 = __fmaf_rn(x[4 * j + 1], c.y, acc2);
      acc2 = __fmaf_rn(x[4 * j + 2], c.z, acc2); acc2 = __fmaf_rn(x[4 * j + 3], c.w, acc2);
      acc3 = __fmaf_rn(x[4 * j], d.x, acc3); acc3 = __fmaf_rn(x[4 * j + 1], d.y, acc3);
      acc3 = __fmaf_rn(x[4 * j + 2], d.z, acc3); acc3 = __fmaf_rn(x[4 * j + 3], d.w, acc3);
    }
    const float d0 = __fmaf_rn(-2.f, acc0, __fadd_rn(csq[k + 0], xsq));
    const float d1 = __fmaf_rn(-2.f, acc1, __fadd_rn(csq[k + 1], xsq));
    const float d2 = __fmaf_rn(-2.f, acc2, __fadd_rn(csq[k + 2], xsq));
    const float d3 = __fmaf_rn(-2.f, acc3, __fadd_rn(csq[k + 3], xsq));
    if (d0 < best) { best = d0; besti = k; }
    if (d1 < best) { best = d1; besti = k + 1; }
    if (d2 < best) { best = d2; besti = k + 2; }
    if (d3 < best) { best = d3; besti = k + 3; }
  }

  idx_out[(frame * num + g) * hw + s] = (int64_t)besti;
  if (zq_out) {
    float* zp = zq_out + base;
    const float* cr = cb + (size_t)besti * D;
#pragma unroll
    for (int j = 0; j < D; ++j) zp[(long long)j * ch_stride] = cr[j];
  }
  if (zq_bf16) {
    __nv_bfloat16* zp = zq_bf16 + base;
    const float* cr = cb + (size_t)besti * D;
#pragma unroll
    for (int j = 0; j < D; ++j) zp[(long long)j * ch_stride] = __float2bfloat16(cr[j]);
  }
  if (counts) atomicAdd(counts + (size_t)g * K + besti, 1.f);
  if (sums) {
    float* sp = sums + ((size_t)g * K + besti) * D;
#pragma unroll
    for (int j = 0; j < D; ++j) atomicAdd(sp + j, x[j]);
  }
}

// Generic path (any D % 8 == 0, any K): codebook streamed from L2. Correctness path for the
// non-DVQ configurations (CODEBOOK.NUM == 1, D == 256: configs/vqvae/Base-VQVAE.yaml); not tuned.
// |c_k|^2 of the group is computed by the block into shared memory (K floats); z_e / zq are addressed with a
// position stride and a channel stride, so NCHW (1, hw) and channels-last (num*D, 1) share the kernel.
__global__ void vq_argmin_generic_kernel(const float* __restrict__ z_e,
                                         const float* __restrict__ codebook,
                                         int64_t* __restrict__ idx_out, float* __restrict__ zq_out,
                                         __nv_bfloat16* __restrict__ zq_bf16,
                                         float* __restrict__ counts, float* __restrict__ sums,
                                         long long total_pos, int num, int K, int D, int hw,
                                         long long pos_stride, long long ch_stride) {
  extern __shared__ float csq_s[];
  const int g = blockIdx.y;
  const float* cbg = codebook + (size_t)g * K * D;
  for (int k = threadIdx.x; k < K; k += blockDim.x) csq_s[k] = sqnorm_aten_order_rt(cbg + (size_t)k * D, 1, D);
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total_pos) return;
  const long long frame = p / hw;
  const int s = (int)(p - frame * hw);
  const long long base = frame * (long long)num * D * hw + (long long)s * pos_stride + (long long)g * D * ch_stride;
  const float* xp = z_e + base;
  const float xsq = sqnorm_aten_order_rt(xp, ch_stride, D);
  float best = INFINITY;
  int besti = 0;
  for (int k = 0; k < K; ++k) {
    const float* cr = cbg + (size_t)k * D;
    float acc = 0.f;
    for (int j = 0; j < D; ++j) acc = __fmaf_rn(xp[(long long)j * ch_stride], __ldg(cr + j), acc);
    const float d = __fmaf_rn(-2.f, acc, __fadd_rn(csq_s[k], xsq));
    if (d < best) { best = d; besti = k; }
  }
  idx_out[(frame * num + g) * hw + s] = (int64_t)besti;
  const float* cr = cbg + (size_t)besti * D;
  if (zq_out || zq_bf16) {
    for (int j = 0; j < D; ++j) {
      const float v = cr[j];
      if (zq_out) zq_out[base + (long long)j * ch_stride] = v;
      if (zq_bf16) zq_bf16[base + (long long)j * ch_stride] = __float2bfloat16(v);
    }
  }
  if (counts) atomicAdd(counts + (size_t)g * K + besti, 1.f);
  if (sums) {
    float* sp = sums + ((size_t)g * K + besti) * D;
    for (int j = 0; j < D; ++j) atomicAdd(sp + j, xp[(long long)j * ch_stride]);
  }
}

__global__ void vq_csq_kernel(const float* __restrict__ codebook, float* __restrict__ csq, int rows,
                              int D) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < rows) csq[k] = sqnorm_aten_order_rt(codebook + (size_t)k * D, 1, D);
}

__global__ void vq_gather_kernel(const int64_t* __restrict__ idx, const float* __restrict__ codebook,
                                 float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16,
                                 long long total_pos, int num, int K, int D, int hw, long long pos_stride,
                                 long long ch_stride) {
  const int g = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total_pos) return;
  const long long frame = p / hw;
  const int s = (int)(p - frame * hw);
  long long code = idx[(frame * num + g) * hw + s];
  code = code < 0 ? 0 : (code >= K ? K - 1 : code);
  const float* cr = codebook + ((size_t)g * K + code) * D;
  const long long base = frame * (long long)num * D * hw + (long long)s * pos_stride + (long long)g * D * ch_stride;
  for (int j = 0; j < D; ++j) {
    const float v = __ldg(cr + j);
    if (out) out[base + (long long)j * ch_stride] = v;
    if (out_bf16) out_bf16[base + (long long)j * ch_stride] = __float2bfloat16(v);
  }
}

// channels-last variant: 16 threads per position, one float4 of the code row each -> 256 B contiguous per position
__global__ void __launch_bounds__(256)
vq_gather_nhwc_kernel(const int64_t* __restrict__ idx, const float* __restrict__ codebook, float* __restrict__ out,
                      __nv_bfloat16* __restrict__ out_bf16, long long total_pos, int num, int K, int hw) {
  const int g = blockIdx.y;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long p = t >> 4;
  if (p >= total_pos) return;
  const int j4 = (int)(t & 15);
  const long long frame = p / hw;
  const int s = (int)(p - frame * hw);
  long long code = idx[(frame * num + g) * hw + s];
  code = code < 0 ? 0 : (code >= K ? K - 1 : code);
  const float4 v = __ldg(reinterpret_cast<const float4*>(codebook + ((size_t)g * K + code) * 64) + j4);
  const long long o = (p * num + g) * 64 + 4 * j4;
  if (out) *reinterpret_cast<float4*>(out + o) = v;
  if (out_bf16) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out_bf16 + o) = u;
  }
}

// EMA update (vq_embedding.py:48-59) in two launches: (1) one block per codebook group updates running_size in place
// and reduces n = sum(running_size) in a fixed order (deterministic), (2) an elementwise grid over the K x D entries
// updates running_sum and the codebook.  (One block per group doing both walked its 32 768 entries 128 deep with two
// divisions each: 82 us for 1 MB of state.)
__global__ void vq_ema_size_kernel(float* __restrict__ running_size, const float* __restrict__ counts,
                                   float* __restrict__ n_out, int K, float decay, float omd) {
  extern __shared__ float red[];
  const int g = blockIdx.x;
  float* rs = running_size + (size_t)g * K;
  const float* cnt = counts + (size_t)g * K;
  float part = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    // running_size.mul_(decay).add_(1 - decay, size): two rounded ops, as in ATen
    const float v = __fadd_rn(__fmul_rn(rs[k], decay), __fmul_rn(omd, cnt[k]));
    rs[k] = v;
    part += v;
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) n_out[g] = red[0];
}

__global__ void __launch_bounds__(256)
vq_ema_apply_kernel(float* __restrict__ codebook, const float* __restrict__ running_size,
                    float* __restrict__ running_sum, const float* __restrict__ sums, const float* __restrict__ n_in,
                    int K, int D, float decay, float omd, float eps, float k_eps) {
  const int g = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * D) return;
  const float n = n_in[g];
  const float denom = n + k_eps;
  const size_t o = (size_t)g * K * D + i;
  const float v = __fadd_rn(__fmul_rn(running_sum[o], decay), __fmul_rn(omd, sums[o]));
  running_sum[o] = v;
  const float size_ = (running_size[(size_t)g * K + i / D] + eps) / denom * n;
  codebook[o] = v / size_;
}


// EMA statistics of the codebook update (vq_embedding.py:44-56): counts[g,k] += #{pos : idx = k},
// sums[g,k,:] += sum of the z_e rows assigned to code k.  One CTA = one codebook group and a contiguous chunk
// of positions; the K x 64 sums are privatised in shared memory (conflict-free: lane l owns dims 2l, 2l+1), so
// global memory sees one vectorised red.add per touched row and CTA instead of 64 scalar atomics per position.
__global__ void __launch_bounds__(1024)
vq_ema_stats_kernel(const float* __restrict__ z_e, const int64_t* __restrict__ idx, float* __restrict__ counts,
                    float* __restrict__ sums, long long total, int num, int K, int hw, long long pos_stride,
                    long long ch_stride, int chunk) {
  extern __shared__ float s_acc[];  // [K][64] sums, then [K] counts
  float* s_cnt = s_acc + (size_t)K * 64;
  const int g = blockIdx.y;
  const int nw = blockDim.x >> 5;  // warps per block (few, large blocks: every block flushes K x 64 sums at the end)
  for (int i = threadIdx.x; i < K * 65; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long p0 = (long long)blockIdx.x * chunk;
  const long long p1 = p0 + chunk < total ? p0 + chunk : total;
  constexpr int U = 4;  // positions in flight per warp: index and row loads of all U issue before the first atomic
  for (long long pb = p0 + warp; pb < p1; pb += (long long)nw * U) {
    int kk[U];
    float x0[U], x1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pos = pb + (long long)nw * u;
      kk[u] = -1;
      if (pos < p1) {
        const long long frame = pos / hw, sp = pos - frame * hw;
        kk[u] = (int)idx[(frame * num + g) * hw + sp];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pos = pb + (long long)nw * u;
      x0[u] = x1[u] = 0.f;
      if (kk[u] >= 0) {
        const long long frame = pos / hw, sp = pos - frame * hw;
        const float* x = z_e + (ch_stride == 1 ? pos * pos_stride + (long long)g * 64
                                               : (frame * num * 64 + (long long)g * 64) * hw + sp);
        x0[u] = x[(2 * lane) * ch_stride];
        x1[u] = x[(2 * lane + 1) * ch_stride];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (kk[u] < 0) continue;  // warp-uniform
      atomicAdd(&s_acc[kk[u] * 64 + 2 * lane], x0[u]);
      atomicAdd(&s_acc[kk[u] * 64 + 2 * lane + 1], x1[u]);
      if (lane == 0) atomicAdd(&s_cnt[kk[u]], 1.f);
    }
  }
  __syncthreads();
  for (int k = warp; k < K; k += nw) {
    const float c = s_cnt[k];
    if (c == 0.f) continue;  // warp-uniform
    if (lane == 0 && counts) atomicAdd(counts + (size_t)g * K + k, c);
    if (sums) {
      float* dst = sums + ((size_t)g * K + k) * 64 + 2 * lane;
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(s_acc[k * 64 + 2 * lane]),
                   "f"(s_acc[k * 64 + 2 * lane + 1])
                   : "memory");
    }
  }
}

// Codebook gradient of the non-EMA objective (MODEL.CODEBOOK.EMA False): loss = mse(z_q, sg[z_e]) reaches the
// codebook through index_select (vq_embedding.py:61-64, vqvae.py:84-85); d/de_k = (2 / numel) * sum over the positions
// assigned to k of (e_k - z_e) = scale * (counts_k * e_k - sums_k) -- the statistics the EMA path already produces.
__global__ void __launch_bounds__(256)
vq_codebook_grad_kernel(const float* __restrict__ counts, const float* __restrict__ sums,
                        const float* __restrict__ codebook, float* __restrict__ grad, float scale, int rows, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  grad[i] = scale * (counts[i / D] * codebook[i] - sums[i]);
}

}  // namespace

static int vq_argmin_impl(const float* z_e, const float* codebook, int64_t* idx_out, float* zq_out,
                          void* zq_bf16, float* counts, float* sums, int n, int num, int K, int D, int hw,
                          bool nhwc, cudaStream_t stream) {
  LVT_CHECK_ARG(n >= 0 && num > 0 && K > 0 && D > 0 && hw > 0, "lvt_vq_argmin: bad shape");
  LVT_CHECK_ARG(D % 8 == 0, "lvt_vq_argmin: D must be a multiple of 8 (got %d)", D);
  if (n == 0) return LVT_OK;
  LVT_CHECK_ARG(z_e && codebook && idx_out, "lvt_vq_argmin: null pointer");
  {
    // EMA statistics: privatised in shared memory by a second small kernel (per-position global atomics inside the
    // search kernel cost 3.9 ms per 131 k positions) whenever K x 64 floats fit
    const bool split_stats = (counts || sums) && D == 64 && (size_t)K * 65 * 4 <= 200 * 1024;
    const int rc = lvt_vq_argmin_tc_try(z_e, codebook, idx_out, zq_out, zq_bf16, split_stats ? nullptr : counts,
                                        split_stats ? nullptr : sums, n, num, K, D, hw, nhwc, stream);
    if (rc < 0) return rc;
    if (rc == 0) {
      if (split_stats) {
        const long long total = (long long)n * hw;
        const size_t smem = (size_t)K * 65 * 4;
        static bool configured = false;
        if (!configured) {
          LVT_CHECK_CUDA(cudaFuncSetAttribute(vq_ema_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          configured = true;
        }
        static int stats_blocks = 0;
        if (!stats_blocks) {
          const char* e = getenv("LVT_VQ_STATS_BLOCKS");
          stats_blocks = e ? atoi(e) : 148;
          if (stats_blocks < 1) stats_blocks = 148;
        }
        // one 1024-thread block per SM (133 KB of privatised sums each): 93 us for 131 072 positions against 182 us
        // with 74 blocks -- the shared-memory float atomics (CAS loops), not the flushes, bound this kernel
        int per_group = stats_blocks / num > 0 ? stats_blocks / num : 1;
        int chunk = (int)((total + per_group - 1) / per_group);
        if (chunk < 256) chunk = 256;
        dim3 grid(lvt_ceil_div(total, chunk), num);
        vq_ema_stats_kernel<<<grid, 1024, smem, stream>>>(z_e, idx_out, counts, sums, total, num, K, hw,
                                                         nhwc ? (long long)num * D : 1, nhwc ? 1 : hw, chunk);
        LVT_CHECK_LAUNCH();
        lvt_count_launch(1);
      }
      return LVT_OK;  // 1 = shape not covered by the tensor-core kernel -> SIMT kernel below
    }
  }
  const long long total = (long long)n * hw;
  const long long pos_stride = nhwc ? (long long)num * D : 1, ch_stride = nhwc ? 1 : hw;
  if (D == 64 && K % 4 == 0 && (size_t)K * D * 4 + K * 4 <= 200 * 1024) {
    const size_t smem = (size_t)K * D * 4 + (size_t)K * 4;
    static bool configured = false;
    if (!configured) {
      LVT_CHECK_CUDA(cudaFuncSetAttribute(vq_argmin_smem_kernel<64>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
    dim3 grid(lvt_ceil_div(total, VQ_THREADS), num);
    vq_argmin_smem_kernel<64><<<grid, VQ_THREADS, smem, stream>>>(
        z_e, codebook, idx_out, zq_out, reinterpret_cast<__nv_bfloat16*>(zq_bf16), counts, sums, total, num, K, hw,
        pos_stride, ch_stride);
    LVT_CHECK_LAUNCH();
    lvt_count_launch(1);
    return LVT_OK;
  }
  LVT_CHECK_ARG((size_t)K * 4 <= 48 * 1024, "lvt_vq_argmin: the generic (D != 64) path keeps |c|^2 of %d codes in shared memory", K);
  dim3 grid(lvt_ceil_div(total, 128), num);
  vq_argmin_generic_kernel<<<grid, 128, (size_t)K * 4, stream>>>(z_e, codebook, idx_out, zq_out,
                                                                 reinterpret_cast<__nv_bfloat16*>(zq_bf16), counts, sums, total,
                                                                 num, K, D, hw, pos_stride, ch_stride);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vq_argmin(const float* z_e, const float* codebook, int64_t* idx_out,
                             float* zq_out, float* counts, float* sums, int n, int num, int K, int D,
                             int hw, void* stream_) {
  return vq_argmin_impl(z_e, codebook, idx_out, zq_out, nullptr, counts, sums, n, num, K, D, hw, false,
                        reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int lvt_vq_argmin_nhwc(const float* z_e, const float* codebook, int64_t* idx_out, float* zq_out,
                                  void* zq_bf16, float* counts, float* sums, int n, int num, int K, int D,
                                  int hw, void* stream_) {
  return vq_argmin_impl(z_e, codebook, idx_out, zq_out, zq_bf16, counts, sums, n, num, K, D, hw, true,
                        reinterpret_cast<cudaStream_t>(stream_));
}

static int vq_gather_impl(const int64_t* idx, const float* codebook, float* out, void* out_bf16, int n, int num,
                          int K, int D, int hw, bool nhwc, cudaStream_t stream) {
  LVT_CHECK_ARG(n >= 0 && num > 0 && K > 0 && D > 0 && hw > 0, "lvt_vq_gather: bad shape");
  if (n == 0) return LVT_OK;
  LVT_CHECK_ARG(idx && codebook && (out || out_bf16), "lvt_vq_gather: null pointer");
  const long long total = (long long)n * hw;
  if (nhwc && D == 64) {
    dim3 grid16(lvt_ceil_div(total * 16, 256), num);
    vq_gather_nhwc_kernel<<<grid16, 256, 0, stream>>>(idx, codebook, out, reinterpret_cast<__nv_bfloat16*>(out_bf16), total,
                                                      num, K, hw);
    LVT_CHECK_LAUNCH();
    lvt_count_launch(1);
    return LVT_OK;
  }
  dim3 grid(lvt_ceil_div(total, 256), num);
  vq_gather_kernel<<<grid, 256, 0, stream>>>(idx, codebook, out, reinterpret_cast<__nv_bfloat16*>(out_bf16), total,
                                             num, K, D, hw, nhwc ? (long long)num * D : 1, nhwc ? 1 : hw);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vq_gather(const int64_t* idx, const float* codebook, float* out, int n, int num,
                             int K, int D, int hw, void* stream_) {
  return vq_gather_impl(idx, codebook, out, nullptr, n, num, K, D, hw, false, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int lvt_vq_gather_nhwc(const int64_t* idx, const float* codebook, float* out, void* out_bf16, int n,
                                  int num, int K, int D, int hw, void* stream_) {
  return vq_gather_impl(idx, codebook, out, out_bf16, n, num, K, D, hw, true, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int lvt_vq_ema_update(float* codebook, float* running_size, float* running_sum,
                                 const float* counts, const float* sums, int num, int K, int D,
                                 double decay, double eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LVT_CHECK_ARG(num > 0 && K > 0 && D > 0, "lvt_vq_ema_update: bad shape");
  LVT_CHECK_ARG(codebook && running_size && running_sum && counts && sums, "lvt_vq_ema_update: null pointer");
  const int threads = 256;
  LVT_CHECK_ARG(num <= 64, "lvt_vq_ema_update: at most 64 codebook groups (%d)", num);
  static float* n_scratch = nullptr;  // [64] sum(running_size) per group, between the two launches
  if (!n_scratch) LVT_CHECK_CUDA(cudaMalloc(&n_scratch, 64 * sizeof(float)));
  vq_ema_size_kernel<<<num, threads, threads * sizeof(float), stream>>>(running_size, counts, n_scratch, K, (float)decay,
                                                                        (float)(1.0 - decay));
  LVT_CHECK_LAUNCH();
  vq_ema_apply_kernel<<<dim3(lvt_ceil_div(K * D, threads), num), threads, 0, stream>>>(
      codebook, running_size, running_sum, sums, n_scratch, K, D, (float)decay, (float)(1.0 - decay), (float)eps,
      (float)(K * eps));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(2);
  return LVT_OK;
}

extern "C" int lvt_vq_codebook_grad(const float* counts, const float* sums, const float* codebook, float* grad,
                                    float scale, int rows, int D, void* stream_) {
  LVT_CHECK_ARG(counts && sums && codebook && grad && rows > 0 && D > 0, "lvt_vq_codebook_grad: bad argument");
  vq_codebook_grad_kernel<<<lvt_ceil_div(rows * D, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      counts, sums, codebook, grad, scale, rows, D);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}
