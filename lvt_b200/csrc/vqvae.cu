// lvt_b200 :: VQ-VAE edge kernels (the 3-channel ends of ResEncoder / ResDecoder, the losses and
// small element-wise pieces).  The 128/256-channel convolutions run as implicit GEMMs on the tensor
// cores (gemm.cu conv modes); these kernels are bandwidth-bound.
// Reference: vidgen/modeling/encoder/resencoder.py:46-52, generator/resdecoder.py:48-57,66-69,
// meta_arch/ae.py:34-36,151-168, meta_arch/vqvae.py:66-91, loss/loss.py:5-20.
//
// "phase-major" layout of a 32x32 feature map: [hp][wp][n][16][16][C], pixel (2*h2+hp, 2*w2+wp).
#include "../../include/lvt_b200.h"
#include "common.cuh"

extern void lvt_count_launch(int n);

namespace {

LVT_DEVICE_INLINE float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Conv2d(3 -> nf/2, k4, s2, p1) input side: normalise (x-mean)/std and write the im2col matrix
// A1 [4*n*256, 64] bf16 (48 real columns k = (kh*4+kw)*3 + c, 16 zero columns), rows in phase-major
// order of the 32x32 output grid.  One thread per (row, tap).
// SPLIT: every value v is written as the bf16 pair hi = bf16(v), lo = bf16(v - hi) in three 64-column segments
// [hi | lo | hi] (row stride 192): the A operand of the 3-term split product of the high-precision encoder.
template <bool SPLIT>
__global__ void __launch_bounds__(256)
in_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ A, int n, float mean, float inv_std) {
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long rows = (long long)4 * n * 256;
  if (gid >= rows * 16) return;
  const int tap = (int)(gid & 15);
  const long long row = gid >> 4;
  const int pos = (int)(row & 255);
  const long long r2 = row >> 8;
  const int img = (int)(r2 % n);
  const int phase = (int)(r2 / n);
  const int oh = 2 * (pos >> 4) + (phase >> 1), ow = 2 * (pos & 15) + (phase & 1);
  const int kh = tap >> 2, kw = tap & 3;
  const int ih = 2 * oh - 1 + kh, iw = 2 * ow - 1 + kw;
  float v[3] = {0.f, 0.f, 0.f};
  if (ih >= 0 && ih < 64 && iw >= 0 && iw < 64) {
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (x[(((long long)img * 3 + c) * 64 + ih) * 64 + iw] - mean) * inv_std;
  }
  if constexpr (SPLIT) {
    __nv_bfloat16* a = A + row * 192 + tap * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const __nv_bfloat16 hi = __float2bfloat16(v[c]);
      a[c] = hi;
      a[64 + c] = __float2bfloat16(v[c] - __bfloat162float(hi));
      a[128 + c] = hi;
    }
    if (tap == 15) {
#pragma unroll
      for (int seg = 0; seg < 3; ++seg)
#pragma unroll
        for (int c = 48; c < 64; ++c) A[row * 192 + seg * 64 + c] = __float2bfloat16(0.f);
    }
  } else {
    __nv_bfloat16* a = A + row * 64 + tap * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) a[c] = __float2bfloat16(v[c]);
    if (tap == 15) {
#pragma unroll
      for (int c = 48; c < 64; ++c) A[row * 64 + c] = __float2bfloat16(0.f);
    }
  }
}

// The plain (training) form of the same im2col with one 16-byte store per thread: thread = (row, 8 consecutive
// elements of the 64-wide row); element e < 48 is (tap e / 3, channel e % 3), the rest is zero padding.  (One thread per
// (row, tap) wrote three 2-byte values each: 82 us for 67 MB at 512 frames.)
__global__ void __launch_bounds__(256)
in_im2col_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ A, int n, float mean, float inv_std) {
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long rows = (long long)4 * n * 256;
  if (gid >= rows * 8) return;
  const int chunk = (int)(gid & 7);
  const long long row = gid >> 3;
  const int pos = (int)(row & 255);
  const long long r2 = row >> 8;
  const int img = (int)(r2 % n);
  const int phase = (int)(r2 / n);
  const int oh = 2 * (pos >> 4) + (phase >> 1), ow = 2 * (pos & 15) + (phase & 1);
  float v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int e = chunk * 8 + u;
    v[u] = 0.f;
    if (e < 48) {
      const int tap = e / 3, c = e - tap * 3;
      const int ih = 2 * oh - 1 + (tap >> 2), iw = 2 * ow - 1 + (tap & 3);
      if (ih >= 0 && ih < 64 && iw >= 0 && iw < 64)
        v[u] = (__ldg(x + (((long long)img * 3 + c) * 64 + ih) * 64 + iw) - mean) * inv_std;
    }
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]);
  o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]);
  o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(A + row * 64 + chunk * 8) = o;
}

// v = [relu](in [+ add]) -> out_f32 (optional) and the bf16 split of v in three C-wide segments of a 3C-wide row:
// activations [hi | lo | hi], weights [hi | hi | lo], so that one GEMM over K = 3C contracts
// a_hi w_hi + a_lo w_hi + a_hi w_lo = a w up to the dropped a_lo w_lo term (~2^-17 relative instead of bf16's 2^-9).
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ in, const float* __restrict__ add, __nv_bfloat16* __restrict__ out,
              float* __restrict__ out_f32, long long rows, int C, int relu, int weight_pattern) {
  const int c4 = C >> 2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < rows * c4; i += (long long)gridDim.x * 256) {
    const long long r = i / c4;
    const int c = (int)(i - r * c4) * 4;
    float4 v = *reinterpret_cast<const float4*>(in + r * C + c);
    if (add) {
      const float4 a = *reinterpret_cast<const float4*>(add + r * C + c);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + r * C + c) = v;
    const float e[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hi[j] = __float2bfloat16(e[j]);
      lo[j] = __float2bfloat16(e[j] - __bfloat162float(hi[j]));
    }
    __nv_bfloat16* o = out + r * 3 * C + c;
    const uint2 uh = make_uint2(((uint32_t)__bfloat16_as_ushort(hi[1]) << 16) | __bfloat16_as_ushort(hi[0]),
                                ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16) | __bfloat16_as_ushort(hi[2]));
    const uint2 ul = make_uint2(((uint32_t)__bfloat16_as_ushort(lo[1]) << 16) | __bfloat16_as_ushort(lo[0]),
                                ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16) | __bfloat16_as_ushort(lo[2]));
    *reinterpret_cast<uint2*>(o) = uh;
    *reinterpret_cast<uint2*>(o + C) = weight_pattern ? uh : ul;
    *reinterpret_cast<uint2*>(o + 2 * C) = weight_pattern ? ul : uh;
  }
}

// ConvTranspose2d(C -> 3, k4, s2, p1) + tanh, second half.  The contraction over the C input channels is a
// tensor-core GEMM Y[row, (kh*4+kw)*3 + co] = sum_c act[row, c] * w[c][co][kh][kw] over the phase-major 32x32
// input pixels (rows); this kernel gathers, per OUTPUT pixel, the 2 x 2 (input pixel, tap) pairs that reach it:
// output row oy = 2Q+py takes kh = (py+1)%2 + 2a from input row iy = Q + (py+1-kh)/2.
// Y fp32 [4*n*256, 64]; out fp32 NCHW [n, 3, 64, 64].
__global__ void __launch_bounds__(256)
out_col2im_tanh_kernel(const float* __restrict__ Y, const float* __restrict__ bias, float* __restrict__ out, int n) {
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= (long long)n * 4096) return;
  const int img = (int)(pix >> 12);
  const int oy = (int)((pix >> 6) & 63), ox = (int)(pix & 63);
  const int py = oy & 1, px = ox & 1, Q = oy >> 1, R = ox >> 1;
  float acc[3] = {bias[0], bias[1], bias[2]};
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int kh = (py + 1) % 2 + 2 * a;
    const int iy = Q + (py + 1 - kh) / 2;
    if (iy < 0 || iy >= 32) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int kw = (px + 1) % 2 + 2 * b;
      const int ix = R + (px + 1 - kw) / 2;
      if (ix < 0 || ix >= 32) continue;
      const int phase = (iy & 1) * 2 + (ix & 1);
      const long long row = (((long long)phase * n + img) * 16 + (iy >> 1)) * 16 + (ix >> 1);
      const float* yp = Y + row * 64 + (kh * 4 + kw) * 3;
      acc[0] += yp[0];
      acc[1] += yp[1];
      acc[2] += yp[2];
    }
  }
#pragma unroll
  for (int co = 0; co < 3; ++co) out[(((long long)img * 3 + co) * 64 + oy) * 64 + ox] = tanhf(acc[co]);
}

// dpre[n,3,64,64] = dL/d(pre-tanh) for the reconstruction MSE (loss.py:20, vqvae.py:79):
//   L = lambda * mean((xt - xn)^2), xn = (x - mean)/std  =>  dpre = 2*lambda/numel * (xt - xn) * (1 - xt^2)
// also accumulates the loss value and the bias gradient.
__global__ void __launch_bounds__(256)
recon_loss_kernel(const float* __restrict__ xt, const float* __restrict__ x, float* __restrict__ dpre,
                  float* __restrict__ loss, float* __restrict__ dbias, long long numel, float mean,
                  float inv_std, float lambda) {
  __shared__ float s_l[8], s_b[8][3];
  float l = 0.f, b[3] = {0.f, 0.f, 0.f};
  const float scale = 2.f * lambda / (float)numel;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < numel; i += (long long)gridDim.x * 256) {
    const float t = xt[i];
    const float d = t - (x[i] - mean) * inv_std;
    l += d * d;
    const float g = scale * d * (1.f - t * t);
    if (dpre) dpre[i] = g;
    const int c = (int)((i >> 12) % 3);
    b[0] += c == 0 ? g : 0.f;
    b[1] += c == 1 ? g : 0.f;
    b[2] += c == 2 ? g : 0.f;
  }
  l = warp_sum(l);
#pragma unroll
  for (int c = 0; c < 3; ++c) b[c] = warp_sum(b[c]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_l[warp] = l;
    s_b[warp][0] = b[0]; s_b[warp][1] = b[1]; s_b[warp][2] = b[2];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f, tb[3] = {0.f, 0.f, 0.f};
    for (int w = 0; w < 8; ++w) {
      t += s_l[w];
      tb[0] += s_b[w][0]; tb[1] += s_b[w][1]; tb[2] += s_b[w][2];
    }
    atomicAdd(loss, t * lambda / (float)numel);
    if (dbias) {
      atomicAdd(dbias, tb[0]); atomicAdd(dbias + 1, tb[1]); atomicAdd(dbias + 2, tb[2]);
    }
  }
}

// Backward of the output ConvTranspose2d, first half: G[row, co*16 + kh*4 + kw] = dpre[co, oy, ox] at the
// output pixel (2*iy - 1 + kh, 2*ix - 1 + kw) that tap (kh, kw) of input pixel `row` reaches (0 outside the
// image; bf16, 48 columns + 16 zeros; rows in the phase-major order of the 32x32 input).  Both gradients are
// then tensor-core GEMMs:  dW[c][k] = sum_rows act[row, c] * G[row, k]  and
// dact[row, c] = relu'(act[row, c]) * sum_k G[row, k] * w[c][k].
__global__ void __launch_bounds__(256)
out_convt_g_kernel(const float* __restrict__ dpre, __nv_bfloat16* __restrict__ G, int n) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;  // one thread = one row and two columns
  const long long row = t >> 5;
  if (row >= (long long)n * 1024) return;
  const int k0 = (int)(t & 31) * 2;
  const int pos = (int)(row & 255);
  const long long r2 = row >> 8;
  const int img = (int)(r2 % n), phase = (int)(r2 / n);
  const int iy = 2 * (pos >> 4) + (phase >> 1), ix = 2 * (pos & 15) + (phase & 1);
  float v[2] = {0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = k0 + e;
    if (k < 48) {
      const int co = k >> 4, kh = (k >> 2) & 3, kw = k & 3;
      const int oy = 2 * iy - 1 + kh, ox = 2 * ix - 1 + kw;
      if (oy >= 0 && oy < 64 && ox >= 0 && ox < 64) v[e] = __ldg(dpre + (((long long)img * 3 + co) * 64 + oy) * 64 + ox);
    }
  }
  reinterpret_cast<__nv_bfloat162*>(G)[t] = __floats2bfloat162_rn(v[0], v[1]);
}

// Commitment loss + gradient wrt z_e, merged with the straight-through gradient (vqvae.py:75,86;
// vq_utils.py:50-53):  L = beta * mean((z_e - zq_bar)^2);  dz = dz_st + 2*beta/numel * (z_e - zq_bar)
__global__ void __launch_bounds__(256)
commit_loss_kernel(const float* __restrict__ z_e, const float* __restrict__ zq_bar, const float* __restrict__ dz_st,
                   __nv_bfloat16* __restrict__ dz, float* __restrict__ loss, long long numel, float beta) {
  __shared__ float s_l[8];
  float l = 0.f;
  const float scale = 2.f * beta / (float)numel;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < numel; i += (long long)gridDim.x * 256) {
    const float d = z_e[i] - zq_bar[i];
    l += d * d;
    if (dz) dz[i] = __float2bfloat16((dz_st ? dz_st[i] : 0.f) + scale * d);
  }
  l = warp_sum(l);
  if ((threadIdx.x & 31) == 0) s_l[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_l[w];
    atomicAdd(loss, t * beta / (float)numel);
  }
}

// out = (a [+ b]) * (mask_src > 0), bf16 (ReLU backward joined with a skip-connection gradient)
__global__ void __launch_bounds__(256)
relu_bwd_add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                    const __nv_bfloat16* __restrict__ mask_src, __nv_bfloat16* __restrict__ out, long long n2) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n2; i += (long long)gridDim.x * 256) {
    float2 v = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(a)[i]);
    if (b) {
      const float2 w = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(b)[i]);
      v.x += w.x;
      v.y += w.y;
    }
    const float2 m = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(mask_src)[i]);
    reinterpret_cast<__nv_bfloat162*>(out)[i] = __floats2bfloat162_rn(m.x > 0.f ? v.x : 0.f, m.y > 0.f ? v.y : 0.f);
  }
}

// x (fp32) += a (bf16), four elements per thread: the skip connection of the last encoder ResBlock, whose fp32 output
// (z_e) cannot take a bf16 addend inside the TMA-store GEMM epilogue (resencoder.py:19-21)
__global__ void __launch_bounds__(256)
add_bf16_to_f32_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ a, long long n4) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    const uint2 u = reinterpret_cast<const uint2*>(a)[i];
    const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    v.x += lo.x; v.y += lo.y; v.z += hi.x; v.w += hi.y;
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

// fp32 -> bf16 with optional ReLU (z_q / activations entering a conv)
__global__ void __launch_bounds__(256)
cast_relu_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n, int relu) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float v = in[i];
    out[i] = __float2bfloat16(relu ? fmaxf(v, 0.f) : v);
  }
}

// reconstruction post-processing of the inference path (ae.py:130-139): y*std + mean, clamp to [lo, hi]
__global__ void __launch_bounds__(256)
denorm_clamp_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, float mean, float std,
                    float lo, float hi) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    out[i] = fminf(fmaxf(in[i] * std + mean, lo), hi);
}

int grid_for(long long n) { return (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8); }

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int lvt_vqvae_in_im2col(const float* x, void* a_bf16, int n, float mean, float std, void* stream) {
  LVT_CHECK_ARG(x && a_bf16 && n > 0 && std != 0.f, "lvt_vqvae_in_im2col: bad argument");
  const long long threads = (long long)4 * n * 256 * 8;
  in_im2col_rows_kernel<<<lvt_ceil_div(threads, 256), 256, 0, STREAM(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(a_bf16), n, mean, 1.f / std);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vqvae_in_im2col_split(const float* x, void* a_bf16, int n, float mean, float std, void* stream) {
  LVT_CHECK_ARG(x && a_bf16 && n > 0 && std != 0.f, "lvt_vqvae_in_im2col_split: bad argument");
  const long long threads = (long long)4 * n * 256 * 16;
  in_im2col_kernel<true><<<lvt_ceil_div(threads, 256), 256, 0, STREAM(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(a_bf16), n, mean, 1.f / std);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_split3_bf16(const float* in, const float* add, void* out_bf16, float* out_f32, long long rows, int C,
                               int relu, int weight_pattern, void* stream) {
  LVT_CHECK_ARG(in && out_bf16 && rows > 0 && C > 0 && C % 4 == 0, "lvt_split3_bf16: bad argument");
  split3_kernel<<<grid_for(rows * (C / 4)), 256, 0, STREAM(stream)>>>(in, add, reinterpret_cast<__nv_bfloat16*>(out_bf16),
                                                                      out_f32, rows, C, relu, weight_pattern);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vqvae_out_col2im_tanh(const float* y, const float* bias, float* out, int n, void* stream) {
  LVT_CHECK_ARG(y && bias && out && n > 0, "lvt_vqvae_out_col2im_tanh: bad argument");
  out_col2im_tanh_kernel<<<lvt_ceil_div((long long)n * 4096, 256), 256, 0, STREAM(stream)>>>(y, bias, out, n);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vqvae_recon_loss(const float* x_tilde, const float* x, float* dpre, float* loss, float* dbias,
                                    int n, float mean, float std, float lambda, void* stream) {
  LVT_CHECK_ARG(x_tilde && x && loss && n > 0 && std != 0.f, "lvt_vqvae_recon_loss: bad argument");
  const long long numel = (long long)n * 3 * 4096;
  recon_loss_kernel<<<grid_for(numel), 256, 0, STREAM(stream)>>>(x_tilde, x, dpre, loss, dbias, numel, mean,
                                                                 1.f / std, lambda);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vqvae_out_convt_g(const float* dpre, void* g_bf16, int n, void* stream) {
  LVT_CHECK_ARG(dpre && g_bf16 && n > 0, "lvt_vqvae_out_convt_g: bad argument");
  out_convt_g_kernel<<<lvt_ceil_div((long long)n * 1024 * 32, 256), 256, 0, STREAM(stream)>>>(
      dpre, reinterpret_cast<__nv_bfloat16*>(g_bf16), n);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vqvae_commit_loss(const float* z_e, const float* zq_bar, const float* dz_st, void* dz_bf16,
                                     float* loss, long long numel, float beta, void* stream) {
  LVT_CHECK_ARG(z_e && zq_bar && loss && numel > 0, "lvt_vqvae_commit_loss: bad argument");
  commit_loss_kernel<<<grid_for(numel), 256, 0, STREAM(stream)>>>(z_e, zq_bar, dz_st,
                                                                  reinterpret_cast<__nv_bfloat16*>(dz_bf16), loss, numel,
                                                                  beta);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_relu_bwd_add(const void* a_bf16, const void* b_bf16, const void* mask_src_bf16, void* out_bf16,
                                long long n, void* stream) {
  LVT_CHECK_ARG(a_bf16 && mask_src_bf16 && out_bf16 && n > 0 && n % 2 == 0, "lvt_relu_bwd_add: bad argument");
  relu_bwd_add_kernel<<<grid_for(n / 2), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(a_bf16), reinterpret_cast<const __nv_bfloat16*>(b_bf16),
      reinterpret_cast<const __nv_bfloat16*>(mask_src_bf16), reinterpret_cast<__nv_bfloat16*>(out_bf16), n / 2);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_add_bf16_to_f32(float* x, const void* a_bf16, long long n, void* stream) {
  LVT_CHECK_ARG(x && a_bf16 && n > 0 && n % 4 == 0, "lvt_add_bf16_to_f32: bad argument");
  add_bf16_to_f32_kernel<<<grid_for(n / 4), 256, 0, STREAM(stream)>>>(x, reinterpret_cast<const __nv_bfloat16*>(a_bf16), n / 4);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_cast_relu_bf16(const float* in, void* out_bf16, long long n, int relu, void* stream) {
  LVT_CHECK_ARG(in && out_bf16 && n > 0, "lvt_cast_relu_bf16: bad argument");
  cast_relu_kernel<<<grid_for(n), 256, 0, STREAM(stream)>>>(in, reinterpret_cast<__nv_bfloat16*>(out_bf16), n, relu);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_denorm_clamp(const float* in, float* out, long long n, float mean, float std, float lo, float hi,
                                void* stream) {
  LVT_CHECK_ARG(in && out && n > 0, "lvt_denorm_clamp: bad argument");
  denorm_clamp_kernel<<<grid_for(n), 256, 0, STREAM(stream)>>>(in, out, n, mean, std, lo, hi);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}
