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
  __nv_bfloat16* a = A + row * 64 + tap * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) a[c] = __float2bfloat16(v[c]);
  if (tap == 15) {
#pragma unroll
    for (int c = 48; c < 64; ++c) A[row * 64 + c] = __float2bfloat16(0.f);
  }
}

// ConvTranspose2d(C -> 3, k4, s2, p1) + tanh on a phase-major 32x32 input (already ReLU'd).
// One warp per output pixel; lanes split the C channels; w: [C][3][4][4] fp32.
// out: fp32 NCHW [n, 3, 64, 64].  Output row oy = 2Q+py takes kh = py+1 (mod 2): iy = Q + (py+1-kh)/2.
template <int C>
__global__ void __launch_bounds__(256)
out_convt_fwd_kernel(const __nv_bfloat16* __restrict__ act, const float* __restrict__ w,
                     const float* __restrict__ bias, float* __restrict__ out, int n) {
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)n * 4096) return;
  const int img = (int)(wid >> 12);
  const int oy = (int)((wid >> 6) & 63), ox = (int)(wid & 63);
  const int py = oy & 1, px = ox & 1, Q = oy >> 1, R = ox >> 1;
  float acc[3] = {0.f, 0.f, 0.f};
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
      const __nv_bfloat16* ap = act + ((((long long)phase * n + img) * 16 + (iy >> 1)) * 16 + (ix >> 1)) * C;
      for (int c = lane; c < C; c += 32) {
        const float v = __bfloat162float(ap[c]);
        const float* wp = w + (long long)c * 48 + kh * 4 + kw;
        acc[0] += v * wp[0];
        acc[1] += v * wp[16];
        acc[2] += v * wp[32];
      }
    }
  }
#pragma unroll
  for (int co = 0; co < 3; ++co) acc[co] = warp_sum(acc[co]);
  if (lane < 3) {
    const float v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : acc[2]);
    out[(((long long)img * 3 + lane) * 64 + oy) * 64 + ox] = tanhf(v + bias[lane]);
  }
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

// Backward of the output ConvTranspose2d wrt its (phase-major, ReLU'd) input.
// One warp per INPUT pixel of the 32x32 grid; lanes split channels.  For input (iy, ix) and tap
// (kh, kw) the output pixel is (2*iy - 1 + kh, 2*ix - 1 + kw).
//   dact[c]  = relu'(act[c]) * sum_{co,kh,kw} dpre[co, oy, ox] * w[c][co][kh][kw]     (bf16, phase-major)
//   G[row, co*16 + kh*4 + kw] = dpre[co, oy, ox] (bf16, 48 columns + 16 zeros): the weight gradient is
//   then the tensor-core GEMM dw[c][k] = sum_rows act[row, c] * G[row, k] (no atomics).
template <int C>
__global__ void __launch_bounds__(256)
out_convt_bwd_kernel(const __nv_bfloat16* __restrict__ act, const float* __restrict__ w,
                     const float* __restrict__ dpre, __nv_bfloat16* __restrict__ dact,
                     __nv_bfloat16* __restrict__ G, int n) {
  __shared__ float s_w[C * 48];
  for (int i = threadIdx.x; i < C * 48; i += 256) s_w[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long total = (long long)n * 1024;
  for (long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); wid < total; wid += (long long)gridDim.x * 8) {
    // phase-major enumeration so that reads / writes of act / dact are contiguous per warp
    const int pos = (int)(wid & 255);
    const long long r2 = wid >> 8;
    const int img = (int)(r2 % n), phase = (int)(r2 / n);
    const int iy = 2 * (pos >> 4) + (phase >> 1), ix = 2 * (pos & 15) + (phase & 1);
    float g[48];  // dpre of the 16 taps x 3 channels this input pixel feeds (0 outside the image)
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int oy = 2 * iy - 1 + kh;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int ox = 2 * ix - 1 + kw;
        const bool ok = oy >= 0 && oy < 64 && ox >= 0 && ox < 64;
#pragma unroll
        for (int co = 0; co < 3; ++co)
          g[co * 16 + kh * 4 + kw] = ok ? __ldg(dpre + (((long long)img * 3 + co) * 64 + oy) * 64 + ox) : 0.f;
      }
    }
    if (G) {
      // lanes 0..23 write two of the 48 values each, lanes 24..31 the zero padding
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int k = 0; k < 48; k += 2) {
        if (lane == k / 2) { v0 = g[k]; v1 = g[k + 1]; }
      }
      reinterpret_cast<__nv_bfloat162*>(G + wid * 64)[lane] = __floats2bfloat162_rn(v0, v1);
    }
    const long long base = wid * C;
    for (int c = lane; c < C; c += 32) {
      const float a = __bfloat162float(act[base + c]);
      const float* wp = s_w + c * 48;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 48; ++k) s += g[k] * wp[k];
      dact[base + c] = __float2bfloat16(a > 0.f ? s : 0.f);
    }
  }
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
  const long long threads = (long long)4 * n * 256 * 16;
  in_im2col_kernel<<<lvt_ceil_div(threads, 256), 256, 0, STREAM(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(a_bf16), n, mean, 1.f / std);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vqvae_out_convt_fwd(const void* act_bf16, const float* w, const float* bias, float* out, int n,
                                       int C, void* stream) {
  LVT_CHECK_ARG(act_bf16 && w && bias && out && n > 0, "lvt_vqvae_out_convt_fwd: bad argument");
  LVT_CHECK_ARG(C == 128, "lvt_vqvae_out_convt_fwd: C must be 128 (NF/2 of the shipped configs)");
  out_convt_fwd_kernel<128><<<lvt_ceil_div((long long)n * 4096, 8), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(act_bf16), w, bias, out, n);
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

extern "C" int lvt_vqvae_out_convt_bwd(const void* act_bf16, const float* w, const float* dpre, void* dact_bf16,
                                       void* g_bf16, int n, int C, void* stream) {
  LVT_CHECK_ARG(act_bf16 && w && dpre && dact_bf16 && n > 0, "lvt_vqvae_out_convt_bwd: bad argument");
  LVT_CHECK_ARG(C == 128, "lvt_vqvae_out_convt_bwd: C must be 128");
  const long long warps = (long long)n * 1024;
  const int blocks = (int)(warps / 8 < 148 * 8 ? (warps + 7) / 8 : 148 * 8);
  out_convt_bwd_kernel<128><<<blocks > 0 ? blocks : 1, 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(act_bf16), w, dpre, reinterpret_cast<__nv_bfloat16*>(dact_bf16),
      reinterpret_cast<__nv_bfloat16*>(g_bf16), n);
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
