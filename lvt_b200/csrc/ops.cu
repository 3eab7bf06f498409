// lvt_b200 :: bandwidth-bound kernels of the DSFVT path: LayerNorm, sparse (one-hot) front
// ends, attention helpers, cross-entropy, optimizers, layout packing.  Each replaces the ATen
// calls cited at its entry point (reference paths relative to the reference repo root).
#include "../../include/lvt_b200.h"
#include "common.cuh"

extern void lvt_count_launch(int n);

namespace {

constexpr int kSMs = 148;

LVT_DEVICE_INLINE float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
LVT_DEVICE_INLINE void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
LVT_DEVICE_INLINE void st_bf16x4(__nv_bfloat16* p, float a, float b, float c, float d) {
  uint2 u;
  u.x = pack_bf16x2(a, b);
  u.y = pack_bf16x2(c, d);
  *reinterpret_cast<uint2*>(p) = u;
}
LVT_DEVICE_INLINE float4 ld_bf16x4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, row kept in registers (d = 128*V4, V4 <= 8).
// ------------------------------------------------------------------------------------------
template <int V4>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
              float* __restrict__ mean, float* __restrict__ rstd, int M, float eps) {
  pdl_prologue();
  constexpr int d = V4 * 128;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (size_t)row * d;
  float4 v[V4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    v[i] = ld4(xr + (i * 32 + lane) * 4);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mu = warp_sum(s) * (1.f / d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, e = v[i].w - mu;
    q += (a * a + b * b) + (c * c + e * e);
  }
  const float rs = rsqrtf(warp_sum(q) * (1.f / d) + eps);
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
  __nv_bfloat16* yr = y + (size_t)row * d;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 g = ld4(gamma + c), b = ld4(beta + c);
    st_bf16x4(yr + c, (v[i].x - mu) * rs * g.x + b.x, (v[i].y - mu) * rs * g.y + b.y,
              (v[i].z - mu) * rs * g.z + b.z, (v[i].w - mu) * rs * g.w + b.w);
  }
}

// LayerNorm backward. dx = dres + rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma.
// dgamma/dbeta: per-warp register partials over a strided set of rows, block reduce, atomics.
// DYB: dy arrives as bf16 (written by the GEMM that produced it) instead of fp32.
template <int V4, bool DYB>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x,
              const float* __restrict__ mean, const float* __restrict__ rstd,
              const float* __restrict__ gamma, const float* __restrict__ dres,
              float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum, int M) {
  pdl_prologue();
  constexpr int d = V4 * 128;
  __shared__ float s_red[8][d + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 gm[V4], ag[V4], ab[V4], ac[V4];  // ac: column sums of dx (the bias gradient of the Linear that produced x)
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    gm[i] = ld4(gamma + (i * 32 + lane) * 4);
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ac[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    const float mu = mean[row], rs = rstd[row];
    const float* xr = x + (size_t)row * d;
    float4 xh[V4], g[V4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 xv = ld4(xr + c);
      float4 dv;
      if constexpr (DYB) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_) + (size_t)row * d + c);
        const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        dv = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {
        dv = ld4(reinterpret_cast<const float*>(dy_) + (size_t)row * d + c);
      }
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[i] = make_float4(dv.x * gm[i].x, dv.y * gm[i].y, dv.z * gm[i].z, dv.w * gm[i].w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      ag[i].x += dv.x * xh[i].x; ag[i].y += dv.y * xh[i].y; ag[i].z += dv.z * xh[i].z; ag[i].w += dv.w * xh[i].w;
      ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
    }
    const float m1 = warp_sum(s1) * (1.f / d), m2 = warp_sum(s2) * (1.f / d);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 o = make_float4(rs * (g[i].x - m1 - xh[i].x * m2), rs * (g[i].y - m1 - xh[i].y * m2),
                             rs * (g[i].z - m1 - xh[i].z * m2), rs * (g[i].w - m1 - xh[i].w * m2));
      if (dres) {
        const float4 r = ld4(dres + (size_t)row * d + c);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      if (dx) st4(dx + (size_t)row * d + c, o);
      if (dx_bf16) st_bf16x4(dx_bf16 + (size_t)row * d + c, o.x, o.y, o.z, o.w);
      ac[i].x += o.x; ac[i].y += o.y; ac[i].z += o.z; ac[i].w += o.w;
    }
  }
  // block reduction of the parameter gradients
  for (int pass = 0; pass < (dxsum ? 3 : 2); ++pass) {
#pragma unroll
    for (int i = 0; i < V4; ++i) st4(&s_red[warp][(i * 32 + lane) * 4], pass == 0 ? ag[i] : (pass == 1 ? ab[i] : ac[i]));
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += 256) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += s_red[w][c];
      atomicAdd((pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum)) + c, s);
    }
    __syncthreads();
  }
}

// out[n] += sum_m x[m, n]  (bias gradients).  Block = 32x8 threads over a 64-column strip.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int M, int N,
                   long long ld, int rows_per_block) {
  pdl_prologue();
  __shared__ float s[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + tx * 2;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float a = 0.f, b = 0.f;
  if (c < N) {
    for (int r = r0 + ty; r < r1; r += 8) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + (size_t)r * ld + c));
      a += v.x;
      b += v.y;
    }
  }
  s[ty][tx * 2] = a;
  s[ty][tx * 2 + 1] = b;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s[w][threadIdx.x];
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < N) atomicAdd(out + cc, t);
  }
}

// delta[b, h, i] = sum_d dO[b*L+i, h*da+d] * O[b*L+i, h*da+d]   (softmax backward row term)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O,
                  float* __restrict__ delta, int nb, int H, int L, int da) {
  pdl_prologue();
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long total = (long long)nb * L * H;
  if (wid >= total) return;
  const int h = (int)(wid % H);
  const long long row = wid / H;  // b*L + i
  const __nv_bfloat16* a = dO + (size_t)row * H * da + (size_t)h * da;
  const __nv_bfloat16* b = O + (size_t)row * H * da + (size_t)h * da;
  float s = 0.f;
  for (int c = lane * 4; c < da; c += 128) {
    const float4 u = ld_bf16x4(a + c), v = ld_bf16x4(b + c);
    s += (u.x * v.x + u.y * v.y) + (u.z * v.z + u.w * v.w);
  }
  s = warp_sum(s);
  if (lane == 0) {
    const long long bb = row / L;
    const int i = (int)(row - bb * L);
    delta[(bb * H + h) * L + i] = s;
  }
}

// Gradient of the relative-position banks (get_B, vt_attention.py:169-174):
// dbank_x[h, off] += sum_{b} sum_{i,j : off_x(i,j) == off} dS[b, h, i, j].
// grid (L*L/2048, H, zsplit), 256 threads; a thread owns 8 consecutive keys j of one query i,
// streams them over its batch range with 16 B loads, then bins through shared memory.
__global__ void __launch_bounds__(256)
relpos_bank_grad_kernel(const __nv_bfloat16* __restrict__ dS, float* __restrict__ dbt,
                        float* __restrict__ dbh, float* __restrict__ dbw, int nb, int H, int bt,
                        int bh, int bw) {
  pdl_prologue();
  const int L = bt * bh * bw;  // == 256
  const int h = blockIdx.y;
  const int e0 = blockIdx.x * 2048 + threadIdx.x * 8;  // element of the L x L plane
  const int i = e0 / L, j0 = e0 % L;
  const int nt = 2 * bt - 1, nh = 2 * bh - 1, nw = 2 * bw - 1;
  const int per = (nb + gridDim.z - 1) / gridDim.z;
  const int b0 = blockIdx.z * per, b1 = min(nb, b0 + per);
  const size_t plane = (size_t)L * L;
  const __nv_bfloat16* p = dS + (size_t)h * plane + e0;
  const size_t bstride = (size_t)H * plane;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int b = b0;
  for (; b + 4 <= b1; b += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(p + (size_t)(b + u) * bstride);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const __nv_bfloat162* q = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(q[k]);
        acc[2 * k] += f.x;
        acc[2 * k + 1] += f.y;
      }
    }
  }
  for (; b < b1; ++b) {
    const uint4 v = *reinterpret_cast<const uint4*>(p + (size_t)b * bstride);
    const __nv_bfloat162* q = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(q[k]);
      acc[2 * k] += f.x;
      acc[2 * k + 1] += f.y;
    }
  }
  // bin through per-warp private shared-memory bins; equal consecutive bins of a thread are merged
  // first (8 consecutive keys share tj and, for bw >= 8, hj), and a single-entry t-bank (bt == 1)
  // is reduced with shuffles, so same-address atomic serialisation stays small.
  __shared__ float wbins[8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = lane; k < 64; k += 32) wbins[warp][k] = 0.f;
  __syncwarp();
  const int hw = bh * bw;
  const int ti = i / hw, hi = (i / bw) % bh, wi = i % bw;
  float* wb = wbins[warp];
  int prev_t = -1, prev_h = -1;
  float run_t = 0.f, run_h = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int j = j0 + u;
    const int tj = j / hw, hj = (j / bw) % bh, wj = j % bw;
    const int bin_t = ti - tj + bt - 1, bin_h = nt + hi - hj + bh - 1;
    if (bin_t != prev_t) {
      if (prev_t >= 0 && bt > 1) atomicAdd(&wb[prev_t], run_t);
      prev_t = bin_t;
      run_t = (bt > 1) ? 0.f : run_t;
    }
    run_t += acc[u];
    if (bin_h != prev_h) {
      if (prev_h >= 0) atomicAdd(&wb[prev_h], run_h);
      prev_h = bin_h;
      run_h = 0.f;
    }
    run_h += acc[u];
    atomicAdd(&wb[nt + nh + wi - wj + bw - 1], acc[u]);
  }
  atomicAdd(&wb[prev_h], run_h);
  if (bt > 1) {
    atomicAdd(&wb[prev_t], run_t);
  } else {
    run_t = warp_sum(run_t);
    if (lane == 0) wb[0] += run_t;
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < nt + nh + nw) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += wbins[w][t];
    if (t < nt) atomicAdd(dbt + h * nt + t, v);
    else if (t < nt + nh) atomicAdd(dbh + h * nh + (t - nt), v);
    else atomicAdd(dbw + h * nw + (t - nt - nh), v);
  }
}

// ------------------------------------------------------------------------------------------
// VTEncoder front end (videotransformer.py:41-53): Conv3d over a one-hot input == sum of the
// selected weight rows; padded entries contribute nothing.  wt is the conv weight re-laid out as
// [nc][kt][kh][kw][nv][de] so that a selected row is contiguous.  One warp per output position.
// ------------------------------------------------------------------------------------------
struct EncFrontDims {
  int B, nc, nv, de, Tc, Hc, Wc, kt, kh, kw, st, sh, sw, to, ho, wo, pad_value;
};

// code of tap `tap` (= ((c*kt + i)*kh + j)*kw + l, the row-block order of the re-laid-out weight) at output (b, to, ho, wo)
LVT_DEVICE_INLINE long long enc_tap_code(const int64_t* __restrict__ ctx, const EncFrontDims& D, int b, int to, int ho,
                                         int wo, int tap) {
  const int l = tap % D.kw;
  int r = tap / D.kw;
  const int j = r % D.kh;
  r /= D.kh;
  const int i = r % D.kt, c = r / D.kt;
  const int tt = to * D.st + i, hh = ho * D.sh + j, ww = wo * D.sw + l;
  return ctx[((((size_t)b * D.nc + c) * D.Tc + tt) * D.Hc + hh) * D.Wc + ww];
}

// One warp per output position.  The taps' codes are read by the lanes (one tap each, 32 independent loads) and handed
// round by shuffle, and the weight rows are fetched four at a time: two dependent L2 round trips per 32 taps instead of
// two per tap.  The rows are still added in tap order, so the result is bit for bit what the sequential loop gave.
__global__ void __launch_bounds__(256)
enc_front_fwd_kernel(const int64_t* __restrict__ ctx, const int64_t* __restrict__ slice_idx,
                     const float* __restrict__ wt, const float* __restrict__ bias,
                     const float* __restrict__ slice_emb, __nv_bfloat16* __restrict__ out,
                     EncFrontDims D) {
  pdl_prologue();
  const long long pos = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long per = (long long)D.to * D.ho * D.wo;
  if (pos >= (long long)D.B * per) return;
  const int b = (int)(pos / per);
  int r = (int)(pos - b * per);
  const int to = r / (D.ho * D.wo);
  r -= to * D.ho * D.wo;
  const int ho = r / D.wo, wo = r - ho * D.wo;
  const float* se = slice_emb + (size_t)slice_idx[b] * D.de;
  const int ntap = D.nc * D.kt * D.kh * D.kw;
  for (int c0 = lane * 4; c0 - lane * 4 < D.de; c0 += 128) {  // every lane runs every pass (shuffles inside)
    const bool act = c0 < D.de;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) {
      acc = ld4(bias + c0);
      const float4 s4 = ld4(se + c0);
      acc.x += s4.x; acc.y += s4.y; acc.z += s4.z; acc.w += s4.w;
    }
    for (int t0 = 0; t0 < ntap; t0 += 32) {
      const long long mycode = (t0 + lane < ntap) ? enc_tap_code(ctx, D, b, to, ho, wo, t0 + lane) : (long long)D.pad_value;
      const int nn = min(32, ntap - t0);
      for (int u = 0; u < nn; u += 4) {
        float4 w4[4];
        bool ok[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const long long code = __shfl_sync(0xffffffffu, mycode, min(u + v, 31));
          ok[v] = act && u + v < nn && code != D.pad_value;
          if (ok[v]) w4[v] = ld4(wt + ((size_t)(t0 + u + v) * D.nv + (size_t)code) * D.de + c0);
        }
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (ok[v]) { acc.x += w4[v].x; acc.y += w4[v].y; acc.z += w4[v].z; acc.w += w4[v].w; }
      }
    }
    if (act) st_bf16x4(out + (size_t)pos * D.de + c0, acc.x, acc.y, acc.z, acc.w);
  }
}

// backward: scatter d_out rows into the re-laid-out weight gradient and the slice embedding (codes as in the forward).
__global__ void __launch_bounds__(256)
enc_front_bwd_kernel(const int64_t* __restrict__ ctx, const int64_t* __restrict__ slice_idx,
                     const float* __restrict__ dout, float* __restrict__ dwt,
                     float* __restrict__ dslice_emb, EncFrontDims D) {
  pdl_prologue();
  const long long pos = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long per = (long long)D.to * D.ho * D.wo;
  if (pos >= (long long)D.B * per) return;
  const int b = (int)(pos / per);
  int r = (int)(pos - b * per);
  const int to = r / (D.ho * D.wo);
  r -= to * D.ho * D.wo;
  const int ho = r / D.wo, wo = r - ho * D.wo;
  float* se = dslice_emb + (size_t)slice_idx[b] * D.de;
  const int ntap = D.nc * D.kt * D.kh * D.kw;
  for (int c0 = lane * 4; c0 - lane * 4 < D.de; c0 += 128) {
    const bool act = c0 < D.de;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) {
      g = ld4(dout + (size_t)pos * D.de + c0);
      red_add_f32x4(se + c0, g.x, g.y, g.z, g.w);
    }
    for (int t0 = 0; t0 < ntap; t0 += 32) {
      const long long mycode = (t0 + lane < ntap) ? enc_tap_code(ctx, D, b, to, ho, wo, t0 + lane) : (long long)D.pad_value;
      const int nn = min(32, ntap - t0);
      for (int u = 0; u < nn; ++u) {
        const long long code = __shfl_sync(0xffffffffu, mycode, u);
        if (act && code != D.pad_value)
          red_add_f32x4(dwt + ((size_t)(t0 + u) * D.nv + (size_t)code) * D.de + c0, g.x, g.y, g.z, g.w);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// VTDecoder front end (videotransformer.py:80-89,96-97 + vt_utils.py:183-200): summed channel
// embeddings, then the causal 3-D conv as an implicit-GEMM A operand: row m, tap q holds the
// embedding sum at position m + off_q (zero outside the slice).  taps = (dt,dh,dw) offsets.
// ------------------------------------------------------------------------------------------
struct DecFrontDims {
  int B, nc, nv, de, t, h, w, ntaps;
};

__global__ void __launch_bounds__(256)
dec_front_fwd_kernel(const int64_t* __restrict__ slc, const float* __restrict__ emb,
                     const int* __restrict__ taps, __nv_bfloat16* __restrict__ out, DecFrontDims D) {
  pdl_prologue();
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int thw = D.t * D.h * D.w;
  const long long total = (long long)D.B * thw * D.ntaps;
  if (wid >= total) return;
  const int q = (int)(wid % D.ntaps);
  const long long m = wid / D.ntaps;
  const int b = (int)(m / thw);
  int r = (int)(m - (long long)b * thw);
  const int tt = r / (D.h * D.w) + taps[3 * q];
  r %= D.h * D.w;
  const int hh = r / D.w + taps[3 * q + 1], ww = r % D.w + taps[3 * q + 2];
  const bool inside = tt >= 0 && tt < D.t && hh >= 0 && hh < D.h && ww >= 0 && ww < D.w;
  __nv_bfloat16* o = out + ((size_t)m * D.ntaps + q) * D.de;
  for (int c0 = lane * 4; c0 < D.de; c0 += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside) {
      const int p = (tt * D.h + hh) * D.w + ww;
      for (int k = 0; k < D.nc; ++k) {
        const long long code = slc[((size_t)b * D.nc + k) * thw + p];
        const float4 e = ld4(emb + ((size_t)k * D.nv + code) * D.de + c0);
        acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
      }
    }
    st_bf16x4(o + c0, acc.x, acc.y, acc.z, acc.w);
  }
}

// backward: demb[p] = sum_q dA[p - off_q, q] (gather form, no atomics), then one atomic scatter
// per channel table.
__global__ void __launch_bounds__(256)
dec_front_bwd_kernel(const int64_t* __restrict__ slc, const float* __restrict__ dA,
                     const int* __restrict__ taps, float* __restrict__ demb_tab, DecFrontDims D) {
  pdl_prologue();
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int thw = D.t * D.h * D.w;
  if (m >= (long long)D.B * thw) return;
  const int b = (int)(m / thw);
  const int p = (int)(m - (long long)b * thw);
  const int pt = p / (D.h * D.w), ph = (p / D.w) % D.h, pw = p % D.w;
  for (int c0 = lane * 4; c0 < D.de; c0 += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < D.ntaps; ++q) {
      const int tt = pt - taps[3 * q], hh = ph - taps[3 * q + 1], ww = pw - taps[3 * q + 2];
      if (tt < 0 || tt >= D.t || hh < 0 || hh >= D.h || ww < 0 || ww >= D.w) continue;
      const size_t src = (size_t)b * thw + (tt * D.h + hh) * D.w + ww;
      const float4 g = ld4(dA + (src * D.ntaps + q) * D.de + c0);
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
    for (int k = 0; k < D.nc; ++k) {
      const long long code = slc[((size_t)b * D.nc + k) * thw + p];
      red_add_f32x4(demb_tab + ((size_t)k * D.nv + code) * D.de + c0, acc.x, acc.y, acc.z, acc.w);
    }
  }
}

// ------------------------------------------------------------------------------------------
// ChannelPredictor (videotransformer.py:148-150): U[k]([y | onehot(codes of channels < k)]) =
// dense GEMM part (done by lvt_gemm_bf16, bias included) + gathered rows of the transposed
// one-hot block ut[(j*nv + code), d]; then ReLU.  One warp per row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
chpred_combine_fwd_kernel(const float* __restrict__ u, const float* __restrict__ ut,
                          const int64_t* __restrict__ slc, __nv_bfloat16* __restrict__ a, int M,
                          int nc, int nv, int d, int thw, int k) {
  pdl_prologue();
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  const int b = (int)(m / thw), p = (int)(m % thw);
  for (int c0 = lane * 4; c0 < d; c0 += 128) {
    float4 acc = ld4(u + (size_t)m * d + c0);
    for (int j = 0; j < k; ++j) {
      const long long code = slc[((size_t)b * nc + j) * thw + p];
      const float4 w4 = ld4(ut + ((size_t)j * nv + code) * d + c0);
      acc.x += w4.x; acc.y += w4.y; acc.z += w4.z; acc.w += w4.w;
    }
    st_bf16x4(a + (size_t)m * d + c0, fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f),
              fmaxf(acc.w, 0.f));
  }
}

__global__ void __launch_bounds__(256)
chpred_combine_bwd_kernel(const __nv_bfloat16* __restrict__ du, const int64_t* __restrict__ slc,
                          float* __restrict__ dut, int M, int nc, int nv, int d, int thw, int k) {
  pdl_prologue();
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  const int b = (int)(m / thw), p = (int)(m % thw);
  for (int c0 = lane * 4; c0 < d; c0 += 128) {
    const float4 g = ld_bf16x4(du + (size_t)m * d + c0);
    for (int j = 0; j < k; ++j) {
      const long long code = slc[((size_t)b * nc + j) * thw + p];
      red_add_f32x4(dut + ((size_t)j * nv + code) * d + c0, g.x, g.y, g.z, g.w);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Cross-entropy over nv classes with ignore mask (meta_arch/vt.py:305-312):
//   loss = 1/nc * sum_k mean_{valid rows} CE(logits_k[row], slice[b,k,pos])
// cnt[0] = number of valid rows (the ignore mask is shared by all channels).
// ------------------------------------------------------------------------------------------
__global__ void count_valid_kernel(const uint8_t* __restrict__ ignore, int n, int* __restrict__ cnt) {
  pdl_prologue();
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    c += ignore[i] ? 0 : 1;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(cnt, c);
}

template <int V4>  // nv = 128*V4
__global__ void __launch_bounds__(256)
cross_entropy_kernel(const float* __restrict__ logits, const int64_t* __restrict__ slc,
                     const uint8_t* __restrict__ ignore, const int* __restrict__ cnt,
                     __nv_bfloat16* __restrict__ dlogits, float* __restrict__ loss, int M, int nc,
                     int thw) {
  pdl_prologue();
  constexpr int nv = V4 * 128;
  __shared__ float s_loss[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * 8 + warp;  // k*M + m
  float row_loss = 0.f;
  if (wid < (long long)nc * M) {
    const int k = (int)(wid / M);
    const int m = (int)(wid - (long long)k * M);
    const int b = m / thw, p = m - b * thw;
    const float scale = 1.f / ((float)(*cnt) * (float)nc);
    __nv_bfloat16* dl = dlogits ? dlogits + (size_t)wid * nv : nullptr;
    if (ignore[(size_t)b * thw + p]) {
      if (dl) {
#pragma unroll
        for (int i = 0; i < V4; ++i) st_bf16x4(dl + (i * 32 + lane) * 4, 0.f, 0.f, 0.f, 0.f);
      }
    } else {
      const float* lr = logits + (size_t)wid * nv;
      const int tgt = (int)slc[((size_t)b * nc + k) * thw + p];
      float4 v[V4];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < V4; ++i) {
        v[i] = ld4(lr + (i * 32 + lane) * 4);
        mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
      }
      mx = warp_max(mx);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < V4; ++i) {
        v[i].x = __expf(v[i].x - mx); v[i].y = __expf(v[i].y - mx);
        v[i].z = __expf(v[i].z - mx); v[i].w = __expf(v[i].w - mx);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
      s = warp_sum(s);
      const float inv = 1.f / s;
      const float lt = lr[tgt];
      row_loss = (mx + __logf(s) - lt) * scale;
      if (dl) {
#pragma unroll
        for (int i = 0; i < V4; ++i) {
          const int c = (i * 32 + lane) * 4;
          float4 g = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
          if (tgt >= c && tgt < c + 4) (&g.x)[tgt - c] -= 1.f;
          st_bf16x4(dl + c, g.x * scale, g.y * scale, g.z * scale, g.w * scale);
        }
      }
    }
  }
  if (lane == 0) s_loss[warp] = row_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_loss[w];
    if (t != 0.f) atomicAdd(loss, t);
  }
}

// ------------------------------------------------------------------------------------------
// Optimizers (torch.optim semantics; solver/build.py:62-72) over flat fp32 buffers, fused with
// the refresh of the bf16 shadow weights the GEMMs read.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ sq,
               float* __restrict__ buf, __nv_bfloat16* __restrict__ pb, long long n4, float lr,
               float alpha, float momentum, float eps, float gscale) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 P = ld4(p + 4 * i), G = ld4(g + 4 * i), S = ld4(sq + 4 * i), Bf = ld4(buf + 4 * i);
    float* Pp = &P.x; float* Gp = &G.x; float* Sp = &S.x; float* Bp = &Bf.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = Gp[j] * gscale;
      Sp[j] = alpha * Sp[j] + (1.f - alpha) * gj * gj;
      const float avg = sqrtf(Sp[j]) + eps;
      Bp[j] = momentum * Bp[j] + gj / avg;
      Pp[j] -= lr * Bp[j];
    }
    st4(p + 4 * i, P); st4(sq + 4 * i, S); st4(buf + 4 * i, Bf);
    if (pb) st_bf16x4(pb + 4 * i, P.x, P.y, P.z, P.w);
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, __nv_bfloat16* __restrict__ pb, long long n4, float lr,
            float beta1, float beta2, float eps, float bc1, float rsqrt_bc2, float gscale) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 P = ld4(p + 4 * i), G = ld4(g + 4 * i), Mm = ld4(m + 4 * i), V = ld4(v + 4 * i);
    float* Pp = &P.x; float* Gp = &G.x; float* Mp = &Mm.x; float* Vp = &V.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = Gp[j] * gscale;
      Mp[j] = beta1 * Mp[j] + (1.f - beta1) * gj;
      Vp[j] = beta2 * Vp[j] + (1.f - beta2) * gj * gj;
      const float denom = sqrtf(Vp[j]) * rsqrt_bc2 + eps;
      Pp[j] -= (lr / bc1) * (Mp[j] / denom);
    }
    st4(p + 4 * i, P); st4(m + 4 * i, Mm); st4(v + 4 * i, V);
    if (pb) st_bf16x4(pb + 4 * i, P.x, P.y, P.z, P.w);
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n4) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ld4(in + 4 * i);
    st_bf16x4(out + 4 * i, v.x, v.y, v.z, v.w);
  }
}

// out[o(i)] (=|+=) in[s(i)] over a 4-D index space; layout packing of small weight tensors.
struct Permute4 {
  int d[4];
  long long is[4], os[4];
};
template <bool BF16, bool ACC>
__global__ void __launch_bounds__(256)
permute4_kernel(const float* __restrict__ in, void* __restrict__ out, Permute4 P, long long total) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int i3 = (int)(r % P.d[3]); r /= P.d[3];
    const int i2 = (int)(r % P.d[2]); r /= P.d[2];
    const int i1 = (int)(r % P.d[1]); r /= P.d[1];
    const int i0 = (int)r;
    const long long so = i0 * P.is[0] + i1 * P.is[1] + i2 * P.is[2] + i3 * P.is[3];
    const long long oo = i0 * P.os[0] + i1 * P.os[1] + i2 * P.os[2] + i3 * P.os[3];
    const float v = in[so];
    if (BF16) reinterpret_cast<__nv_bfloat16*>(out)[oo] = __float2bfloat16(v);
    else if (ACC) reinterpret_cast<float*>(out)[oo] += v;
    else reinterpret_cast<float*>(out)[oo] = v;
  }
}

// Many small re-layout jobs in ONE launch (the per-step weight packs / gradient folds of the engines are 20-70
// launches of a few microseconds each): block b works on 1024 elements of the job whose block range contains b.
// Class conditioning of VTEncoder (videotransformer.py:29-33,54-57): the class embedding is concatenated to every
// position's de channels before the 1x1x1 projector, i.e. each sample gets the extra bias W[:, de:] . E_class[cls].
// cb[b, o] = sum_i W2[o, i] * emb[cls[b], i]  (W2 = the second de columns of the (d, 2de) projector weight, fp32)
__global__ void __launch_bounds__(256)
class_bias_kernel(const float* __restrict__ w2, long long ldw, const float* __restrict__ emb,
                  const int64_t* __restrict__ cls, float* __restrict__ cb, int B, int d, int de) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (o >= d) return;
  const float* e = emb + (long long)cls[b] * de;
  const float* w = w2 + (long long)o * ldw;
  float acc = 0.f;
  for (int i = 0; i < de; ++i) acc = fmaf(w[i], e[i], acc);
  cb[(long long)b * d + o] = acc;
}
// x[r, :] += cb[r / rows_per_group, :]
__global__ void __launch_bounds__(256)
rows_add_group_bias_kernel(float* __restrict__ x, const float* __restrict__ cb, long long M, int d4, int rows_per_group) {
  const long long total = M * d4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d4;
    const int c = (int)(i - r * d4);
    float4 v = ld4(x + 4 * i);
    const float4 bvec = ld4(cb + ((r / rows_per_group) * d4 + c) * 4);
    v.x += bvec.x; v.y += bvec.y; v.z += bvec.z; v.w += bvec.w;
    st4(x + 4 * i, v);
  }
}
// out[g, c] = sum over the rows of group g of x[g * rows + r, c]  (bf16 in, fp32 out; one block per (64 columns, group))
__global__ void __launch_bounds__(256)
colsum_groups_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int rows, int N) {
  __shared__ float s[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + tx * 2, g = blockIdx.y;
  float a = 0.f, b = 0.f;
  if (c < N) {
    const __nv_bfloat16* xg = x + (size_t)g * rows * N;
    for (int r = ty; r < rows; r += 8) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xg + (size_t)r * N + c));
      a += v.x;
      b += v.y;
    }
  }
  s[ty][tx * 2] = a;
  s[ty][tx * 2 + 1] = b;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s[w][threadIdx.x];
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < N) out[(size_t)g * N + cc] = t;
  }
}
// gradients of the class term from S[b, o] = sum over the sample's positions of d(projector output):
// dW2[o, i] += sum_b S[b, o] * emb[cls[b], i];  demb[cls[b], i] += sum_o S[b, o] * W2[o, i]
__global__ void __launch_bounds__(256)
class_grad_kernel(const float* __restrict__ S, const float* __restrict__ emb, const int64_t* __restrict__ cls,
                  const float* __restrict__ w2, float* __restrict__ dw2, long long ldw, float* __restrict__ demb,
                  int B, int d, int de) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < d * de) {  // one (o, i) of dW2
    const int o = t / de, i = t - o * de;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(S[(long long)b * d + o], emb[(long long)cls[b] * de + i], acc);
    dw2[(long long)o * ldw + i] += acc;
  } else if (t < d * de + B * de) {  // one (b, i) of demb
    const int u = t - d * de, b = u / de, i = u - b * de;
    float acc = 0.f;
    for (int o = 0; o < d; ++o) acc = fmaf(S[(long long)b * d + o], w2[(long long)o * ldw + i], acc);
    atomicAdd(demb + (long long)cls[b] * de + i, acc);  // several samples may share a class
  }
}

// dst[r, :] = src[idx[r], :] for rows of row_bytes (a multiple of 16) bytes: the token re-ordering between raster
// order of the slice grid and block-major order of the general tiled BlockLocalAttention (vt_attention.py:189-200).
__global__ void __launch_bounds__(256)
rows_gather_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const int* __restrict__ idx, int M, int vec) {
  pdl_prologue();
  const long long total = (long long)M * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / vec), c = (int)(i - (long long)r * vec);
    dst[i] = src[(long long)idx[r] * vec + c];
  }
}

__global__ void __launch_bounds__(256)
permute4_batch_kernel(const LvtPermuteJob* __restrict__ jobs, int n_jobs) {
  pdl_prologue();
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {  // last job with first_block <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= (long long)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const LvtPermuteJob J = jobs[lo];
  const long long total = (long long)J.dims[0] * J.dims[1] * J.dims[2] * J.dims[3];
  const long long base = ((long long)blockIdx.x - J.first_block) * 1024;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long i = base + u * 256 + threadIdx.x;
    if (i >= total) break;
    int i0, i1, i2, i3;
    if (total < (1ll << 31)) {  // 32-bit index arithmetic (three 64-bit divisions per element bound this kernel)
      unsigned r = (unsigned)i;
      i3 = (int)(r % (unsigned)J.dims[3]); r /= (unsigned)J.dims[3];
      i2 = (int)(r % (unsigned)J.dims[2]); r /= (unsigned)J.dims[2];
      i1 = (int)(r % (unsigned)J.dims[1]); r /= (unsigned)J.dims[1];
      i0 = (int)r;
    } else {
      long long r = i;
      i3 = (int)(r % J.dims[3]); r /= J.dims[3];
      i2 = (int)(r % J.dims[2]); r /= J.dims[2];
      i1 = (int)(r % J.dims[1]); r /= J.dims[1];
      i0 = (int)r;
    }
    const long long so = i0 * J.in_strides[0] + i1 * J.in_strides[1] + i2 * J.in_strides[2] + i3 * J.in_strides[3];
    const long long oo = i0 * J.out_strides[0] + i1 * J.out_strides[1] + i2 * J.out_strides[2] + i3 * J.out_strides[3];
    const float v = J.in[so];
    if (J.out_is_bf16) reinterpret_cast<__nv_bfloat16*>(J.out)[oo] = __float2bfloat16(v);
    else if (J.accumulate) reinterpret_cast<float*>(J.out)[oo] += v;
    else reinterpret_cast<float*>(J.out)[oo] = v;
  }
}

template <typename F>
int dispatch_v4(int n128, F f) {
  switch (n128) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 4: return f(std::integral_constant<int, 4>());
    case 8: return f(std::integral_constant<int, 8>());
    default: lvt_set_error("unsupported width %d (need 128, 256, 512 or 1024)", n128 * 128); return LVT_ERR_INVALID;
  }
}


// ------------------------------------------------------------------------------------------
// One channel of ChannelPredictor.sample (videotransformer.py:161-185): categorical draw from
// softmax(logits / temp) at ONE position per sequence, written straight into the slice buffer.
// Same arithmetic as torch.multinomial(softmax(.), 1) on CUDA: argmax_i p_i / q_i with q ~ Exp(1) drawn by
// the caller (torch's generator, so both paths consume the same random stream); first index wins ties.
// One block per sequence; pos is read from device memory (CUDA-graph replay with a moving position).
// logit * (1 / temp) is rounded BEFORE the maximum is subtracted (__fmul_rn: no contraction into an FMA, whose exact
// product would leave exp() of the maximum entry at exp(rounding residual) -- harmless at temp 1, overflow / all-zero at
// temp -> 0), as ATen's elementwise division and softmax do.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sample_pixel_kernel(const float* __restrict__ logits, const float* __restrict__ q, const int64_t* __restrict__ pos_ptr,
                    int64_t* __restrict__ slc, int thw, int nv, int nc, int k, float inv_temp, int logit_rows) {
  pdl_prologue();
  __shared__ float s_f[8];
  __shared__ int s_i[8];
  const int b = blockIdx.x;
  const int pos = (int)pos_ptr[0];
  // logits: [B * thw, nv] (full pass, logit_rows == thw) or one row per sequence (incremental pass, logit_rows == 1)
  const float* row = logits + ((long long)b * logit_rows + (logit_rows > 1 ? pos : 0)) * nv;
  const float* qr = q + (long long)b * nv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < nv; i += 256) mx = fmaxf(mx, __fmul_rn(row[i], inv_temp));
  mx = warp_max(mx);
  if (lane == 0) s_f[warp] = mx;
  __syncthreads();
  mx = s_f[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_f[w]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < nv; i += 256) sum += expf(__fmul_rn(row[i], inv_temp) - mx);
  sum = warp_sum(sum);
  if (lane == 0) s_f[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += s_f[w];
  __syncthreads();
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int i = threadIdx.x; i < nv; i += 256) {
    const float v = (expf(__fmul_rn(row[i], inv_temp) - mx) / sum) / qr[i];
    if (v > best) { best = v; besti = i; }  // ascending i per thread: the first maximum is kept
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  if (lane == 0) { s_f[warp] = best; s_i[warp] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_f[w] > best || (s_f[w] == best && s_i[w] < besti)) { best = s_f[w]; besti = s_i[w]; }
    slc[((long long)b * nc + k) * thw + pos] = besti;
  }
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int lvt_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16,
                                 float* mean, float* rstd, int M, int d, float eps, void* stream) {
  LVT_CHECK_ARG(x && gamma && beta && y_bf16 && mean && rstd && M > 0, "lvt_layernorm_fwd: bad argument");
  LVT_CHECK_ARG(d % 128 == 0, "lvt_layernorm_fwd: d must be a multiple of 128");
  int rc = dispatch_v4(d / 128, [&](auto v4) {
    LVT_CHECK_CUDA(lvt_launch(ln_fwd_kernel<decltype(v4)::value>, dim3(lvt_ceil_div(M, 8)), dim3(256), 0, STREAM(stream), 
        x, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y_bf16), mean, rstd, M, eps));
    return LVT_OK;
  });
  if (rc) return rc;
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

static int layernorm_bwd_impl(const void* dy, bool dy_bf16, const float* x, const float* mean, const float* rstd,
                              const float* gamma, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma,
                              float* dbeta, float* dxsum, int M, int d, void* stream) {
  LVT_CHECK_ARG(dy && x && mean && rstd && gamma && dgamma && dbeta && M > 0, "lvt_layernorm_bwd: bad argument");
  LVT_CHECK_ARG(d % 128 == 0 && d <= 512, "lvt_layernorm_bwd: d must be 128, 256 or 512");
  const int blocks = min(lvt_ceil_div(M, 8), kSMs * 4);
  int rc = dispatch_v4(d / 128, [&](auto v4) {
    constexpr int V4 = decltype(v4)::value;
    if constexpr (V4 <= 4) {
      if (dy_bf16)
        LVT_CHECK_CUDA(lvt_launch(ln_bwd_kernel<V4, true>, dim3(blocks), dim3(256), 0, STREAM(stream), dy, x, mean, rstd,
                                  gamma, dres, dx_f32, reinterpret_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, dxsum, M));
      else
        LVT_CHECK_CUDA(lvt_launch(ln_bwd_kernel<V4, false>, dim3(blocks), dim3(256), 0, STREAM(stream), dy, x, mean, rstd,
                                  gamma, dres, dx_f32, reinterpret_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, dxsum, M));
    }
    return LVT_OK;
  });
  if (rc) return rc;
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                                 const float* gamma, const float* dres, float* dx_f32, void* dx_bf16,
                                 float* dgamma, float* dbeta, int M, int d, void* stream) {
  return layernorm_bwd_impl(dy, false, x, mean, rstd, gamma, dres, dx_f32, dx_bf16, dgamma, dbeta, nullptr, M, d, stream);
}

extern "C" int lvt_layernorm_bwd_bf16dy(const void* dy_bf16, const float* x, const float* mean, const float* rstd,
                                        const float* gamma, const float* dres, float* dx_f32, void* dx_bf16,
                                        float* dgamma, float* dbeta, int M, int d, void* stream) {
  return layernorm_bwd_impl(dy_bf16, true, x, mean, rstd, gamma, dres, dx_f32, dx_bf16, dgamma, dbeta, nullptr, M, d, stream);
}

extern "C" int lvt_layernorm_bwd_ex(const void* dy, int dy_is_bf16, const float* x, const float* mean, const float* rstd,
                                    const float* gamma, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma,
                                    float* dbeta, float* dx_colsum, int M, int d, void* stream) {
  return layernorm_bwd_impl(dy, dy_is_bf16 != 0, x, mean, rstd, gamma, dres, dx_f32, dx_bf16, dgamma, dbeta, dx_colsum, M, d,
                            stream);
}

extern "C" int lvt_colsum_bf16(const void* x, float* out, int M, int N, long long ld, void* stream) {
  LVT_CHECK_ARG(x && out && M > 0 && N > 0 && N % 2 == 0 && ld % 2 == 0, "lvt_colsum_bf16: bad argument");
  const int strips = lvt_ceil_div(N, 64);
  int ysplit = max(1, min(lvt_ceil_div(M, 64), (kSMs * 4) / strips));
  const int rows_per_block = lvt_ceil_div(M, ysplit);
  ysplit = lvt_ceil_div(M, rows_per_block);
  LVT_CHECK_CUDA(lvt_launch(colsum_bf16_kernel, dim3(dim3(strips, ysplit)), dim3(256), 0, STREAM(stream), 
      reinterpret_cast<const __nv_bfloat16*>(x), out, M, N, ld, rows_per_block));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_attn_delta(const void* dO, const void* O, float* delta, int nb, int H, int L, int da,
                              void* stream) {
  LVT_CHECK_ARG(dO && O && delta && nb > 0 && da % 4 == 0, "lvt_attn_delta: bad argument");
  const long long warps = (long long)nb * L * H;
  LVT_CHECK_CUDA(lvt_launch(attn_delta_kernel, dim3(lvt_ceil_div(warps, 8)), dim3(256), 0, STREAM(stream), 
      reinterpret_cast<const __nv_bfloat16*>(dO), reinterpret_cast<const __nv_bfloat16*>(O), delta, nb, H, L, da));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_relpos_bank_grad(const void* dS, float* dbank_t, float* dbank_h, float* dbank_w,
                                    int nb, int H, int bt, int bh, int bw, void* stream) {
  LVT_CHECK_ARG(dS && dbank_t && dbank_h && dbank_w && nb > 0, "lvt_relpos_bank_grad: bad argument");
  LVT_CHECK_ARG(bt * bh * bw == 256 && 2 * (bt + bh + bw) - 3 <= 64, "lvt_relpos_bank_grad: block must hold 256 positions and <= 64 offsets");
  const int zsplit = nb >= 16 ? 4 : 1;
  LVT_CHECK_CUDA(lvt_launch(relpos_bank_grad_kernel, dim3(dim3(32, H, zsplit)), dim3(256), 0, STREAM(stream), 
      reinterpret_cast<const __nv_bfloat16*>(dS), dbank_t, dbank_h, dbank_w, nb, H, bt, bh, bw));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

static int enc_dims(EncFrontDims* D, int B, int nc, int nv, int de, const int* ctx_shape, const int* kernel,
                    const int* stride, int pad_value) {
  D->B = B; D->nc = nc; D->nv = nv; D->de = de;
  D->Tc = ctx_shape[0]; D->Hc = ctx_shape[1]; D->Wc = ctx_shape[2];
  D->kt = kernel[0]; D->kh = kernel[1]; D->kw = kernel[2];
  D->st = stride[0]; D->sh = stride[1]; D->sw = stride[2];
  D->to = (D->Tc - D->kt) / D->st + 1; D->ho = (D->Hc - D->kh) / D->sh + 1; D->wo = (D->Wc - D->kw) / D->sw + 1;
  D->pad_value = pad_value;
  LVT_CHECK_ARG(B > 0 && de % 4 == 0 && D->to > 0 && D->ho > 0 && D->wo > 0, "vt_enc_front: bad shape");
  return LVT_OK;
}

extern "C" int lvt_vt_enc_front_fwd(const int64_t* context, const int64_t* slice_idx, const float* wt,
                                    const float* bias, const float* slice_emb, void* out_bf16, int B,
                                    int nc, int nv, int de, const int* ctx_shape, const int* kernel,
                                    const int* stride, int pad_value, void* stream) {
  LVT_CHECK_ARG(context && slice_idx && wt && bias && slice_emb && out_bf16, "lvt_vt_enc_front_fwd: null pointer");
  EncFrontDims D;
  int rc = enc_dims(&D, B, nc, nv, de, ctx_shape, kernel, stride, pad_value);
  if (rc) return rc;
  const long long npos = (long long)B * D.to * D.ho * D.wo;
  LVT_CHECK_CUDA(lvt_launch(enc_front_fwd_kernel, dim3(lvt_ceil_div(npos, 8)), dim3(256), 0, STREAM(stream), 
      context, slice_idx, wt, bias, slice_emb, reinterpret_cast<__nv_bfloat16*>(out_bf16), D));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vt_enc_front_bwd(const int64_t* context, const int64_t* slice_idx, const float* dout,
                                    float* dwt, float* dslice_emb, int B, int nc, int nv, int de,
                                    const int* ctx_shape, const int* kernel, const int* stride,
                                    int pad_value, void* stream) {
  LVT_CHECK_ARG(context && slice_idx && dout && dwt && dslice_emb, "lvt_vt_enc_front_bwd: null pointer");
  EncFrontDims D;
  int rc = enc_dims(&D, B, nc, nv, de, ctx_shape, kernel, stride, pad_value);
  if (rc) return rc;
  const long long npos = (long long)B * D.to * D.ho * D.wo;
  LVT_CHECK_CUDA(lvt_launch(enc_front_bwd_kernel, dim3(lvt_ceil_div(npos, 8)), dim3(256), 0, STREAM(stream), context, slice_idx, dout, dwt,
                                                                        dslice_emb, D));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vt_dec_front_fwd(const int64_t* slc, const float* emb, const int* taps, void* out_bf16,
                                    int B, int nc, int nv, int de, int t, int h, int w, int ntaps,
                                    void* stream) {
  LVT_CHECK_ARG(slc && emb && taps && out_bf16 && B > 0 && ntaps > 0 && de % 4 == 0, "lvt_vt_dec_front_fwd: bad argument");
  DecFrontDims D{B, nc, nv, de, t, h, w, ntaps};
  const long long warps = (long long)B * t * h * w * ntaps;
  LVT_CHECK_CUDA(lvt_launch(dec_front_fwd_kernel, dim3(lvt_ceil_div(warps, 8)), dim3(256), 0, STREAM(stream), 
      slc, emb, taps, reinterpret_cast<__nv_bfloat16*>(out_bf16), D));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vt_dec_front_bwd(const int64_t* slc, const float* dA, const int* taps, float* demb,
                                    int B, int nc, int nv, int de, int t, int h, int w, int ntaps,
                                    void* stream) {
  LVT_CHECK_ARG(slc && dA && taps && demb && B > 0 && ntaps > 0 && de % 4 == 0, "lvt_vt_dec_front_bwd: bad argument");
  DecFrontDims D{B, nc, nv, de, t, h, w, ntaps};
  LVT_CHECK_CUDA(lvt_launch(dec_front_bwd_kernel, dim3(lvt_ceil_div((long long)B * t * h * w, 8)), dim3(256), 0, STREAM(stream), slc, dA, taps, demb, D));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_chpred_combine_fwd(const float* u, const float* ut, const int64_t* slc, void* a_bf16,
                                      int M, int nc, int nv, int d, int thw, int k, void* stream) {
  LVT_CHECK_ARG(u && slc && a_bf16 && (k == 0 || ut) && M > 0 && d % 4 == 0 && k >= 0 && k < nc, "lvt_chpred_combine_fwd: bad argument");
  LVT_CHECK_CUDA(lvt_launch(chpred_combine_fwd_kernel, dim3(lvt_ceil_div(M, 8)), dim3(256), 0, STREAM(stream), 
      u, ut, slc, reinterpret_cast<__nv_bfloat16*>(a_bf16), M, nc, nv, d, thw, k));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_chpred_combine_bwd(const void* du_bf16, const int64_t* slc, float* dut, int M, int nc,
                                      int nv, int d, int thw, int k, void* stream) {
  LVT_CHECK_ARG(du_bf16 && slc && M > 0 && d % 4 == 0 && k >= 0 && k < nc, "lvt_chpred_combine_bwd: bad argument");
  if (k == 0) return LVT_OK;
  LVT_CHECK_ARG(dut, "lvt_chpred_combine_bwd: null dut");
  LVT_CHECK_CUDA(lvt_launch(chpred_combine_bwd_kernel, dim3(lvt_ceil_div(M, 8)), dim3(256), 0, STREAM(stream), 
      reinterpret_cast<const __nv_bfloat16*>(du_bf16), slc, dut, M, nc, nv, d, thw, k));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_cross_entropy(const float* logits, const int64_t* slc, const uint8_t* ignore,
                                 void* dlogits_bf16, float* loss, int* count_scratch, int B, int nc,
                                 int nv, int thw, void* stream) {
  LVT_CHECK_ARG(logits && slc && ignore && loss && count_scratch && B > 0, "lvt_cross_entropy: bad argument");
  LVT_CHECK_ARG(nv % 128 == 0, "lvt_cross_entropy: nv must be a multiple of 128");
  const int M = B * thw;
  LVT_CHECK_CUDA(cudaMemsetAsync(count_scratch, 0, sizeof(int), STREAM(stream)));
  LVT_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), STREAM(stream)));
  LVT_CHECK_CUDA(lvt_launch(count_valid_kernel, dim3(min(lvt_ceil_div(M, 256), kSMs)), dim3(256), 0, STREAM(stream), ignore, M, count_scratch));
  int rc = dispatch_v4(nv / 128, [&](auto v4) {
    LVT_CHECK_CUDA(lvt_launch(cross_entropy_kernel<decltype(v4)::value>, dim3(lvt_ceil_div((long long)nc * M, 8)), dim3(256), 0, STREAM(stream), 
        logits, slc, ignore, count_scratch, reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), loss, M, nc, thw));
    return LVT_OK;
  });
  if (rc) return rc;
  LVT_CHECK_LAUNCH();
  lvt_count_launch(2);
  return LVT_OK;
}

static int flat_grid(long long n4) { return (int)min((long long)kSMs * 8, (n4 + 255) / 256); }

extern "C" int lvt_rmsprop_step(float* p, const float* g, float* sq, float* buf, void* p_bf16, long long n,
                                float lr, float alpha, float momentum, float eps, float grad_scale,
                                void* stream) {
  LVT_CHECK_ARG(p && g && sq && buf && n > 0 && n % 4 == 0, "lvt_rmsprop_step: bad argument (n must be a multiple of 4)");
  LVT_CHECK_CUDA(lvt_launch(rmsprop_kernel, dim3(flat_grid(n / 4)), dim3(256), 0, STREAM(stream), p, g, sq, buf, reinterpret_cast<__nv_bfloat16*>(p_bf16),
                                                              n / 4, lr, alpha, momentum, eps, grad_scale));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr,
                             float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  LVT_CHECK_ARG(p && g && m && v && n > 0 && n % 4 == 0 && step >= 1, "lvt_adam_step: bad argument");
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  LVT_CHECK_CUDA(lvt_launch(adam_kernel, dim3(flat_grid(n / 4)), dim3(256), 0, STREAM(stream), p, g, m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16), n / 4,
                                                           lr, beta1, beta2, eps, (float)bc1,
                                                           (float)(1.0 / sqrt(bc2)), grad_scale));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_cast_bf16(const float* in, void* out_bf16, long long n, void* stream) {
  LVT_CHECK_ARG(in && out_bf16 && n > 0 && n % 4 == 0, "lvt_cast_bf16: bad argument");
  LVT_CHECK_CUDA(lvt_launch(cast_bf16_kernel, dim3(flat_grid(n / 4)), dim3(256), 0, STREAM(stream), in, reinterpret_cast<__nv_bfloat16*>(out_bf16), n / 4));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_permute4(const float* in, void* out, int out_is_bf16, int accumulate, const int* dims,
                            const long long* in_strides, const long long* out_strides, void* stream) {
  LVT_CHECK_ARG(in && out && dims && in_strides && out_strides, "lvt_permute4: null pointer");
  LVT_CHECK_ARG(!(out_is_bf16 && accumulate), "lvt_permute4: accumulate needs fp32 output");
  Permute4 P;
  long long total = 1;
  for (int i = 0; i < 4; ++i) {
    P.d[i] = dims[i]; P.is[i] = in_strides[i]; P.os[i] = out_strides[i];
    total *= dims[i];
  }
  LVT_CHECK_ARG(total > 0, "lvt_permute4: empty");
  const int grid = (int)min((long long)kSMs * 8, (total + 255) / 256);
  if (out_is_bf16) LVT_CHECK_CUDA(lvt_launch(permute4_kernel<true, false>, dim3(grid), dim3(256), 0, STREAM(stream), in, out, P, total));
  else if (accumulate) LVT_CHECK_CUDA(lvt_launch(permute4_kernel<false, true>, dim3(grid), dim3(256), 0, STREAM(stream), in, out, P, total));
  else LVT_CHECK_CUDA(lvt_launch(permute4_kernel<false, false>, dim3(grid), dim3(256), 0, STREAM(stream), in, out, P, total));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_permute4_batch(const LvtPermuteJob* jobs_dev, int n_jobs, int total_blocks, void* stream) {
  LVT_CHECK_ARG(jobs_dev && n_jobs > 0 && total_blocks > 0, "lvt_permute4_batch: bad argument");
  LVT_CHECK_CUDA(lvt_launch(permute4_batch_kernel, dim3(total_blocks), dim3(256), 0, STREAM(stream), jobs_dev, n_jobs));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vt_sample_pixel(const float* logits, const float* q_exp, const int64_t* pos, int64_t* slice, int B,
                                   int thw, int nv, int nc, int k, float temp, int logit_rows, void* stream) {
  LVT_CHECK_ARG(logits && q_exp && pos && slice && B > 0 && thw > 0 && nv > 0 && k >= 0 && k < nc && temp > 0.f &&
                    (logit_rows == thw || logit_rows == 1),
                "lvt_vt_sample_pixel: bad argument");
  LVT_CHECK_CUDA(lvt_launch(sample_pixel_kernel, dim3(B), dim3(256), 0, STREAM(stream), logits, q_exp, pos, slice, thw, nv, nc, k,
                            1.f / temp, logit_rows));
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_rows_gather(const void* src, void* dst, const int* idx, int M, int row_bytes, void* stream) {
  LVT_CHECK_ARG(src && dst && idx && M > 0 && row_bytes > 0 && row_bytes % 16 == 0 && src != dst, "lvt_rows_gather: bad argument");
  const int vec = row_bytes / 16;
  LVT_CHECK_CUDA(lvt_launch(rows_gather_kernel, dim3(flat_grid((long long)M * vec)), dim3(256), 0, STREAM(stream),
                            reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), idx, M, vec));
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vt_class_bias(const float* w2, long long ldw, const float* emb, const int64_t* cls, float* cb, int B,
                                 int d, int de, void* stream) {
  LVT_CHECK_ARG(w2 && emb && cls && cb && B > 0 && d > 0 && de > 0, "lvt_vt_class_bias: bad argument");
  class_bias_kernel<<<dim3(lvt_ceil_div(d, 256), B), 256, 0, STREAM(stream)>>>(w2, ldw, emb, cls, cb, B, d, de);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_rows_add_group_bias(float* x, const float* cb, long long M, int d, int rows_per_group, void* stream) {
  LVT_CHECK_ARG(x && cb && M > 0 && d % 4 == 0 && rows_per_group > 0 && M % rows_per_group == 0,
                "lvt_rows_add_group_bias: bad argument");
  rows_add_group_bias_kernel<<<flat_grid(M * (d / 4)), 256, 0, STREAM(stream)>>>(x, cb, M, d / 4, rows_per_group);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_colsum_groups_bf16(const void* x, float* out, int groups, int rows, int N, void* stream) {
  LVT_CHECK_ARG(x && out && groups > 0 && rows > 0 && N > 0 && N % 2 == 0, "lvt_colsum_groups_bf16: bad argument");
  colsum_groups_bf16_kernel<<<dim3(lvt_ceil_div(N, 64), groups), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), out, rows, N);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_vt_class_grad(const float* S, const float* emb, const int64_t* cls, const float* w2, float* dw2,
                                 long long ldw, float* demb, int B, int d, int de, void* stream) {
  LVT_CHECK_ARG(S && emb && cls && w2 && dw2 && demb && B > 0 && d > 0 && de > 0, "lvt_vt_class_grad: bad argument");
  class_grad_kernel<<<lvt_ceil_div(d * de + B * de, 256), 256, 0, STREAM(stream)>>>(S, emb, cls, w2, dw2, ldw, demb, B, d, de);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}
