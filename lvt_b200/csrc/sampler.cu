// lvt_b200 :: incremental (K/V-cached) decoding step of the DSFVT sampler.
//
// Sampling draws one latent position at a time (VideoTransformerModel.sample_video, meta_arch/vt.py:107-134); the
// reference recomputes the whole 256-token decoder pass for every position.  Position p only needs row p of every
// decoder layer plus the keys / values of rows <= p, so the per-position step becomes a chain of skinny
// (B rows x N) products against bf16 weights.  These are weight-streaming problems (M = B <= 16 rows), not tensor
// core work: the kernels below spread the N x K weight read over many CTAs and keep the rows in shared memory.
// Activations that are bf16 in the full pass (LayerNorm outputs, q/k/v, P, O, the FFN hidden) are rounded to bf16 at
// the same places, so the incremental logits differ from the full pass only by summation order.
// The current position is read from device memory (one CUDA graph replayed per position).
#include <stdint.h>

#include "../../include/lvt_b200.h"
#include "common.cuh"

extern void lvt_count_launch(int n);

namespace {

constexpr int RB = 16;  // max rows (sequences) per call

LVT_DEVICE_INLINE float round_bf16(float v) { return __bfloat162float(__float2bfloat16(v)); }

// y[b, n] = epi( sum_k in(b, k) * W[n, k] ),  in = [LayerNorm](x) [rounded to bf16];  one warp per output column.
__global__ void __launch_bounds__(256)
rows_linear_kernel(const LvtRowsLinear a) {
  extern __shared__ float xs[];  // [B][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pos = a.pos ? a.pos[0] : 0;
  const int B = a.B, K = a.K;
  for (int i = threadIdx.x; i < B * K; i += 256) {  // all threads: the B rows -> shared memory (coalesced)
    const int b = i / K, k = i - b * K;
    const long long off = b * a.x_ldb + pos * a.x_pos_mul + k;
    xs[i] = a.x_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.x)[off])
                     : reinterpret_cast<const float*>(a.x)[off];
  }
  __syncthreads();
  if (a.ln_gamma || a.round_in) {
    for (int b = warp; b < B; b += 8) {  // one warp per row: LayerNorm / rounding in shared memory
      float* row = xs + b * K;
      if (a.ln_gamma) {
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s += row[k];
        const float mean = warp_sum(s) / K;
        float q = 0.f;
        for (int k = lane; k < K; k += 32) {
          const float d = row[k] - mean;
          q += d * d;
        }
        const float rstd = rsqrtf(warp_sum(q) / K + a.ln_eps);
        for (int k = lane; k < K; k += 32) row[k] = (row[k] - mean) * rstd * a.ln_gamma[k] + a.ln_beta[k];
      }
      if (a.round_in)
        for (int k = lane; k < K; k += 32) row[k] = round_bf16(row[k]);
    }
  }
  __syncthreads();
  const int n = blockIdx.x * 8 + warp;
  if (n >= a.N) return;
  float acc[RB];
#pragma unroll
  for (int b = 0; b < RB; ++b) acc[b] = 0.f;
  const __nv_bfloat162* wrow = reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(a.w_bf16) + n * a.w_ld);
#pragma unroll 4
  for (int k2 = lane; k2 < K / 2; k2 += 32) {  // (independent weight loads: several in flight)
    const float2 w = __bfloat1622float2(wrow[k2]);
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      if (b < B) {
        const float2 xv = *reinterpret_cast<const float2*>(xs + b * K + 2 * k2);
        acc[b] = fmaf(xv.x, w.x, acc[b]);
        acc[b] = fmaf(xv.y, w.y, acc[b]);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < RB; ++b)
    if (b < B) acc[b] = warp_sum(acc[b]);
  if (lane < B) {
    const int b = lane;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < RB; ++i)
      if (i == b) v = acc[i];
    if (a.bias) v += a.bias[n];
    if (a.res) v += a.res[b * a.res_ldb + pos * a.res_pos_mul + n];
    for (int j = 0; j < a.g_count; ++j) {  // one-hot half of ChannelPredictor.U[k]: a row gather per earlier channel
      const long long code = a.slice[((long long)b * a.nc + j) * a.thw + pos];
      v += a.gtab[((long long)j * a.nv + code) * a.N + n];
    }
    if (a.relu) v = fmaxf(v, 0.f);
    if (a.round_out) v = round_bf16(v);
    a.out[b * a.out_ldb + n] = v;
  }
}

// q | k | v of row `pos`: out column (hb, j) = sum_k LN(x)[b, k] * W[hb][k][j], W as the reference stores w_q / w_k / w_v
// ((head, d, da): x @ w).  One CTA per head block hb in [0, 3H); 8 k-slices x 64 threads x 2 columns per CTA.
__global__ void __launch_bounds__(512)
rows_qkv_kernel(const float* __restrict__ x, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps,
                const __nv_bfloat16* __restrict__ w, float* __restrict__ q_out, __nv_bfloat16* __restrict__ k_cache,
                __nv_bfloat16* __restrict__ v_cache, const int64_t* __restrict__ pos_ptr, int B, int H, int d, int L) {
  constexpr int DA = 128, KS = 8;
  extern __shared__ float sh[];  // [B][d] rows, then [KS][B][128] partial sums
  float* xs = sh;
  float* part = sh + B * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pos = (int)pos_ptr[0];
  for (int i = threadIdx.x; i < B * d; i += 512) xs[i] = x[i];
  __syncthreads();
  for (int b = warp; b < B; b += 16) {
    float* row = xs + b * d;
    float s = 0.f;
    for (int k = lane; k < d; k += 32) s += row[k];
    const float mean = warp_sum(s) / d;
    float q = 0.f;
    for (int k = lane; k < d; k += 32) {
      const float dv = row[k] - mean;
      q += dv * dv;
    }
    const float rstd = rsqrtf(warp_sum(q) / d + eps);
    for (int k = lane; k < d; k += 32) row[k] = round_bf16((row[k] - mean) * rstd * ln_g[k] + ln_b[k]);
  }
  __syncthreads();
  const int hb = blockIdx.x;
  const int j2 = threadIdx.x & 63, ks = threadIdx.x >> 6;  // column pair, k-slice
  const int kper = d / KS;
  const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(w + ((long long)hb * d + ks * kper) * DA) + j2;
  float acc0[RB], acc1[RB];
#pragma unroll
  for (int b = 0; b < RB; ++b) acc0[b] = acc1[b] = 0.f;
  for (int k0 = 0; k0 < kper; k0 += 8) {  // 8 independent 4-byte weight loads in flight, then the FMAs
    float2 wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) wv[u] = __bfloat1622float2(wp[(long long)(k0 + u) * (DA / 2)]);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int b = 0; b < RB; ++b) {
        if (b < B) {
          const float xv = xs[b * d + ks * kper + k0 + u];
          acc0[b] = fmaf(xv, wv[u].x, acc0[b]);
          acc1[b] = fmaf(xv, wv[u].y, acc1[b]);
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < RB; ++b) {
    if (b < B) {
      part[(ks * B + b) * DA + 2 * j2] = acc0[b];
      part[(ks * B + b) * DA + 2 * j2 + 1] = acc1[b];
    }
  }
  __syncthreads();
  const int which = hb / H, head = hb % H;
  for (int i = threadIdx.x; i < B * DA; i += 512) {
    const int b = i / DA, j = i - b * DA;
    float v = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < KS; ++s2) v += part[(s2 * B + b) * DA + j];
    v = round_bf16(v);
    if (which == 0) q_out[((long long)b * H + head) * DA + j] = v;
    else (which == 1 ? k_cache : v_cache)[(((long long)b * H + head) * L + pos) * DA + j] = __float2bfloat16(v);
  }
}

// One query row against the cached keys / values of rows <= pos (ScaledDotProductAttention with the causal mask and
// the relative-position bias of BlockLocalAttention.get_B, vt_attention.py:61-81,169-174).  One CTA per (sequence, head).
__global__ void __launch_bounds__(256)
attn_row_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ k_cache,
                const __nv_bfloat16* __restrict__ v_cache, const float* __restrict__ bank_t,
                const float* __restrict__ bank_h, const float* __restrict__ bank_w, int bt, int bh, int bw,
                const int64_t* __restrict__ pos_ptr, float scale, float* __restrict__ o, int H, int L) {
  constexpr int DA = 128;
  __shared__ float qs[DA];
  __shared__ float ps[256];
  __shared__ float red[8];
  const int bhid = blockIdx.x, head = bhid % H;
  const int pos = (int)pos_ptr[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < DA) qs[threadIdx.x] = q[(long long)bhid * DA + threadIdx.x];
  __syncthreads();
  const int j = threadIdx.x;
  float s = -1e4f;  // masked_fill(-1e4) of the keys after the query (their exp underflows to 0, as in the reference)
  if (j < L && j <= pos) {
    const uint4* kr = reinterpret_cast<const uint4*>(k_cache + ((long long)bhid * L + j) * DA);
    float acc = 0.f;
#pragma unroll 4
    for (int c = 0; c < DA / 8; ++c) {
      const uint4 u = kr[c];
      const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv[e]));
        acc = fmaf(qs[8 * c + 2 * e], kv.x, acc);
        acc = fmaf(qs[8 * c + 2 * e + 1], kv.y, acc);
      }
    }
    const int ti = pos / (bh * bw), hi = (pos / bw) % bh, wi = pos % bw;
    const int tj = j / (bh * bw), hj = (j / bw) % bh, wj = j % bw;
    s = acc * scale + bank_t[head * (2 * bt - 1) + ti - tj + bt - 1] + bank_h[head * (2 * bh - 1) + hi - hj + bh - 1] +
        bank_w[head * (2 * bw - 1) + wi - wj + bw - 1];
  }
  float mx = warp_max(s);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  const float e = (j < L) ? expf(s - mx) : 0.f;
  float sum = warp_sum(e);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  ps[j] = round_bf16(e / sum);
  __syncthreads();
  {
    // O = P V: four quarters of the key range x 64 threads x 2 output columns, 8 independent loads in flight
    __shared__ float osum[4][DA];
    const int d2 = threadIdx.x & 63, jq = threadIdx.x >> 6;
    const __nv_bfloat162* vc = reinterpret_cast<const __nv_bfloat162*>(v_cache + (long long)bhid * L * DA) + d2;
    const int n = min(pos + 1, L);
    const int per = (n + 3) >> 2;
    const int j0 = min(jq * per, n), j1 = min(j0 + per, n);
    float acc0 = 0.f, acc1 = 0.f;
    int jj = j0;
    for (; jj + 8 <= j1; jj += 8) {
      float2 vv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) vv[u] = __bfloat1622float2(vc[(long long)(jj + u) * (DA / 2)]);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc0 = fmaf(ps[jj + u], vv[u].x, acc0);
        acc1 = fmaf(ps[jj + u], vv[u].y, acc1);
      }
    }
    for (; jj < j1; ++jj) {
      const float2 vv = __bfloat1622float2(vc[(long long)jj * (DA / 2)]);
      acc0 = fmaf(ps[jj], vv.x, acc0);
      acc1 = fmaf(ps[jj], vv.y, acc1);
    }
    osum[jq][2 * d2] = acc0;
    osum[jq][2 * d2 + 1] = acc1;
    __syncthreads();
    if (threadIdx.x < DA) {
      const int dcol = threadIdx.x;
      o[(long long)bhid * DA + dcol] = round_bf16((osum[0][dcol] + osum[1][dcol]) + (osum[2][dcol] + osum[3][dcol]));
    }
  }
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int lvt_rows_linear(const LvtRowsLinear* a, void* stream) {
  LVT_CHECK_ARG(a && a->x && a->w_bf16 && a->out && a->B > 0 && a->B <= RB && a->N > 0 && a->K > 0 && a->K % 64 == 0 &&
                    a->w_ld % 2 == 0,
                "lvt_rows_linear: bad argument (1 <= B <= 16, K %% 64 == 0)");
  LVT_CHECK_ARG(a->g_count == 0 || (a->gtab && a->slice && a->pos), "lvt_rows_linear: gather needs gtab, slice and pos");
  LVT_CHECK_ARG((a->x_pos_mul == 0 && a->res_pos_mul == 0) || a->pos, "lvt_rows_linear: position offsets need pos");
  const size_t smem = (size_t)a->B * a->K * sizeof(float);
  LVT_CHECK_ARG(smem <= 96 * 1024, "lvt_rows_linear: B * K too large");
  static bool configured = false;
  if (!configured) {
    LVT_CHECK_CUDA(cudaFuncSetAttribute(rows_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  rows_linear_kernel<<<lvt_ceil_div(a->N, 8), 256, smem, STREAM(stream)>>>(*a);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_rows_qkv(const float* x, const float* ln_gamma, const float* ln_beta, float eps, const void* w_bf16,
                            float* q_out, void* k_cache_bf16, void* v_cache_bf16, const int64_t* pos, int B, int H, int d,
                            int da, int L, void* stream) {
  LVT_CHECK_ARG(x && ln_gamma && ln_beta && w_bf16 && q_out && k_cache_bf16 && v_cache_bf16 && pos && B > 0 && B <= RB &&
                    H > 0 && da == 128 && d % 64 == 0 && L > 0,
                "lvt_rows_qkv: bad argument (1 <= B <= 16, da == 128)");
  const size_t smem = ((size_t)B * d + 8 * (size_t)B * 128) * sizeof(float);
  LVT_CHECK_ARG(smem <= 96 * 1024, "lvt_rows_qkv: B * d too large");
  static bool configured = false;
  if (!configured) {
    LVT_CHECK_CUDA(cudaFuncSetAttribute(rows_qkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  rows_qkv_kernel<<<3 * H, 512, smem, STREAM(stream)>>>(x, ln_gamma, ln_beta, eps, reinterpret_cast<const __nv_bfloat16*>(w_bf16),
                                                        q_out, reinterpret_cast<__nv_bfloat16*>(k_cache_bf16),
                                                        reinterpret_cast<__nv_bfloat16*>(v_cache_bf16), pos, B, H, d, L);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}

extern "C" int lvt_attn_row(const float* q, const void* k_cache_bf16, const void* v_cache_bf16, const float* bank_t,
                            const float* bank_h, const float* bank_w, int bt, int bh, int bw, const int64_t* pos, float scale,
                            float* o, int B, int H, int L, int da, void* stream) {
  LVT_CHECK_ARG(q && k_cache_bf16 && v_cache_bf16 && bank_t && bank_h && bank_w && pos && o && B > 0 && H > 0 && da == 128 &&
                    L > 0 && L <= 256 && bt * bh * bw == L,
                "lvt_attn_row: bad argument (da == 128, L = bt*bh*bw <= 256)");
  attn_row_kernel<<<B * H, 256, 0, STREAM(stream)>>>(q, reinterpret_cast<const __nv_bfloat16*>(k_cache_bf16),
                                                     reinterpret_cast<const __nv_bfloat16*>(v_cache_bf16), bank_t, bank_h, bank_w, bt,
                                                     bh, bw, pos, scale, o, H, L);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}
