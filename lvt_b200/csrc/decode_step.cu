// lvt_b200 :: the whole per-position step of the incremental DSFVT sampler as ONE persistent kernel.
//
// Sampling draws one latent position at a time (VideoTransformerModel.sample_video, meta_arch/vt.py:107-134 around
// VideoTransformer mode "sample_pixel", videotransformer.py:161-185,240-246).  With the K/V caches of sampler.cu one
// position is a chain of ~58 skinny products (one row per sequence through the masked decoder, then per channel the
// predictor and the categorical draw): as separate launches they cost 5-9 us each, i.e. the sampler was bound by launch
// latency (0.50 ms per position at B = 1).  Here the same stages run inside one kernel of 32 CTAs x 512 threads with a
// grid barrier between dependent stages; the arithmetic of every stage is the arithmetic of the corresponding kernel
// of sampler.cu / ops.cu (same lane -> k mapping, same bf16 rounding points, same reduction order), so the sampled
// codes are identical to the launch-per-stage path.
//   stage 0      row `pos` of the masked-conv im2col (embed-sum of the taps, built in shared memory) x conv weights + y0s
//   per layer    LayerNorm + q | k | v (k, v into the caches)  |  attention of the row against the cached rows  |
//                projection + residual  |  LayerNorm + FFN1 + ReLU  |  FFN2 + residual
//   per channel  LayerNorm + U[k] (+ one-hot row gather) + ReLU  |  P[k] logits  |  argmax(softmax / Exp(1) noise)
// Activations written by one stage and read by the next go through L2 (ld.global.cg: no stale L1 lines).
#include <stdint.h>

#include "../../include/lvt_b200.h"
#include "common.cuh"

extern void lvt_count_launch(int n);

namespace {

constexpr int RB = 16;       // max rows (sequences)
constexpr int NT = 512;      // threads per CTA
constexpr int NW = NT / 32;  // warps per CTA
constexpr int GRID = 32;     // CTAs (co-resident: the kernel runs alone on its stream)

LVT_DEVICE_INLINE float round_bf16(float v) { return __bfloat162float(__float2bfloat16(v)); }

// Barrier over the whole grid: state[0] counts arrivals and only grows inside a launch -- barrier number k completes
// when it reaches k * gridDim.x -- so a CTA needs ONE fire-and-forget red plus the polling loop per barrier (the
// previous generation barrier also read the generation word first and had the last arriver reset / publish: two
// more dependent L2 round trips; 2.5 us per barrier, 52 barriers per position).  state[1] counts the CTAs that have
// left the kernel; the last one puts both words back to zero, so the same two words serve every launch.
LVT_DEVICE_INLINE void grid_sync(unsigned* state, unsigned& k) {
  ++k;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned target = k * gridDim.x;
    __threadfence();  // this CTA's writes of the stage before the barrier
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(state) : "memory");
    unsigned spins = 0, v;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(state) : "memory");
      if (v >= target) break;
      if (++spins > (1u << 26)) {
        printf("lvt_b200: decode step grid barrier timeout (block %d)\n", blockIdx.x);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}
LVT_DEVICE_INLINE void grid_exit(unsigned* state) {
  if (threadIdx.x == 0 && atomicAdd(state + 1, 1u) == gridDim.x - 1) {  // every other CTA is past its last barrier
    state[0] = 0u;
    state[1] = 0u;
    __threadfence();
  }
}

struct RowsArgs {
  const float* x;          // [B, K] fp32 activations (nullptr: the rows are already in shared memory)
  int K, N;
  const float* ln_g; const float* ln_b; float ln_eps;
  int round_in;
  const __nv_bfloat16* w; long long w_ld;
  const float* bias;
  const float* res; long long res_ldb, res_off;
  const float* gtab; int g_count;
  int relu, round_out;
  float* out;
};

// The weights of a stage do not depend on the previous stage: every warp requests the weight row of its (first) output
// column BEFORE the grid barrier, so the L2 round trip overlaps the barrier and the staging of the rows.
constexpr int WPRE = 16;  // bf16x2 words per lane: K <= 1024
LVT_DEVICE_INLINE void rows_prefetch(const RowsArgs& a, uint32_t (&wreg)[WPRE]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * NW + warp;
  if (n >= a.N) return;
  const uint32_t* wrow = reinterpret_cast<const uint32_t*>(a.w + n * a.w_ld);
#pragma unroll
  for (int i = 0; i < WPRE; ++i)
    if (lane + 32 * i < a.K / 2) wreg[i] = __ldg(wrow + lane + 32 * i);
}

// rows_linear_kernel (sampler.cu) with 16 warps per CTA and the output columns strided over the grid
template <int R>
LVT_DEVICE_INLINE void stage_rows(const LvtDecodeStep& p, const RowsArgs& a, float* xs, long long pos,
                                  const uint32_t (&wreg)[WPRE]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = p.B, K = a.K;
  if ((int)blockIdx.x * NW >= a.N) return;  // (uniform per CTA)
  if (a.x) {
    for (int i = threadIdx.x; i < B * K; i += NT) xs[i] = __ldcg(a.x + i);
    __syncthreads();
  }
  if (a.ln_g || a.round_in) {
    for (int b = warp; b < B; b += NW) {
      float* row = xs + b * K;
      if (a.ln_g) {
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s += row[k];
        const float mean = warp_sum(s) / K;
        float q = 0.f;
        for (int k = lane; k < K; k += 32) {
          const float d = row[k] - mean;
          q += d * d;
        }
        const float rstd = rsqrtf(warp_sum(q) / K + a.ln_eps);
        for (int k = lane; k < K; k += 32) row[k] = (row[k] - mean) * rstd * a.ln_g[k] + a.ln_b[k];
      }
      if (a.round_in)
        for (int k = lane; k < K; k += 32) row[k] = round_bf16(row[k]);
    }
    __syncthreads();
  }
  for (int n = blockIdx.x * NW + warp; n < a.N; n += GRID * NW) {
    float acc[R];
#pragma unroll
    for (int b = 0; b < R; ++b) acc[b] = 0.f;
    const __nv_bfloat162* wrow = reinterpret_cast<const __nv_bfloat162*>(a.w + n * a.w_ld);
    if (n < GRID * NW) {  // first column of this warp: its weights were requested before the barrier (same k order)
#pragma unroll
      for (int i = 0; i < WPRE; ++i) {
        const int k2 = lane + 32 * i;
        if (k2 < K / 2) {
          const float2 w = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wreg[i]));
#pragma unroll
          for (int b = 0; b < R; ++b) {
            if (b < B) {
              const float2 xv = *reinterpret_cast<const float2*>(xs + b * K + 2 * k2);
              acc[b] = fmaf(xv.x, w.x, acc[b]);
              acc[b] = fmaf(xv.y, w.y, acc[b]);
            }
          }
        }
      }
      for (int k2 = lane + 32 * WPRE; k2 < K / 2; k2 += 32) {  // K > 1024 (masked conv with many live taps)
        const float2 w = __bfloat1622float2(wrow[k2]);
#pragma unroll
        for (int b = 0; b < R; ++b) {
          if (b < B) {
            const float2 xv = *reinterpret_cast<const float2*>(xs + b * K + 2 * k2);
            acc[b] = fmaf(xv.x, w.x, acc[b]);
            acc[b] = fmaf(xv.y, w.y, acc[b]);
          }
        }
      }
    } else {
#pragma unroll 4
      for (int k2 = lane; k2 < K / 2; k2 += 32) {
        const float2 w = __bfloat1622float2(wrow[k2]);
#pragma unroll
        for (int b = 0; b < R; ++b) {
          if (b < B) {
            const float2 xv = *reinterpret_cast<const float2*>(xs + b * K + 2 * k2);
            acc[b] = fmaf(xv.x, w.x, acc[b]);
            acc[b] = fmaf(xv.y, w.y, acc[b]);
          }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < R; ++b)
      if (b < B) acc[b] = warp_sum(acc[b]);
    if (lane < B) {
      const int b = lane;
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < R; ++i)
        if (i == b) v = acc[i];
      if (a.bias) v += a.bias[n];
      if (a.res) v += __ldcg(a.res + b * a.res_ldb + a.res_off + n);
      for (int j = 0; j < a.g_count; ++j) {
        const long long code = __ldcg(p.slice + ((long long)b * p.nc + j) * p.L + pos);
        v += a.gtab[((long long)j * p.nv + code) * a.N + n];
      }
      if (a.relu) v = fmaxf(v, 0.f);
      if (a.round_out) v = round_bf16(v);
      a.out[b * a.N + n] = v;
    }
  }
}

// first 32 of the 64 weight words a thread of the q | k | v stage reads for its first head block, requested before the barrier
LVT_DEVICE_INLINE void qkv_prefetch(const LvtDecodeStep& p, const LvtDecodeLayer& ly, uint32_t (&wq)[32]) {
  constexpr int DA = 128, KS = 8;
  const int hb = blockIdx.x;
  if (hb >= 3 * p.H) return;
  const int j2 = threadIdx.x & 63, ks = threadIdx.x >> 6;
  const int kper = p.d / KS;
  const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(ly.w_qkv) +
                                                                     ((long long)hb * p.d + ks * kper) * DA) + j2;
#pragma unroll
  for (int u = 0; u < 32; ++u)
    if (u < kper) wq[u] = __ldg(reinterpret_cast<const uint32_t*>(wp + (long long)u * (DA / 2)));
}

// rows_qkv_kernel (sampler.cu): one head block per CTA iteration
template <int R>
LVT_DEVICE_INLINE void stage_qkv(const LvtDecodeStep& p, const LvtDecodeLayer& ly, const float* x, float* sh, int pos,
                                 const uint32_t (&wq)[32]) {
  constexpr int DA = 128, KS = 8;
  const int B = p.B, d = p.d, H = p.H, L = p.L;
  if ((int)blockIdx.x >= 3 * H) return;
  float* xs = sh;
  float* part = sh + B * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < B * d; i += NT) xs[i] = __ldcg(x + i);
  __syncthreads();
  for (int b = warp; b < B; b += NW) {
    float* row = xs + b * d;
    float s = 0.f;
    for (int k = lane; k < d; k += 32) s += row[k];
    const float mean = warp_sum(s) / d;
    float q = 0.f;
    for (int k = lane; k < d; k += 32) {
      const float dv = row[k] - mean;
      q += dv * dv;
    }
    const float rstd = rsqrtf(warp_sum(q) / d + p.ln_eps);
    for (int k = lane; k < d; k += 32) row[k] = round_bf16((row[k] - mean) * rstd * ly.ln1_g[k] + ly.ln1_b[k]);
  }
  __syncthreads();
  const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(ly.w_qkv);
  __nv_bfloat16* k_cache = reinterpret_cast<__nv_bfloat16*>(ly.k_cache);
  __nv_bfloat16* v_cache = reinterpret_cast<__nv_bfloat16*>(ly.v_cache);
  for (int hb = blockIdx.x; hb < 3 * H; hb += GRID) {
    const int j2 = threadIdx.x & 63, ks = threadIdx.x >> 6;
    const int kper = d / KS;
    const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(w + ((long long)hb * d + ks * kper) * DA) + j2;
    float acc0[R], acc1[R];
#pragma unroll
    for (int b = 0; b < R; ++b) acc0[b] = acc1[b] = 0.f;
    for (int k0 = 0; k0 < kper; k0 += 32) {  // 32 independent weight loads in flight (same k order as rows_qkv_kernel)
      uint32_t wr[32];
      if (k0 == 0 && hb == (int)blockIdx.x) {
#pragma unroll
        for (int u = 0; u < 32; ++u) wr[u] = wq[u];
      } else {
#pragma unroll
        for (int u = 0; u < 32; ++u)
          if (k0 + u < kper) wr[u] = __ldg(reinterpret_cast<const uint32_t*>(wp + (long long)(k0 + u) * (DA / 2)));
      }
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        if (k0 + u < kper) {
          const float2 wv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wr[u]));
#pragma unroll
          for (int b = 0; b < R; ++b) {
            if (b < B) {
              const float xv = xs[b * d + ks * kper + k0 + u];
              acc0[b] = fmaf(xv, wv.x, acc0[b]);
              acc1[b] = fmaf(xv, wv.y, acc1[b]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < R; ++b) {
      if (b < B) {
        part[(ks * B + b) * DA + 2 * j2] = acc0[b];
        part[(ks * B + b) * DA + 2 * j2 + 1] = acc1[b];
      }
    }
    __syncthreads();
    const int which = hb / H, head = hb % H;
    for (int i = threadIdx.x; i < B * DA; i += NT) {
      const int b = i / DA, j = i - b * DA;
      float v = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < KS; ++s2) v += part[(s2 * B + b) * DA + j];
      v = round_bf16(v);
      if (which == 0) p.q[((long long)b * H + head) * DA + j] = v;
      else (which == 1 ? k_cache : v_cache)[(((long long)b * H + head) * L + pos) * DA + j] = __float2bfloat16(v);
    }
    __syncthreads();
  }
}

// attn_row_kernel (sampler.cu): one (sequence, head) per CTA iteration; threads >= 256 only keep the barriers company
LVT_DEVICE_INLINE void stage_attn(const LvtDecodeStep& p, const LvtDecodeLayer& ly, int pos) {
  constexpr int DA = 128;
  __shared__ float qs[DA];
  __shared__ float ps[256];
  __shared__ float red[8];
  __shared__ float osum[4][DA];
  const int H = p.H, L = p.L, bt = p.bt, bh = p.bh, bw = p.bw;
  const __nv_bfloat16* k_cache = reinterpret_cast<const __nv_bfloat16*>(ly.k_cache);
  const __nv_bfloat16* v_cache = reinterpret_cast<const __nv_bfloat16*>(ly.v_cache);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool act = threadIdx.x < 256;
  for (int bhid = blockIdx.x; bhid < p.B * H; bhid += GRID) {
    const int head = bhid % H;
    if (threadIdx.x < DA) qs[threadIdx.x] = __ldcg(p.q + (long long)bhid * DA + threadIdx.x);
    __syncthreads();
    const int j = threadIdx.x;
    float s = -1e4f;
    if (act && j < L && j <= pos) {
      const uint4* kr = reinterpret_cast<const uint4*>(k_cache + ((long long)bhid * L + j) * DA);
      float acc = 0.f;
      uint4 ku[DA / 8];
#pragma unroll
      for (int c = 0; c < DA / 8; ++c) ku[c] = __ldcg(kr + c);
#pragma unroll
      for (int c = 0; c < DA / 8; ++c) {
        const uint4 u = ku[c];
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
      s = acc * p.scale + ly.bank_t[head * (2 * bt - 1) + ti - tj + bt - 1] + ly.bank_h[head * (2 * bh - 1) + hi - hj + bh - 1] +
          ly.bank_w[head * (2 * bw - 1) + wi - wj + bw - 1];
    }
    float mx = warp_max(s);
    if (act && lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    const float e = (act && j < L) ? expf(s - mx) : 0.f;
    float sum = warp_sum(e);
    if (act && lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w];
    if (act) ps[j] = round_bf16(e / sum);
    __syncthreads();
    if (act) {
      const int d2 = threadIdx.x & 63, jq = threadIdx.x >> 6;
      const __nv_bfloat162* vc = reinterpret_cast<const __nv_bfloat162*>(v_cache + (long long)bhid * L * DA) + d2;
      const int n = min(pos + 1, L);
      const int per = (n + 3) >> 2;
      const int j0 = min(jq * per, n), j1 = min(j0 + per, n);
      float acc0 = 0.f, acc1 = 0.f;
      int jj = j0;
      for (; jj + 32 <= j1; jj += 32) {  // (sequential accumulation over the keys, as attn_row_kernel)
        uint32_t raw[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) raw[u] = __ldcg(reinterpret_cast<const uint32_t*>(vc + (long long)(jj + u) * (DA / 2)));
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[u]));
          acc0 = fmaf(ps[jj + u], vv.x, acc0);
          acc1 = fmaf(ps[jj + u], vv.y, acc1);
        }
      }
      for (; jj + 8 <= j1; jj += 8) {
        float2 vv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t raw = __ldcg(reinterpret_cast<const uint32_t*>(vc + (long long)(jj + u) * (DA / 2)));
          vv[u] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc0 = fmaf(ps[jj + u], vv[u].x, acc0);
          acc1 = fmaf(ps[jj + u], vv[u].y, acc1);
        }
      }
      for (; jj < j1; ++jj) {
        const uint32_t raw = __ldcg(reinterpret_cast<const uint32_t*>(vc + (long long)jj * (DA / 2)));
        const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
        acc0 = fmaf(ps[jj], vv.x, acc0);
        acc1 = fmaf(ps[jj], vv.y, acc1);
      }
      osum[jq][2 * d2] = acc0;
      osum[jq][2 * d2 + 1] = acc1;
    }
    __syncthreads();
    if (threadIdx.x < DA) {
      const int dcol = threadIdx.x;
      p.o[(long long)bhid * DA + dcol] = round_bf16((osum[0][dcol] + osum[1][dcol]) + (osum[2][dcol] + osum[3][dcol]));
    }
    __syncthreads();
  }
}

// sample_pixel_kernel (ops.cu) on the one-row-per-sequence logits: one sequence per CTA iteration, 256 active threads
LVT_DEVICE_INLINE void stage_sample(const LvtDecodeStep& p, int k, int pos) {
  __shared__ float s_f[8];
  __shared__ int s_i[8];
  const int nv = p.nv;
  const float inv_temp = 1.f / p.temp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool act = threadIdx.x < 256;
  for (int b = blockIdx.x; b < p.B; b += GRID) {
    const float* row = p.logits + (long long)b * nv;
    const float* qr = p.q_exp + ((long long)k * p.B + b) * nv;
    float mx = -INFINITY;
    if (act)
      for (int i = threadIdx.x; i < nv; i += 256) mx = fmaxf(mx, __fmul_rn(__ldcg(row + i), inv_temp));
    mx = warp_max(mx);
    if (act && lane == 0) s_f[warp] = mx;
    __syncthreads();
    mx = s_f[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_f[w]);
    __syncthreads();
    float sum = 0.f;
    if (act)
      for (int i = threadIdx.x; i < nv; i += 256) sum += expf(__fmul_rn(__ldcg(row + i), inv_temp) - mx);
    sum = warp_sum(sum);
    if (act && lane == 0) s_f[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += s_f[w];
    __syncthreads();
    float best = -INFINITY;
    int besti = 0x7fffffff;
    if (act)
      for (int i = threadIdx.x; i < nv; i += 256) {
        const float v = (expf(__fmul_rn(__ldcg(row + i), inv_temp) - mx) / sum) / qr[i];
        if (v > best) { best = v; besti = i; }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if (act && lane == 0) { s_f[warp] = best; s_i[warp] = besti; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w)
        if (s_f[w] > best || (s_f[w] == best && s_i[w] < besti)) { best = s_f[w]; besti = s_i[w]; }
      p.slice[((long long)b * p.nc + k) * p.L + pos] = besti;
    }
    __syncthreads();
  }
}

template <int R>
__global__ void __launch_bounds__(NT, 1)
decode_step_kernel(const __grid_constant__ LvtDecodeStep p) {
  extern __shared__ float sh[];
  const int pos = (int)p.pos[0];
  const int B = p.B, d = p.d;
  const int K0 = p.ntaps * p.de;
  auto bf = [](const void* w) { return reinterpret_cast<const __nv_bfloat16*>(w); };
  auto args_conv = [&]() {
    RowsArgs a = {};
    a.K = K0; a.N = d; a.w = bf(p.conv_w); a.w_ld = K0;
    a.res = p.y0s; a.res_ldb = (long long)p.L * d; a.res_off = (long long)pos * d;
    a.out = p.xa;
    return a;
  };
  auto args_proj = [&](const LvtDecodeLayer& ly, const float* x) {
    RowsArgs a = {};
    a.x = p.o; a.K = p.H * p.da; a.N = d; a.w = bf(ly.w_proj); a.w_ld = p.H * p.da;
    a.res = x; a.res_ldb = d;
    a.out = p.hbuf;
    return a;
  };
  auto args_ffn1 = [&](const LvtDecodeLayer& ly) {
    RowsArgs a = {};
    a.x = p.hbuf; a.K = d; a.N = d;
    a.ln_g = ly.ln2_g; a.ln_b = ly.ln2_b; a.ln_eps = p.ln_eps; a.round_in = 1;
    a.w = bf(ly.w_ffn1); a.w_ld = d; a.bias = ly.b_ffn1; a.relu = 1; a.round_out = 1;
    a.out = p.a1;
    return a;
  };
  auto args_ffn3 = [&](const LvtDecodeLayer& ly, float* y) {
    RowsArgs a = {};
    a.x = p.a1; a.K = d; a.N = d; a.w = bf(ly.w_ffn3); a.w_ld = d; a.bias = ly.b_ffn3;
    a.res = p.hbuf; a.res_ldb = d;
    a.out = y;
    return a;
  };
  auto args_U = [&](int k, const float* x) {
    RowsArgs a = {};
    a.x = x; a.K = d; a.N = d;
    a.ln_g = p.lnp_g; a.ln_b = p.lnp_b; a.ln_eps = p.ln_eps; a.round_in = 1;
    a.w = bf(p.U[k]); a.w_ld = p.U_ld[k]; a.bias = p.U_bias[k];
    a.gtab = p.gtab[k]; a.g_count = k; a.relu = 1; a.round_out = 1;
    a.out = p.abuf;
    return a;
  };
  auto args_P = [&](int k) {
    RowsArgs a = {};
    a.x = p.abuf; a.K = d; a.N = p.nv; a.w = bf(p.P[k]); a.w_ld = d; a.bias = p.P_bias[k];
    a.out = p.logits;
    return a;
  };
  int nstamp = 0;
  auto stamp = [&]() {
    if (p.prof && blockIdx.x == 0 && threadIdx.x == 0 && nstamp < 128) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.prof[nstamp] = t;
    }
    ++nstamp;
  };
  stamp();
  unsigned nbar = 0;    // grid barriers passed so far in this launch
  uint32_t wreg[WPRE];  // weights of the next rows stage, requested before the barrier in front of it
  uint32_t wq[32];      // ... of the next q | k | v stage
  // ---- stage 0: row `pos` of embed-sum + causal im2col (dec_front_fwd_kernel, ops.cu: bf16-rounded), then the conv row
  {
    const RowsArgs a = args_conv();
    rows_prefetch(a, wreg);
    if (p.n_layers > 0) qkv_prefetch(p, p.layer[0], wq);
    const int hw = p.h * p.w;
    if ((int)blockIdx.x * NW < d) {
      for (int i = threadIdx.x; i < B * p.ntaps * (p.de / 4); i += NT) {
        const int c0 = (i % (p.de / 4)) * 4;
        const int q = (i / (p.de / 4)) % p.ntaps;
        const int b = i / ((p.de / 4) * p.ntaps);
        int r = pos;
        const int tt = r / hw + p.taps[3 * q];
        r %= hw;
        const int hh = r / p.w + p.taps[3 * q + 1], ww = r % p.w + p.taps[3 * q + 2];
        const bool inside = tt >= 0 && tt < p.t && hh >= 0 && hh < p.h && ww >= 0 && ww < p.w;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inside) {
          const int pp = (tt * p.h + hh) * p.w + ww;
          for (int k = 0; k < p.nc; ++k) {
            const long long code = __ldcg(p.slice + ((size_t)b * p.nc + k) * p.L + pp);
            const float4 e = *reinterpret_cast<const float4*>(p.emb + ((size_t)k * p.nv + code) * p.de + c0);
            acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
          }
        }
        float* o = sh + (size_t)b * K0 + q * p.de + c0;
        o[0] = round_bf16(acc.x); o[1] = round_bf16(acc.y); o[2] = round_bf16(acc.z); o[3] = round_bf16(acc.w);
      }
      __syncthreads();
    }
    stage_rows<R>(p, a, sh, pos, wreg);
  }
  stamp(); grid_sync(p.barrier, nbar); stamp();
  float* x = p.xa;
  float* y = p.xb;
  for (int i = 0; i < p.n_layers; ++i) {
    const LvtDecodeLayer& ly = p.layer[i];
    stage_qkv<R>(p, ly, x, sh, pos, wq);
    {
      const RowsArgs a = args_proj(ly, x);
      rows_prefetch(a, wreg);  // (stays in registers across the attention stage)
      stamp(); grid_sync(p.barrier, nbar); stamp();
      stage_attn(p, ly, pos);
      stamp(); grid_sync(p.barrier, nbar); stamp();
      stage_rows<R>(p, a, sh, pos, wreg);
    }
    {
      const RowsArgs a = args_ffn1(ly);
      rows_prefetch(a, wreg);
      stamp(); grid_sync(p.barrier, nbar); stamp();
      stage_rows<R>(p, a, sh, pos, wreg);
    }
    {
      const RowsArgs a = args_ffn3(ly, y);
      rows_prefetch(a, wreg);
      stamp(); grid_sync(p.barrier, nbar); stamp();
      stage_rows<R>(p, a, sh, pos, wreg);
    }
    float* t = x; x = y; y = t;
    if (i + 1 < p.n_layers) {
      qkv_prefetch(p, p.layer[i + 1], wq);
    } else if (p.do_sample) {
      const RowsArgs a = args_U(0, x);
      rows_prefetch(a, wreg);
    }
    if (i + 1 < p.n_layers || p.do_sample) stamp(); grid_sync(p.barrier, nbar); stamp();
  }
  if (!p.do_sample) {
    grid_exit(p.barrier);
    return;
  }
  if (p.n_layers == 0) {
    const RowsArgs a = args_U(0, x);
    rows_prefetch(a, wreg);
  }
  for (int k = 0; k < p.nc; ++k) {
    {
      const RowsArgs a = args_U(k, x);
      stage_rows<R>(p, a, sh, pos, wreg);
    }
    {
      const RowsArgs a = args_P(k);
      rows_prefetch(a, wreg);
      stamp(); grid_sync(p.barrier, nbar); stamp();
      stage_rows<R>(p, a, sh, pos, wreg);
    }
    stamp(); grid_sync(p.barrier, nbar); stamp();
    stage_sample(p, k, pos);
    if (k + 1 < p.nc) {
      const RowsArgs a = args_U(k + 1, x);
      rows_prefetch(a, wreg);
      stamp(); grid_sync(p.barrier, nbar); stamp();
    }
  }
  grid_exit(p.barrier);
}

}  // namespace

extern "C" int lvt_vt_decode_step(const LvtDecodeStep* p, void* stream) {
  LVT_CHECK_ARG(p != nullptr, "lvt_vt_decode_step: null descriptor");
  LVT_CHECK_ARG(p->B > 0 && p->B <= RB && p->da == 128 && p->d % 64 == 0 && p->d >= 64 && p->H > 0 && p->L > 0 && p->L <= 256 &&
                    p->bt * p->bh * p->bw == p->L && p->t * p->h * p->w == p->L && p->n_layers >= 0 && p->n_layers <= 8 &&
                    p->nc > 0 && p->nc <= 4 && p->nv > 0 && p->de % 4 == 0 && p->ntaps > 0 && (p->ntaps * p->de) % 64 == 0 &&
                    (p->H * p->da) % 64 == 0 && p->temp > 0.f,
                "lvt_vt_decode_step: bad shape (1 <= B <= 16, da == 128, L <= 256, <= 8 layers, <= 4 channels)");
  LVT_CHECK_ARG(p->pos && p->slice && p->emb && p->taps && p->conv_w && p->y0s && p->xa && p->xb && p->hbuf && p->a1 && p->q &&
                    p->o && p->abuf && p->logits && p->barrier && (!p->do_sample || p->q_exp),
                "lvt_vt_decode_step: null pointer");
  int kmax = p->ntaps * p->de;
  if (p->H * p->da > kmax) kmax = p->H * p->da;
  if (p->d > kmax) kmax = p->d;
  size_t smem = (size_t)p->B * kmax * sizeof(float);
  const size_t smem_qkv = ((size_t)p->B * p->d + 8 * (size_t)p->B * 128) * sizeof(float);
  if (smem_qkv > smem) smem = smem_qkv;
  LVT_CHECK_ARG(smem <= 160 * 1024, "lvt_vt_decode_step: B * K too large for shared memory");
  static bool configured = false;
  if (!configured) {
    LVT_CHECK_CUDA(cudaFuncSetAttribute(decode_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    LVT_CHECK_CUDA(cudaFuncSetAttribute(decode_step_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    LVT_CHECK_CUDA(cudaFuncSetAttribute(decode_step_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // (row capacity as a template parameter: the per-row accumulators are registers)
  if (p->B == 1) decode_step_kernel<1><<<GRID, NT, smem, st>>>(*p);
  else if (p->B <= 4) decode_step_kernel<4><<<GRID, NT, smem, st>>>(*p);
  else decode_step_kernel<RB><<<GRID, NT, smem, st>>>(*p);
  LVT_CHECK_LAUNCH();
  lvt_count_launch(1);
  return LVT_OK;
}
