/* lvt_b200 — C-ABI of the B200-native hot path of rakhimovv/lvt ("vidgen").
 *
 * The reference has no FFI layer of its own (it is pure PyTorch; setup.py:38-82,110 only
 * reserves the never-built extension name `vidgen._C`).  Its extension mechanism is the
 * registry + from_config contract of the vidgen/modeling packages; the module classes there call ATen ops.
 * Every entry point below replaces one group of those ATen call sites (cited per function as
 * reference file:line, relative to the reference repo root) and is what a `vidgen._C` binding
 * would bind.  INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C: pointers are DEVICE pointers unless the name ends in _host; sizes are ints /
 *     long long; `stream` is a cudaStream_t passed as void*.
 *   - the caller owns every buffer; no entry point allocates device memory or synchronises
 *     (except where stated), so calls compose with the caller's stream and CUDA-graph capture.
 *   - return value: 0 on success, negative LvtStatus on failure; lvt_last_error() returns a
 *     thread-local, NUL-terminated description of the last failure.
 *   - bf16 tensors are raw uint16 storage (__nv_bfloat16).
 */
#ifndef LVT_B200_H_
#define LVT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVT_B200_ABI_VERSION 1

/* ------------------------------------------------------------------------------------------
 * library / device
 * ---------------------------------------------------------------------------------------- */
int lvt_abi_version(void);
const char* lvt_last_error(void);
/* 0 if the current CUDA device is sm_100 (B200); negative otherwise. No CPU fallback exists. */
int lvt_device_check(void);
/* number of kernels this library has launched since load / last reset (bench `gpu_launches`) */
/* SM budget of the persistent tensor-core kernels (GEMM, attention): n > 0 sizes their grids for n SMs, 0 = all.
   Used while an NCCL gradient all-reduce holds a few SMs next to the backward (the launches captured into a CUDA
   graph keep the budget they were captured with). */
void lvt_set_sm_limit(int n);
long long lvt_launch_count(void);
void lvt_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------
 * VQ codebook  (vidgen/modeling/vq/vq_utils.py:5-24,34-65; vq_embedding.py:23-66,69-99)
 * ---------------------------------------------------------------------------------------- */
/* Nearest-codebook-entry search, DVQEmbedding.forward(mode="") semantics:
 *   z_e      [n, num*D, hw]  fp32, NCHW (hw = H*W flattened) — read in place, no NHWC copy
 *   codebook [num, K, D]     fp32 (ve[i].embedding.weight stacked)
 *   idx_out  [n, num, hw]    int64
 * Distances are evaluated with exactly the reference's fp32 arithmetic
 * (fl(fl(|c|^2+|x|^2) - 2*dot), dot = sequential FMA chain; |.|^2 in ATen's 4x8-lane order,
 * see oracle/vq_oracle.c), first minimum wins => indices are bit-exact.
 * Optional outputs (pass NULL to skip):
 *   zq_out   [n, num*D, hw]  fp32 NCHW: codebook rows gathered at idx (vq_st forward value)
 *   counts   [num, K] fp32 and sums [num, K, D] fp32: per-code histogram / per-code sum of
 *            z_e, ACCUMULATED (+=) into the buffers (EMA statistics, vq_embedding.py:44-55)
 */
int lvt_vq_argmin(const float* z_e, const float* codebook, int64_t* idx_out, float* zq_out,
                  float* counts, float* sums, int n, int num, int K, int D, int hw, void* stream);

/* EMA codebook update, vq_embedding.py:48-59 (after the cross-rank sum of counts/sums):
 *   running_size = decay*running_size + (1-decay)*counts
 *   running_sum  = decay*running_sum  + (1-decay)*sums
 *   n = sum(running_size); size_ = (running_size+eps)/(n+K*eps)*n
 *   codebook = running_sum / size_[:,None]
 * All buffers [num, K(,D)] fp32, updated in place.                                          */
int lvt_vq_ema_update(float* codebook, float* running_size, float* running_sum,
                      const float* counts, const float* sums, int num, int K, int D, double decay,
                      double eps, void* stream);

/* Codebook gradient when MODEL.CODEBOOK.EMA is False: the autograd of index_select under
 * loss = mse(z_q, sg[z_e]) (vq_embedding.py:61-64, meta_arch/vqvae.py:84-85; the codebook argument of vq_st is
 * detached, so vq_utils.py:55-63 contributes nothing there):  grad[k, :] = scale * (counts[k] * codebook[k, :]
 * - sums[k, :]) with scale = 2 / numel(z_e) and counts / sums the per-code statistics lvt_vq_argmin accumulates.
 * rows = num * K; all buffers fp32.                                                          */
int lvt_vq_codebook_grad(const float* counts, const float* sums, const float* codebook, float* grad,
                         float scale, int rows, int D, void* stream);

/* Codebook gather ("emb" mode / z_q_bar): idx [n, num, hw] int64 -> out [n, num*D, hw] fp32
 * NCHW (vq_embedding.py:61-64, 92-97 followed by vqvae.py:104's permute).                   */
int lvt_vq_gather(const int64_t* idx, const float* codebook, float* out, int n, int num, int K,
                  int D, int hw, void* stream);

/* ------------------------------------------------------------------------------------------
 * bf16 tensor-core GEMM (tcgen05.mma + TMEM + TMA), the engine under every Linear / bmm /
 * 1x1x1 Conv3d / masked conv of the DSFVT path
 * (vt_attention.py:63-80,120-128,138; videotransformer.py:52-57,97-99,139-160).
 *
 *   D[z][m, n] = epilogue( alpha * sum_k A[z][m, k] * B[z][n, k] )     m<M, n<N, k<K, z<batch
 *
 * Operand addressing (elements; A shown, B identical with n in place of m):
 *   K-major  (a_mn_major = 0): contiguous coordinate c = k, strided coordinate r = m
 *   MN-major (a_mn_major = 1): contiguous coordinate c = m, strided coordinate r = k
 *   offset = (z / a_zdiv) * a_s_zhi + (z % a_zdiv) * a_s_zlo
 *          + (c / a_cin) * a_s_blk + r * a_ld + (c % a_cin)
 *   (a_cin = block width along the contiguous coordinate; pass the full extent and s_blk = 0
 *    for a plain 2-D matrix.)  The output and all epilogue tensors share one such addressing
 *   (o_*), with c = n and r = m.
 * ---------------------------------------------------------------------------------------- */
enum {
  LVT_EPI_LINEAR = 0,  /* v = alpha*acc [+bias] [+res]; [relu]; [*(aux>0)]                   */
  LVT_EPI_SOFTMAX = 1, /* attention probabilities: N == 256 keys per row (vt_attention.py:63-79)
                          v = alpha*acc + relpos_bias; causal fill -1e4; P = softmax_row(v)  */
  LVT_EPI_DS = 2       /* attention backward: v = aux(P) * (alpha*acc - delta[row])          */
};
enum {
  LVT_GEMM_RELU = 1,       /* out = max(v, 0)                                                */
  LVT_GEMM_MASK = 2,       /* v *= (aux_bf16 > 0)  (ReLU backward)                           */
  LVT_GEMM_ATOMIC = 4,     /* out_f32 += v with red.global.add (split-K / grad accumulate)   */
  LVT_GEMM_CAUSAL = 8,     /* SOFTMAX mode: mask keys j > query i                            */
  LVT_GEMM_AUX_ADD = 16,   /* v += aux_bf16 (bf16 residual, ResBlock skip), before the ReLU  */
  LVT_GEMM_ROWDOT = 32     /* rowdot[...] = sum over each rd_block columns of v * aux_bf16: the softmax-backward
                              row term delta = rowsum(dO * O) (vt_attention.py:75-80) produced by the GEMM
                              that computes dO; bf16 output, batch == 1                          */
};

typedef struct LvtGemm {
  int M, N, K, batch;
  int splits; /* split-K factor (>1 requires LVT_GEMM_ATOMIC and out_bf16 == NULL); < 0: chosen by the library */
  /* A */
  const void* a; int a_mn_major; int a_cin; int a_zdiv;
  long long a_ld, a_s_blk, a_s_zlo, a_s_zhi;
  /* B */
  const void* b; int b_mn_major; int b_cin; int b_zdiv;
  long long b_ld, b_s_blk, b_s_zlo, b_s_zhi;
  /* epilogue */
  int mode; int flags; float alpha;
  float* out_f32;          /* optional */
  void* out_bf16;          /* optional */
  const float* res;        /* optional fp32 residual, output addressing (may alias out_f32) */
  const void* aux_bf16;    /* optional bf16 tensor, output addressing (mask source / P)      */
  const float* bias;       /* optional; bias[(m % bias_mod) * N + n] if bias_mod > 0 else bias[n] */
  int bias_mod;
  int o_cin; int o_zdiv;
  long long o_ld, o_s_blk, o_s_zlo, o_s_zhi;
  /* SOFTMAX / DS modes: per-row vectors indexed [z * M + m] */
  float* lse;              /* SOFTMAX: log-sum-exp out (optional)                            */
  const float* delta;      /* DS: rowsum(dO * O)                                             */
  /* SOFTMAX: relative-position banks of head (z % heads): bank_x[head, 2*bx-1]
     (BlockLocalAttention.get_B, vt_attention.py:169-174); block (bt,bh,bw), bt*bh*bw == 256 */
  const float* bank_t; const float* bank_h; const float* bank_w;
  int bt, bh, bw, heads;
  /* Implicit-GEMM convolution over NHWC bf16 activations (ResEncoder / ResDecoder convs,
     encoder/resencoder.py:46-52, generator/resdecoder.py:48-56 and their autograd): the conv operand
     is an activation tensor [P phases][cv_N][cv_H][cv_W][cv_C] (pixel stride cv_pix_stride elements,
     phase stride cv_s_phase) read through per-tap shifted TMA boxes, out-of-range pixels = 0.
       a_conv: A[m, tap*cv_C + c] = act[ph[tap]][n, h + dh[tap], w + dw[tap], c],  m = (n, h, w)
       b_conv: B[tap*cv_C + c, k] = act[ph[tap]][n, h + dh[tap], w + dw[tap], c],  k = (n, h, w)
               (weight gradient; needs b_mn_major = 1)                                        */
  int a_conv, b_conv;
  int cv_C, cv_W, cv_H, cv_N, cv_P, cv_ntaps;
  long long cv_pix_stride, cv_s_phase;
  signed char cv_dh[16], cv_dw[16], cv_ph[16];
  /* LVT_GEMM_ROWDOT: rowdot[((m / rd_L) * (N / rd_block) + n / rd_block) * rd_L + m % rd_L]
     (= delta[sequence, head, position] for rd_block = da, rd_L = block length)                 */
  float* rowdot; int rd_block, rd_L;
  /* LVT_EPI_SOFTMAX with v != NULL: fused attention forward (vt_attention.py:61-81).  After the softmax the same
     kernel computes O[z] = P[z] @ V[z] with P taken from shared memory: V[z] is [N keys][o2_n] (keys = rows, like
     an MN-major B operand), O goes to o2_bf16 [M, o2_n] (o2_n = da = 128).  out_bf16 (P, needed by the backward)
     becomes optional.
     LVT_EPI_DS with v != NULL: first half of the attention backward in one kernel: dS (out_bf16) as before and
     dQ[z] = alpha * dS[z] @ K[z] with K passed as `v`, dQ as `o2_bf16`; alpha is applied to dQ only.  If bank_t/h/w
     are set (block (1,16,16)) they are the dt/dh/dw_bank GRADIENT buffers [heads, 2*bx-1], accumulated (+=) from dS
     in the epilogue (BlockLocalAttention.get_B, vt_attention.py:169-174).                                     */
  const void* v; int v_cin, v_zdiv; long long v_ld, v_s_zlo, v_s_zhi;
  void* o2_bf16; int o2_n, o2_cin, o2_zdiv; long long o2_ld, o2_s_zlo, o2_s_zhi;
  void* prof;  /* NULL, or 32*8 int64: clock64 timeline of CTA 0 of the fused attention forward (tools/attn_fwd_prof.py) */
} LvtGemm;

int lvt_gemm_bf16(const LvtGemm* g, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused attention backward (autograd of ScaledDotProductAttention.forward, vt_attention.py:61-81, with the
 * relative-position bias of BlockLocalAttention.get_B, vt_attention.py:169-174), one 256-position block per
 * (sequence, head), da = 128.  Neither P nor dS touches HBM: the kernel recomputes P = exp(scale*QK^T + B - lse)
 * from the row log-sum-exp the fused forward saved (LvtGemm.lse) and produces
 *   dV = P^T dO,  dS = P * (dO V^T - delta),  dK = scale * dS^T Q,  dQ = scale * dS K,  dbank_x += sums of dS.
 *   qkv / dqkv  bf16 [nseq*256, ld]: q | k | v at columns 0, heads*128, 2*heads*128, head h at + h*128
 *   dO          bf16 [nseq*256, do_ld], head h at column h*128
 *   lse, delta  fp32 [nseq*heads, 256]  (natural-log lse; delta = rowsum(dO * O), LVT_GEMM_ROWDOT)
 *   bank_x      fp32 [heads, 2*bx-1] values;  dbank_x: their gradients, ACCUMULATED (+=)
 *   scratch     lvt_attn_bwd_scratch_bytes() bytes of device memory (dQ partial sums; contents irrelevant)
 * ---------------------------------------------------------------------------------------- */
typedef struct LvtAttnBwd {
  int nseq, heads;
  int bt, bh, bw;          /* attention block: (1,16,16) or (4,8,8) */
  int causal;              /* keys after the query were filled with -1e4 in the forward */
  float scale;             /* 1/sqrt(da) */
  const void* qkv; long long qkv_ld;
  const void* dO; long long do_ld;
  void* dqkv; long long dqkv_ld;
  const float* lse; const float* delta;
  const float* bank_t; const float* bank_h; const float* bank_w;
  float* dbank_t; float* dbank_h; float* dbank_w;
  float* scratch; long long scratch_bytes;
  void* prof;              /* NULL, or 64*16 int64: clock64 timeline of CTA 0 (tools/attn_bwd_prof.py) */
} LvtAttnBwd;
long long lvt_attn_bwd_scratch_bytes(void);
int lvt_attn_bwd(const LvtAttnBwd* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * DSFVT bandwidth-bound operators
 * ---------------------------------------------------------------------------------------- */
/* nn.LayerNorm(d) forward (vt_attention.py:121,138; videotransformer.py:143): x fp32 [M,d] ->
 * y bf16 [M,d]; mean/rstd [M] are saved for the backward.  d in {128,256,512,1024}.          */
int lvt_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16,
                      float* mean, float* rstd, int M, int d, float eps, void* stream);
/* LayerNorm backward: dx = [dres +] dLN(dy); written as fp32 (dx_f32) and/or bf16 (dx_bf16);
 * dgamma/dbeta [d] are ACCUMULATED (+=).  d in {128,256,512}.                                */
int lvt_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                      const float* gamma, const float* dres, float* dx_f32, void* dx_bf16,
                      float* dgamma, float* dbeta, int M, int d, void* stream);
/* same with dy in bf16 (as written by the GEMM that produced it: halves that round trip).     */
int lvt_layernorm_bwd_bf16dy(const void* dy_bf16, const float* x, const float* mean, const float* rstd,
                             const float* gamma, const float* dres, float* dx_f32, void* dx_bf16,
                             float* dgamma, float* dbeta, int M, int d, void* stream);
/* general form: dy fp32 or bf16; dx_colsum [d] (optional) is ACCUMULATED with the column sums of dx -- the bias
 * gradient of the nn.Linear whose output x is (ffn.3 of the previous BlockLocalAttention layer,
 * vt_attention.py:138), which saves a separate pass over dx.                                 */
int lvt_layernorm_bwd_ex(const void* dy, int dy_is_bf16, const float* x, const float* mean, const float* rstd,
                         const float* gamma, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma,
                         float* dbeta, float* dx_colsum, int M, int d, void* stream);
/* out[n] += sum_m x[m*ld + n], x bf16 (bias gradients of nn.Linear / Conv3d).                */
int lvt_colsum_bf16(const void* x, float* out, int M, int N, long long ld, void* stream);
/* delta[b,h,i] = sum_d dO[b*L+i, h*da+d] * O[b*L+i, h*da+d] (softmax backward row term of
 * ScaledDotProductAttention, vt_attention.py:75-80); dO, O bf16 [nb*L, H*da].                */
int lvt_attn_delta(const void* dO, const void* O, float* delta, int nb, int H, int L, int da,
                   void* stream);
/* Gradient of dt_bank/dh_bank/dw_bank (BlockLocalAttention.get_B, vt_attention.py:169-174):
 * dS bf16 [nb, H, 256, 256]; dbank_x [H, 2*bx-1] ACCUMULATED.                                */
int lvt_relpos_bank_grad(const void* dS, float* dbank_t, float* dbank_h, float* dbank_w, int nb,
                         int H, int bt, int bh, int bw, void* stream);
/* VTEncoder front end (videotransformer.py:41-53): pad-aware one-hot -> Conv3d(nc*nv -> de,
 * kernel, stride) -> + slice_embedding[slice_idx], evaluated as a gather-sum.
 *   context [B, nc, Tc, Hc, Wc] int64 (pad_value entries contribute nothing)
 *   wt      [nc, kt, kh, kw, nv, de] fp32 = encoder.conv.weight permuted (lvt_permute4)
 *   out     bf16 [B*to*ho*wo, de]   (to = (Tc-kt)/st+1, ...);  ctx_shape/kernel/stride: int[3]
 * backward: dout fp32 [rows, de] is scattered (+=) into dwt and dslice_emb.                  */
int lvt_vt_enc_front_fwd(const int64_t* context, const int64_t* slice_idx, const float* wt,
                         const float* bias, const float* slice_emb, void* out_bf16, int B, int nc,
                         int nv, int de, const int* ctx_shape, const int* kernel,
                         const int* stride, int pad_value, void* stream);
int lvt_vt_enc_front_bwd(const int64_t* context, const int64_t* slice_idx, const float* dout,
                         float* dwt, float* dslice_emb, int B, int nc, int nv, int de,
                         const int* ctx_shape, const int* kernel, const int* stride,
                         int pad_value, void* stream);
/* VTDecoder front end (videotransformer.py:80-89 embed_sum + the im2col of MaskedConv3d,
 * vt_utils.py:183-200): out bf16 [B*t*h*w, ntaps*de]; row m, tap q = sum_k emb[k, slice[b,k,
 * pos(m)+taps[q]]] or 0 outside the slice.  emb [nc, nv, de] fp32; taps int[ntaps*3] (device).
 * backward: dA fp32 [rows, ntaps*de] -> demb [nc, nv, de] ACCUMULATED.                       */
int lvt_vt_dec_front_fwd(const int64_t* slice, const float* emb, const int* taps, void* out_bf16,
                         int B, int nc, int nv, int de, int t, int h, int w, int ntaps,
                         void* stream);
int lvt_vt_dec_front_bwd(const int64_t* slice, const float* dA, const int* taps, float* demb, int B,
                         int nc, int nv, int de, int t, int h, int w, int ntaps, void* stream);
/* ChannelPredictor U[k] one-hot half + ReLU (videotransformer.py:148-150):
 * a[m,:] = relu(u[m,:] + sum_{j<k} ut[j*nv + slice[b,j,pos], :]); u fp32 [M,d] (dense half incl.
 * bias), ut fp32 [k*nv, d] = U[k].weight[:, d:]^T, a bf16 [M,d].  backward scatters du (bf16)
 * into dut (+=).                                                                            */
int lvt_chpred_combine_fwd(const float* u, const float* ut, const int64_t* slice, void* a_bf16,
                           int M, int nc, int nv, int d, int thw, int k, void* stream);
int lvt_chpred_combine_bwd(const void* du_bf16, const int64_t* slice, float* dut, int M, int nc,
                           int nv, int d, int thw, int k, void* stream);
/* loss = 1/nc * sum_k mean_valid CE(logits[k], slice[:,k]) with the shared ignore mask
 * (meta_arch/vt.py:305-312).  logits fp32 [nc, B*thw, nv]; slice int64 [B, nc, thw]; ignore
 * uint8 [B, thw]; dlogits bf16 [nc, B*thw, nv] (optional) receives dloss/dlogits; loss: 1 float
 * (overwritten); count_scratch: 1 int.                                                      */
int lvt_cross_entropy(const float* logits, const int64_t* slice, const uint8_t* ignore,
                      void* dlogits_bf16, float* loss, int* count_scratch, int B, int nc, int nv,
                      int thw, void* stream);
/* One channel of ChannelPredictor.sample (videotransformer.py:161-185) at position *pos of every sequence:
 * slice[b, k, *pos] = argmax_i softmax(logits[b*thw + *pos, :] / temp)_i / q_exp[b, i]  with q_exp ~ Exp(1) drawn by the
 * caller — the arithmetic of torch.multinomial(probs, 1), so both consume the same random stream.  logits fp32
 * [B*logit_rows, nv] with logit_rows == thw (full pass) or 1 (one row per sequence, incremental pass); slice int64
 * [B, nc, thw]; pos: ONE int64 in device memory (graph replay with a moving position).                          */
int lvt_vt_sample_pixel(const float* logits, const float* q_exp, const int64_t* pos, int64_t* slice, int B,
                        int thw, int nv, int nc, int k, float temp, int logit_rows, void* stream);

/* ------------------------------------------------------------------------------------------
 * Incremental (K/V-cached) decoding step of the sampler (meta_arch/vt.py:107-134 recomputes the whole decoder
 * pass per position): skinny B-row products against bf16 weights, B <= 16.  *pos = current position.
 * ---------------------------------------------------------------------------------------- */
typedef struct LvtRowsLinear {
  int B, N, K;
  const void* x; long long x_ldb, x_pos_mul; int x_bf16;   /* row b at x + b*x_ldb + (*pos)*x_pos_mul (elements) */
  const float* ln_gamma; const float* ln_beta; float ln_eps; /* optional LayerNorm of the row first              */
  int round_in;                                             /* round the input row to bf16 (as the full pass)  */
  const void* w_bf16; long long w_ld;                       /* W [N, K] rows (nn.Linear layout), bf16          */
  const float* bias;                                        /* optional [N]                                    */
  const float* res; long long res_ldb, res_pos_mul;         /* optional residual row                           */
  const float* gtab; const int64_t* slice; int g_count, nv, nc, thw;
                       /* out[b,n] += sum_{j<g_count} gtab[(j*nv + slice[b,j,*pos])*N + n] (one-hot half of U[k]) */
  int relu, round_out;
  float* out; long long out_ldb;
  const int64_t* pos;
} LvtRowsLinear;
/* out[b, :] = epi([LN](x[b, :]) @ W^T) */
int lvt_rows_linear(const LvtRowsLinear* a, void* stream);
/* q | k | v of row *pos from w_q|w_k|w_v [3H, d, da] (vt_attention.py:98-104,120-124) with the pre-LayerNorm fused:
 * q -> q_out fp32 [B, H*da]; k, v -> caches bf16 [B, H, L, da] at row *pos.                                   */
int lvt_rows_qkv(const float* x, const float* ln_gamma, const float* ln_beta, float eps, const void* w_bf16,
                 float* q_out, void* k_cache_bf16, void* v_cache_bf16, const int64_t* pos, int B, int H, int d,
                 int da, int L, void* stream);
/* o[b, h, :] = softmax(q k^T * scale + B[h, *pos, :] [keys after *pos: -1e4]) v over the cached rows
 * (vt_attention.py:61-81,169-174).                                                                           */
int lvt_attn_row(const float* q, const void* k_cache_bf16, const void* v_cache_bf16, const float* bank_t,
                 const float* bank_h, const float* bank_w, int bt, int bh, int bw, const int64_t* pos, float scale,
                 float* o, int B, int H, int L, int da, void* stream);
/* The whole per-position step of the incremental sampler as ONE persistent kernel (32 CTAs, grid barriers between
 * dependent stages): row *pos through the masked decoder against the K/V caches (the stages of lvt_vt_dec_front_fwd,
 * lvt_rows_linear, lvt_rows_qkv, lvt_attn_row above, same arithmetic), then -- if do_sample -- per channel the
 * predictor (U[k] with the one-hot row gather, P[k]) and the categorical draw of lvt_vt_sample_pixel into
 * slice[b, k, *pos].  q_exp [nc, B, nv]: Exp(1) noise drawn by the caller, channel by channel, BEFORE the call (torch's
 * generator: the random stream of torch.multinomial in meta_arch/vt.py:107-134, videotransformer.py:161-185).
 * Activation scratch xa, xb, hbuf, a1, abuf [B, d], q, o [B, H*da], logits [B, nv] fp32; barrier: 2 zero-initialised
 * unsigned ints owned by this entry point.  B <= 16, da == 128, L = t*h*w = bt*bh*bw <= 256, <= 8 layers, <= 4 channels. */
typedef struct LvtDecodeLayer {
  const float* ln1_g; const float* ln1_b;        /* mha.layer_norm                                   */
  const void* w_qkv;                              /* w_q | w_k | w_v [3H, d, da] bf16                 */
  void* k_cache; void* v_cache;                   /* bf16 [B, H, L, da]                               */
  const float* bank_t; const float* bank_h; const float* bank_w;
  const void* w_proj;                             /* mha.proj.weight [d, H*da] bf16                   */
  const float* ln2_g; const float* ln2_b;        /* ffn.0                                            */
  const void* w_ffn1; const float* b_ffn1;       /* ffn.1 [d, d] bf16, bias fp32                     */
  const void* w_ffn3; const float* b_ffn3;       /* ffn.3                                            */
} LvtDecodeLayer;
typedef struct LvtDecodeStep {
  int B, d, H, da, L, nc, nv, de, ntaps, n_layers;
  int bt, bh, bw, t, h, w;
  float scale, ln_eps, temp;
  int do_sample;                                  /* 0: decoder row only (primed position: fills the K/V caches) */
  const int64_t* pos;                             /* ONE int64 in device memory                       */
  int64_t* slice;                                 /* [B, nc, L]                                       */
  const float* emb;                               /* decoder.ch_embedder [nc, nv, de] fp32            */
  const int* taps;                                /* live taps of the masked conv, int[ntaps*3]       */
  const void* conv_w;                             /* packed live-tap conv weight [d, ntaps*de] bf16   */
  const float* y0s;                               /* [B*L, d]: zl Wlp^T + positional encoding + conv bias */
  LvtDecodeLayer layer[8];
  const float* lnp_g; const float* lnp_b;        /* ch_predictor.layer_norm                          */
  const void* U[4]; long long U_ld[4]; const float* U_bias[4]; const float* gtab[4];
  const void* P[4]; const float* P_bias[4];
  const float* q_exp;
  float* xa; float* xb; float* hbuf; float* a1; float* q; float* o; float* abuf; float* logits;
  unsigned* barrier;
  long long* prof;                                /* NULL, or 128 int64: %globaltimer stamps of CTA 0 per stage (tools/) */
} LvtDecodeStep;
int lvt_vt_decode_step(const LvtDecodeStep* p, void* stream);
/* torch.optim.RMSprop / Adam steps (solver/build.py:62-72) over flat fp32 buffers of n elements
 * (n % 4 == 0), gradients pre-multiplied by grad_scale; p_bf16 (optional) receives the bf16
 * shadow copy the GEMMs read.                                                               */
int lvt_rmsprop_step(float* p, const float* g, float* sq, float* buf, void* p_bf16, long long n,
                     float lr, float alpha, float momentum, float eps, float grad_scale,
                     void* stream);
int lvt_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr,
                  float beta1, float beta2, float eps, int step, float grad_scale, void* stream);
int lvt_cast_bf16(const float* in, void* out_bf16, long long n, void* stream);
/* out[sum i_k*out_strides[k]] (=|+=) in[sum i_k*in_strides[k]] over dims[4] (layout packing of
 * the small weights: one-hot conv, masked conv taps, U[k] one-hot halves).                  */
int lvt_permute4(const float* in, void* out, int out_is_bf16, int accumulate, const int* dims,
                 const long long* in_strides, const long long* out_strides, void* stream);

/* The same for a table of jobs in ONE launch (per-step weight packs / gradient folds).  jobs: DEVICE array; job j covers
 * blocks [first_block, first_block + ceil(numel / 1024)), ascending; total_blocks = their sum.  The jobs of one call must
 * not depend on each other.                                                                  */
typedef struct LvtPermuteJob {
  const float* in; void* out;
  int out_is_bf16, accumulate;
  int dims[4];
  long long in_strides[4], out_strides[4];
  long long first_block;
} LvtPermuteJob;
int lvt_permute4_batch(const LvtPermuteJob* jobs, int n_jobs, int total_blocks, void* stream);

/* Class conditioning of VTEncoder (CLASS_NUM > 0; videotransformer.py:29-33,54-57): the class embedding is
 * concatenated to the de channels of every position before the 1x1x1 projector (weight (d, 2de)), which equals a
 * per-sample bias.  w2 = &W[0][de] with row stride ldw = 2de, emb [class_num, de], cls int64 [B], all fp32.
 *   lvt_vt_class_bias       cb[b, :] = W2 emb[cls[b]]                                  (forward)
 *   lvt_rows_add_group_bias x[r, :] += cb[r / rows_per_group, :]                       (x fp32 [M, d], in place)
 *   lvt_colsum_groups_bf16  out[g, :] = sum of the rows of group g of x (bf16 [groups*rows, N]) (backward: S)
 *   lvt_vt_class_grad       dW2 += S^T emb[cls];  demb[cls[b]] += S[b] W2              (backward)            */
int lvt_vt_class_bias(const float* w2, long long ldw, const float* emb, const int64_t* cls, float* cb, int B, int d,
                      int de, void* stream);
int lvt_rows_add_group_bias(float* x, const float* cb, long long M, int d, int rows_per_group, void* stream);
int lvt_colsum_groups_bf16(const void* x_bf16, float* out, int groups, int rows, int N, void* stream);
int lvt_vt_class_grad(const float* S, const float* emb, const int64_t* cls, const float* w2, float* dw2, long long ldw,
                      float* demb, int B, int d, int de, void* stream);

/* Row gather dst[r, :] = src[idx[r], :] (rows of row_bytes bytes, a multiple of 16; idx int32 [M]; src != dst): the
 * token re-ordering of the general tiled BlockLocalAttention.forward (vt_attention.py:189-200: split the slice grid
 * into blocks, attend inside each block, put the tokens back), applied once around a whole stack of layers.        */
int lvt_rows_gather(const void* src, void* dst, const int* idx, int M, int row_bytes, void* stream);

/* Channels-last variants used inside the VQ-VAE engine (z_e [n*hw, num*D] fp32 as written by the
 * last encoder GEMM): same arithmetic and index layout ([n, num, hw] int64) as lvt_vq_argmin;
 * zq may additionally be produced as bf16 (decoder GEMM operand).                             */
int lvt_vq_argmin_nhwc(const float* z_e, const float* codebook, int64_t* idx_out, float* zq_out,
                       void* zq_bf16, float* counts, float* sums, int n, int num, int K, int D, int hw,
                       void* stream);
int lvt_vq_gather_nhwc(const int64_t* idx, const float* codebook, float* out, void* out_bf16, int n,
                       int num, int K, int D, int hw, void* stream);

/* ------------------------------------------------------------------------------------------
 * VQ-VAE edge operators (3-channel ends of ResEncoder / ResDecoder, losses)
 * "phase-major": a 32x32 feature map stored as [hp][wp][n][16][16][C], pixel (2*h2+hp, 2*w2+wp).
 * ---------------------------------------------------------------------------------------- */
/* normalise (ae.py:34-36) + im2col of Conv2d(3->NF/2, k4, s2, p1) (resencoder.py:47):
 * x fp32 NCHW [n,3,64,64] -> A bf16 [4*n*256, 64] (48 columns k=(kh*4+kw)*3+c, 16 zeros), rows in
 * phase-major order of the 32x32 output.                                                      */
int lvt_vqvae_in_im2col(const float* x, void* a_bf16, int n, float mean, float std, void* stream);
/* High-precision encoder (inference / CodesExtractor: the latents are the wire format between the two models, and a
 * bf16 encoder flips ~1-2 % of the codes of near-tied positions).  Every fp32 value v is carried as the bf16 pair
 * hi = bf16(v), lo = bf16(v - hi); activations are stored [hi | lo | hi], weights [hi | hi | lo] (three C-wide segments),
 * so ONE lvt_gemm_bf16 over K = 3C contracts a_hi w_hi + a_lo w_hi + a_hi w_lo with fp32 accumulation: a w up to
 * ~2^-17 relative.  lvt_split3_bf16: v = [relu](in [+ add]) (fp32 [rows, C]) -> out_bf16 [rows, 3C] (+ out_f32 = v);
 * lvt_vqvae_in_im2col_split: as lvt_vqvae_in_im2col with A [4*n*256, 192] in the [hi | lo | hi] form.            */
int lvt_split3_bf16(const float* in, const float* add, void* out_bf16, float* out_f32, long long rows, int C, int relu,
                    int weight_pattern, void* stream);
int lvt_vqvae_in_im2col_split(const float* x, void* a_bf16, int n, float mean, float std, void* stream);
/* ConvTranspose2d(C->3, k4, s2, p1) + tanh (resdecoder.py:56,68-69), second half: the contraction over the C
 * input channels is one lvt_gemm_bf16 call Y[row, (kh*4+kw)*3+co] = sum_c act[row,c] * w[c][co][kh][kw] over the
 * phase-major 32x32 input pixels; this entry gathers the 2x2 (input pixel, tap) pairs of every output pixel, adds
 * the bias and applies tanh.  y fp32 [4*n*256, 64], out fp32 NCHW [n,3,64,64].                  */
int lvt_vqvae_out_col2im_tanh(const float* y, const float* bias, float* out, int n, void* stream);
/* reconstruction loss lambda*MSE(x_tilde, (x-mean)/std) (loss.py:20, vqvae.py:79): loss += value,
 * dpre = dL/d(pre-tanh) [n,3,64,64] (optional), dbias[3] += column sums (optional).           */
int lvt_vqvae_recon_loss(const float* x_tilde, const float* x, float* dpre, float* loss, float* dbias,
                         int n, float mean, float std, float lambda, void* stream);
/* backward of the output ConvTranspose2d, first half: G bf16 [n*1024, 64] with
 * G[row, co*16+kh*4+kw] = dpre at the output pixel the tap reaches (0 outside), so that both
 * dW[c][co][kh][kw] = sum_rows act[row,c] * G[row, .] and dact = relu'(act) * (G @ w^T) are lvt_gemm_bf16 calls. */
int lvt_vqvae_out_convt_g(const float* dpre, void* g_bf16, int n, void* stream);
/* commitment loss beta*MSE(z_e, zq_bar) (vqvae.py:86) + merged gradient wrt z_e:
 * dz (bf16) = dz_st + 2*beta/numel*(z_e - zq_bar);  loss += value.                            */
int lvt_vqvae_commit_loss(const float* z_e, const float* zq_bar, const float* dz_st, void* dz_bf16,
                          float* loss, long long numel, float beta, void* stream);
/* out = (a [+ b]) * (mask_src > 0), all bf16 (ReLU backward merged with a skip gradient).     */
/* x (fp32, n elements, n % 4 == 0) += a (bf16): the skip connection added to the fp32 output z_e of the last encoder
 * ResBlock (resencoder.py:19-21) after its 1x1 convolution has gone through the TMA-store GEMM epilogue. */
int lvt_add_bf16_to_f32(float* x, const void* a_bf16, long long n, void* stream);
int lvt_relu_bwd_add(const void* a_bf16, const void* b_bf16, const void* mask_src_bf16, void* out_bf16,
                     long long n, void* stream);
int lvt_cast_relu_bf16(const float* in, void* out_bf16, long long n, int relu, void* stream);
/* y*std + mean clamped to [lo, hi] (ae.py:130-139).                                           */
int lvt_denorm_clamp(const float* in, float* out, long long n, float mean, float std, float lo, float hi,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LVT_B200_H_ */
