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
  LVT_GEMM_CAUSAL = 8      /* SOFTMAX mode: mask keys j > query i                            */
};

typedef struct LvtGemm {
  int M, N, K, batch;
  int splits; /* split-K factor (>1 requires LVT_GEMM_ATOMIC and out_bf16 == NULL) */
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
} LvtGemm;

int lvt_gemm_bf16(const LvtGemm* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LVT_B200_H_ */
