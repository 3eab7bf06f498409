"""Thin Python wrappers over the C-ABI (include/lvt_b200.h): raw device pointers of torch
tensors + the current CUDA stream go straight into liblvt_b200.so.  torch is only the owner of
device memory and streams here; no torch op sits on these paths.
"""
import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import LvtAttnBwd, LvtGemm, check, ptr, stream_ptr

EPI_LINEAR, EPI_SOFTMAX, EPI_DS = 0, 1, 2
GEMM_RELU, GEMM_MASK, GEMM_ATOMIC, GEMM_CAUSAL, GEMM_AUX_ADD, GEMM_ROWDOT = 1, 2, 4, 8, 16, 32


def _cuda_contig(t, dtype, name):
    if not t.is_cuda:
        raise _lib.LvtError(f"{name} must be a CUDA tensor (lvt_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.LvtError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.LvtError(f"{name} must be contiguous")
    return t


# --------------------------------------------------------------------------------------------
# VQ codebook
# --------------------------------------------------------------------------------------------
def vq_argmin(z_e, codebook, want_zq=False, counts=None, sums=None):
    """z_e [n, num*D, h, w] fp32 NCHW; codebook [num, K, D] fp32 -> idx [n, num, h, w] int64
    (+ z_q [n, num*D, h, w] gathered from `codebook` when want_zq).  Accumulates EMA statistics
    into counts [num,K] / sums [num,K,D] when given.  (vq_utils.py:7-24, vq_embedding.py:36-55)"""
    lib = _lib.require_device()
    _cuda_contig(z_e, torch.float32, "z_e")
    _cuda_contig(codebook, torch.float32, "codebook")
    n, c, h, w = z_e.shape
    num, K, D = codebook.shape
    if c != num * D:
        raise _lib.LvtError(f"z_e has {c} channels, codebook expects {num}*{D}")
    idx = torch.empty((n, num, h, w), dtype=torch.int64, device=z_e.device)
    zq = torch.empty_like(z_e) if want_zq else None
    check(lib.lvt_vq_argmin(ptr(z_e), ptr(codebook), ptr(idx), ptr(zq), ptr(counts), ptr(sums),
                            n, num, K, D, h * w, stream_ptr()), "lvt_vq_argmin")
    return (idx, zq) if want_zq else idx


def vq_argmin_nhwc(z_e, codebook, hw, want_zq=False, want_zq_bf16=False, counts=None, sums=None):
    """Channels-last variant: z_e [n*hw, num*D] fp32 -> idx [n, num, hw] int64 (+ z_q fp32 / bf16 [n*hw, num*D])."""
    lib = _lib.require_device()
    _cuda_contig(z_e, torch.float32, "z_e")
    _cuda_contig(codebook, torch.float32, "codebook")
    M, c = z_e.shape
    num, K, D = codebook.shape
    assert c == num * D and M % hw == 0
    n = M // hw
    idx = torch.empty((n, num, hw), dtype=torch.int64, device=z_e.device)
    zq = torch.empty_like(z_e) if want_zq else None
    zqb = torch.empty_like(z_e, dtype=torch.bfloat16) if want_zq_bf16 else None
    check(lib.lvt_vq_argmin_nhwc(ptr(z_e), ptr(codebook), ptr(idx), ptr(zq), ptr(zqb), ptr(counts), ptr(sums),
                                 n, num, K, D, hw, stream_ptr()), "lvt_vq_argmin_nhwc")
    return idx, zq, zqb


def vq_gather(idx, codebook):
    """idx [n, num, h, w] int64 -> [n, num*D, h, w] fp32 NCHW (vq_embedding.py:92-97 + permute)."""
    lib = _lib.require_device()
    _cuda_contig(idx, torch.int64, "idx")
    _cuda_contig(codebook, torch.float32, "codebook")
    n, num, h, w = idx.shape
    num2, K, D = codebook.shape
    assert num == num2
    out = torch.empty((n, num * D, h, w), dtype=torch.float32, device=idx.device)
    check(lib.lvt_vq_gather(ptr(idx), ptr(codebook), ptr(out), n, num, K, D, h * w, stream_ptr()),
          "lvt_vq_gather")
    return out


def vq_ema_update(codebook, running_size, running_sum, counts, sums, decay=0.99, eps=1e-5):
    lib = _lib.require_device()
    num, K, D = codebook.shape
    check(lib.lvt_vq_ema_update(ptr(codebook), ptr(running_size), ptr(running_sum), ptr(counts),
                                ptr(sums), num, K, D, float(decay), float(eps), stream_ptr()),
          "lvt_vq_ema_update")


# --------------------------------------------------------------------------------------------
# GEMM
# --------------------------------------------------------------------------------------------
@dataclass
class Operand:
    """Addressing of one GEMM operand / of the output (see `struct LvtGemm`)."""
    data: int                 # device pointer
    ld: int                   # stride of the strided coordinate (elements)
    mn_major: bool = False
    cin: int = 0              # 0 -> full extent (plain 2-D)
    s_blk: int = 0
    zdiv: int = 1
    s_zlo: int = 0
    s_zhi: int = 0


def op_kmajor(t, rows=None):
    """2-D tensor [rows, K] (row stride arbitrary, last dim contiguous) as a K-major operand."""
    assert t.stride(-1) == 1
    return Operand(t.data_ptr(), t.stride(0))


def op_mnmajor(t):
    """2-D tensor [K, rows] (last dim contiguous) as an MN-major operand (i.e. used transposed)."""
    assert t.stride(-1) == 1
    return Operand(t.data_ptr(), t.stride(0), mn_major=True)


@dataclass
class ConvSpec:
    """Implicit-GEMM convolution operand: NHWC bf16 activations [P][n][h][w][C] read through
    per-tap shifted boxes (see `struct LvtGemm`); taps = [(dh, dw, phase), ...]."""
    side: str                 # "a" (forward / data gradient) or "b" (weight gradient)
    C: int
    H: int
    W: int
    n: int
    taps: list
    P: int = 1
    pix_stride: int = 0
    s_phase: int = 0


def gemm(M, N, K, a: Operand, b: Operand, out: Operand, out_f32=None, out_bf16=None, batch=1,
         splits=1, alpha=1.0, mode=EPI_LINEAR, flags=0, bias=None, bias_mod=0, res=None, aux=None,
         lse=None, delta=None, banks=None, block=None, heads=1, conv: Optional[ConvSpec] = None,
         rowdot=None, rd_block=0, rd_L=0, v: Optional[Operand] = None, o2: Optional[Operand] = None, o2_n=0, prof=None):
    """D[z] = epilogue(alpha * A[z] @ B[z]^T); pointers may be torch tensors or ints."""
    lib = _lib.require_device()

    def p(x):
        if x is None:
            return None
        if isinstance(x, int):
            return ctypes.c_void_p(x)
        return ctypes.c_void_p(x.data_ptr())

    g = LvtGemm()
    g.M, g.N, g.K, g.batch, g.splits = M, N, K, batch, splits
    g.a = a.data
    g.a_mn_major = int(a.mn_major)
    g.a_cin = a.cin or (M if a.mn_major else K)
    g.a_zdiv, g.a_ld, g.a_s_blk, g.a_s_zlo, g.a_s_zhi = a.zdiv, a.ld, a.s_blk, a.s_zlo, a.s_zhi
    g.b = b.data
    g.b_mn_major = int(b.mn_major)
    g.b_cin = b.cin or (N if b.mn_major else K)
    g.b_zdiv, g.b_ld, g.b_s_blk, g.b_s_zlo, g.b_s_zhi = b.zdiv, b.ld, b.s_blk, b.s_zlo, b.s_zhi
    g.mode, g.flags, g.alpha = mode, flags, alpha
    g.out_f32, g.out_bf16 = p(out_f32), p(out_bf16)
    g.res, g.aux_bf16, g.bias, g.bias_mod = p(res), p(aux), p(bias), bias_mod
    g.o_cin = out.cin or N
    g.o_zdiv, g.o_ld, g.o_s_blk, g.o_s_zlo, g.o_s_zhi = out.zdiv, out.ld, out.s_blk, out.s_zlo, out.s_zhi
    g.lse, g.delta = p(lse), p(delta)
    if banks is not None:
        g.bank_t, g.bank_h, g.bank_w = p(banks[0]), p(banks[1]), p(banks[2])
        g.bt, g.bh, g.bw = block
    g.heads = heads
    if conv is not None:
        g.a_conv, g.b_conv = int(conv.side == "a"), int(conv.side == "b")
        g.cv_C, g.cv_W, g.cv_H, g.cv_N, g.cv_P, g.cv_ntaps = conv.C, conv.W, conv.H, conv.n, conv.P, len(conv.taps)
        g.cv_pix_stride, g.cv_s_phase = conv.pix_stride or conv.C, conv.s_phase
        for i, (dh, dw, ph) in enumerate(conv.taps):
            g.cv_dh[i], g.cv_dw[i], g.cv_ph[i] = dh, dw, ph
    if v is not None:  # fused attention forward: O = softmax(...) @ V in the same kernel
        g.v, g.v_cin, g.v_zdiv, g.v_ld, g.v_s_zlo, g.v_s_zhi = v.data, v.cin or o2_n, v.zdiv, v.ld, v.s_zlo, v.s_zhi
        g.o2_bf16, g.o2_n, g.o2_cin, g.o2_zdiv = o2.data, o2_n, o2.cin or o2_n, o2.zdiv
        g.o2_ld, g.o2_s_zlo, g.o2_s_zhi = o2.ld, o2.s_zlo, o2.s_zhi
    if rowdot is not None:
        g.rowdot, g.rd_block, g.rd_L = p(rowdot), rd_block, rd_L
        g.flags |= GEMM_ROWDOT
    if prof is not None:
        g.prof = p(prof)
    check(lib.lvt_gemm_bf16(ctypes.byref(g), stream_ptr()), "lvt_gemm_bf16")


_attn_scratch = {}


def attn_bwd_scratch(device=None):
    """Per-device scratch of the fused attention backward (dQ partial sums; lvt_attn_bwd_scratch_bytes())."""
    dev = torch.device(device if device is not None else torch.cuda.current_device())
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _attn_scratch:
        n = int(_lib.require_device().lvt_attn_bwd_scratch_bytes())
        _attn_scratch[key] = torch.empty(n // 4, dtype=torch.float32, device=f"cuda:{key}")
    return _attn_scratch[key]


def attn_bwd(qkv, dO, dqkv, lse, delta, banks, dbanks, nseq, heads, block, causal, scale, qkv_ld=None, do_ld=None,
             scratch=None, prof=None):
    """Fused attention backward (lvt_attn_bwd): dqkv <- (dQ | dK | dV), dbanks += bank gradients; P and dS never
    reach HBM.  Tensors or raw device pointers (ints)."""
    lib = _lib.require_device()

    def p(x):
        return ctypes.c_void_p(x) if isinstance(x, int) else ctypes.c_void_p(x.data_ptr())

    a = LvtAttnBwd()
    a.nseq, a.heads = nseq, heads
    a.bt, a.bh, a.bw = block
    a.causal, a.scale = int(bool(causal)), scale
    a.qkv, a.qkv_ld = p(qkv), qkv_ld or 3 * heads * 128
    a.dO, a.do_ld = p(dO), do_ld or heads * 128
    a.dqkv, a.dqkv_ld = p(dqkv), qkv_ld or 3 * heads * 128
    a.lse, a.delta = p(lse), p(delta)
    a.bank_t, a.bank_h, a.bank_w = (p(b) for b in banks)
    a.dbank_t, a.dbank_h, a.dbank_w = (p(b) for b in dbanks)
    scratch = scratch if scratch is not None else attn_bwd_scratch()
    a.scratch, a.scratch_bytes = p(scratch), scratch.numel() * 4
    a.prof = p(prof) if prof is not None else None
    check(lib.lvt_attn_bwd(ctypes.byref(a), stream_ptr()), "lvt_attn_bwd")


def linear_bf16(x, w, bias=None, relu=False, out_dtype=torch.bfloat16, res=None):
    """y = x @ w^T (+bias) (+res) [relu]; x [M,K] bf16, w [N,K] bf16 (nn.Linear layout)."""
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=out_dtype, device=x.device)
    gemm(M, N, K, op_kmajor(x), op_kmajor(w), Operand(y.data_ptr(), N),
         out_f32=y if out_dtype == torch.float32 else None,
         out_bf16=y if out_dtype == torch.bfloat16 else None,
         bias=bias, res=res, flags=GEMM_RELU if relu else 0)
    return y
