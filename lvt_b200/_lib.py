"""ctypes binding of liblvt_b200.so — the C-ABI declared in include/lvt_b200.h.

There is deliberately no fallback: if the shared library is missing or the device is not a
B200 the product path raises.  (The CPU oracle under oracle/ is test infrastructure only.)
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblvt_b200.so")

_lib = None


class LvtError(RuntimeError):
    pass


class LvtGemm(ctypes.Structure):
    """Mirror of `struct LvtGemm` (include/lvt_b200.h)."""
    _fields_ = [
        ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int), ("batch", ctypes.c_int),
        ("splits", ctypes.c_int),
        ("a", ctypes.c_void_p), ("a_mn_major", ctypes.c_int), ("a_cin", ctypes.c_int),
        ("a_zdiv", ctypes.c_int),
        ("a_ld", ctypes.c_longlong), ("a_s_blk", ctypes.c_longlong), ("a_s_zlo", ctypes.c_longlong),
        ("a_s_zhi", ctypes.c_longlong),
        ("b", ctypes.c_void_p), ("b_mn_major", ctypes.c_int), ("b_cin", ctypes.c_int),
        ("b_zdiv", ctypes.c_int),
        ("b_ld", ctypes.c_longlong), ("b_s_blk", ctypes.c_longlong), ("b_s_zlo", ctypes.c_longlong),
        ("b_s_zhi", ctypes.c_longlong),
        ("mode", ctypes.c_int), ("flags", ctypes.c_int), ("alpha", ctypes.c_float),
        ("out_f32", ctypes.c_void_p), ("out_bf16", ctypes.c_void_p), ("res", ctypes.c_void_p),
        ("aux_bf16", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("bias_mod", ctypes.c_int),
        ("o_cin", ctypes.c_int), ("o_zdiv", ctypes.c_int),
        ("o_ld", ctypes.c_longlong), ("o_s_blk", ctypes.c_longlong), ("o_s_zlo", ctypes.c_longlong),
        ("o_s_zhi", ctypes.c_longlong),
        ("lse", ctypes.c_void_p), ("delta", ctypes.c_void_p),
        ("bank_t", ctypes.c_void_p), ("bank_h", ctypes.c_void_p), ("bank_w", ctypes.c_void_p),
        ("bt", ctypes.c_int), ("bh", ctypes.c_int), ("bw", ctypes.c_int), ("heads", ctypes.c_int),
        ("a_conv", ctypes.c_int), ("b_conv", ctypes.c_int),
        ("cv_C", ctypes.c_int), ("cv_W", ctypes.c_int), ("cv_H", ctypes.c_int), ("cv_N", ctypes.c_int),
        ("cv_P", ctypes.c_int), ("cv_ntaps", ctypes.c_int),
        ("cv_pix_stride", ctypes.c_longlong), ("cv_s_phase", ctypes.c_longlong),
        ("cv_dh", ctypes.c_byte * 16), ("cv_dw", ctypes.c_byte * 16), ("cv_ph", ctypes.c_byte * 16),
        ("rowdot", ctypes.c_void_p), ("rd_block", ctypes.c_int), ("rd_L", ctypes.c_int),
        ("v", ctypes.c_void_p), ("v_cin", ctypes.c_int), ("v_zdiv", ctypes.c_int),
        ("v_ld", ctypes.c_longlong), ("v_s_zlo", ctypes.c_longlong), ("v_s_zhi", ctypes.c_longlong),
        ("o2_bf16", ctypes.c_void_p), ("o2_n", ctypes.c_int), ("o2_cin", ctypes.c_int), ("o2_zdiv", ctypes.c_int),
        ("o2_ld", ctypes.c_longlong), ("o2_s_zlo", ctypes.c_longlong), ("o2_s_zhi", ctypes.c_longlong),
        ("prof", ctypes.c_void_p),
    ]


class LvtAttnBwd(ctypes.Structure):
    """Mirror of `struct LvtAttnBwd` (include/lvt_b200.h)."""
    _fields_ = [
        ("nseq", ctypes.c_int), ("heads", ctypes.c_int),
        ("bt", ctypes.c_int), ("bh", ctypes.c_int), ("bw", ctypes.c_int),
        ("causal", ctypes.c_int), ("scale", ctypes.c_float),
        ("qkv", ctypes.c_void_p), ("qkv_ld", ctypes.c_longlong),
        ("dO", ctypes.c_void_p), ("do_ld", ctypes.c_longlong),
        ("dqkv", ctypes.c_void_p), ("dqkv_ld", ctypes.c_longlong),
        ("lse", ctypes.c_void_p), ("delta", ctypes.c_void_p),
        ("bank_t", ctypes.c_void_p), ("bank_h", ctypes.c_void_p), ("bank_w", ctypes.c_void_p),
        ("dbank_t", ctypes.c_void_p), ("dbank_h", ctypes.c_void_p), ("dbank_w", ctypes.c_void_p),
        ("scratch", ctypes.c_void_p), ("scratch_bytes", ctypes.c_longlong),
        ("prof", ctypes.c_void_p),
    ]


class LvtRowsLinear(ctypes.Structure):
    """Mirror of `struct LvtRowsLinear` (include/lvt_b200.h)."""
    _fields_ = [
        ("B", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("x", ctypes.c_void_p), ("x_ldb", ctypes.c_longlong), ("x_pos_mul", ctypes.c_longlong), ("x_bf16", ctypes.c_int),
        ("ln_gamma", ctypes.c_void_p), ("ln_beta", ctypes.c_void_p), ("ln_eps", ctypes.c_float),
        ("round_in", ctypes.c_int),
        ("w_bf16", ctypes.c_void_p), ("w_ld", ctypes.c_longlong),
        ("bias", ctypes.c_void_p),
        ("res", ctypes.c_void_p), ("res_ldb", ctypes.c_longlong), ("res_pos_mul", ctypes.c_longlong),
        ("gtab", ctypes.c_void_p), ("slice", ctypes.c_void_p), ("g_count", ctypes.c_int), ("nv", ctypes.c_int),
        ("nc", ctypes.c_int), ("thw", ctypes.c_int),
        ("relu", ctypes.c_int), ("round_out", ctypes.c_int),
        ("out", ctypes.c_void_p), ("out_ldb", ctypes.c_longlong),
        ("pos", ctypes.c_void_p),
    ]


class LvtPermuteJob(ctypes.Structure):
    """Mirror of `struct LvtPermuteJob` (include/lvt_b200.h)."""
    _fields_ = [("in_", ctypes.c_void_p), ("out", ctypes.c_void_p), ("out_is_bf16", ctypes.c_int), ("accumulate", ctypes.c_int),
                ("dims", ctypes.c_int * 4), ("in_strides", ctypes.c_longlong * 4), ("out_strides", ctypes.c_longlong * 4),
                ("first_block", ctypes.c_longlong)]


class PermuteBatch:
    """A fixed list of lvt_permute4 jobs (same arguments) replayed as ONE launch: the table is uploaded once."""

    def __init__(self):
        self.jobs, self.table, self.blocks = [], None, 0

    def add(self, src, dst, bf16, acc, dims, istr, ostr):
        ptr_of = lambda x: x if isinstance(x, int) else x.data_ptr()  # noqa: E731
        n = 1
        for d in dims:
            n *= int(d)
        if n > 0:
            self.jobs.append((ptr_of(src), ptr_of(dst), int(bf16), int(acc), tuple(int(d) for d in dims),
                              tuple(int(v) for v in istr), tuple(int(v) for v in ostr), n))

    def run(self, device):
        if not self.jobs:
            return
        if self.table is None:
            arr = (LvtPermuteJob * len(self.jobs))()
            blk = 0
            for a, (src, dst, bf16, acc, dims, istr, ostr, n) in zip(arr, self.jobs):
                a.in_, a.out, a.out_is_bf16, a.accumulate = src, dst, bf16, acc
                a.dims[:], a.in_strides[:], a.out_strides[:] = dims, istr, ostr
                a.first_block = blk
                blk += (n + 1023) // 1024
            self.blocks = blk
            self.table = torch.frombuffer(bytearray(arr), dtype=torch.uint8).to(device)
        check(load().lvt_permute4_batch(ctypes.c_void_p(self.table.data_ptr()), len(self.jobs), self.blocks, stream_ptr()),
              "lvt_permute4_batch")


class LvtDecodeLayer(ctypes.Structure):
    """Mirror of `struct LvtDecodeLayer` (include/lvt_b200.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("ln1_g", "ln1_b", "w_qkv", "k_cache", "v_cache", "bank_t", "bank_h", "bank_w",
                                               "w_proj", "ln2_g", "ln2_b", "w_ffn1", "b_ffn1", "w_ffn3", "b_ffn3")]


class LvtDecodeStep(ctypes.Structure):
    """Mirror of `struct LvtDecodeStep` (include/lvt_b200.h)."""
    _fields_ = ([(n, ctypes.c_int) for n in ("B", "d", "H", "da", "L", "nc", "nv", "de", "ntaps", "n_layers",
                                              "bt", "bh", "bw", "t", "h", "w")]
                + [(n, ctypes.c_float) for n in ("scale", "ln_eps", "temp")]
                + [("do_sample", ctypes.c_int)]
                + [(n, ctypes.c_void_p) for n in ("pos", "slice", "emb", "taps", "conv_w", "y0s")]
                + [("layer", LvtDecodeLayer * 8)]
                + [("lnp_g", ctypes.c_void_p), ("lnp_b", ctypes.c_void_p)]
                + [("U", ctypes.c_void_p * 4), ("U_ld", ctypes.c_longlong * 4), ("U_bias", ctypes.c_void_p * 4),
                   ("gtab", ctypes.c_void_p * 4), ("P", ctypes.c_void_p * 4), ("P_bias", ctypes.c_void_p * 4)]
                + [(n, ctypes.c_void_p) for n in ("q_exp", "xa", "xb", "hbuf", "a1", "q", "o", "abuf", "logits", "barrier", "prof")])


_vp, _i, _ll, _d, _f = (ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double,
                        ctypes.c_float)
# symbol -> (restype, argtypes); one entry per function include/lvt_b200.h declares
SYMBOLS = {
    "lvt_abi_version": (_i, []),
    "lvt_last_error": (ctypes.c_char_p, []),
    "lvt_device_check": (_i, []),
    "lvt_set_sm_limit": (None, [_i]),
    "lvt_launch_count": (_ll, []),
    "lvt_launch_count_reset": (None, []),
    "lvt_vt_sample_pixel": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "lvt_rows_linear": (_i, [_vp, _vp]),
    "lvt_vt_decode_step": (_i, [ctypes.POINTER(LvtDecodeStep), _vp]),
    "lvt_rows_qkv": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "lvt_attn_row": (_i, [_vp] * 6 + [_i, _i, _i, _vp, _f, _vp, _i, _i, _i, _i, _vp]),
    "lvt_vq_argmin": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "lvt_vq_ema_update": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _d, _d, _vp]),
    "lvt_vq_codebook_grad": (_i, [_vp, _vp, _vp, _vp, _f, _i, _i, _vp]),
    "lvt_vq_gather": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "lvt_gemm_bf16": (_i, [ctypes.POINTER(LvtGemm), _vp]),
    "lvt_attn_bwd_scratch_bytes": (_ll, []),
    "lvt_attn_bwd": (_i, [ctypes.POINTER(LvtAttnBwd), _vp]),
    "lvt_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "lvt_layernorm_bwd": (_i, [_vp] * 10 + [_i, _i, _vp]),
    "lvt_layernorm_bwd_bf16dy": (_i, [_vp] * 10 + [_i, _i, _vp]),
    "lvt_layernorm_bwd_ex": (_i, [_vp, _i] + [_vp] * 10 + [_i, _i, _vp]),
    "lvt_colsum_bf16": (_i, [_vp, _vp, _i, _i, _ll, _vp]),
    "lvt_attn_delta": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "lvt_relpos_bank_grad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "lvt_vt_enc_front_fwd": (_i, [_vp] * 6 + [_i] * 4 + [_vp] * 3 + [_i, _vp]),
    "lvt_vt_enc_front_bwd": (_i, [_vp] * 5 + [_i] * 4 + [_vp] * 3 + [_i, _vp]),
    "lvt_vt_dec_front_fwd": (_i, [_vp] * 4 + [_i] * 8 + [_vp]),
    "lvt_vt_dec_front_bwd": (_i, [_vp] * 4 + [_i] * 8 + [_vp]),
    "lvt_chpred_combine_fwd": (_i, [_vp] * 4 + [_i] * 6 + [_vp]),
    "lvt_chpred_combine_bwd": (_i, [_vp] * 3 + [_i] * 6 + [_vp]),
    "lvt_cross_entropy": (_i, [_vp] * 6 + [_i] * 4 + [_vp]),
    "lvt_rmsprop_step": (_i, [_vp] * 5 + [_ll] + [_f] * 5 + [_vp]),
    "lvt_adam_step": (_i, [_vp] * 5 + [_ll] + [_f] * 4 + [_i, _f, _vp]),
    "lvt_cast_bf16": (_i, [_vp, _vp, _ll, _vp]),
    "lvt_permute4": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "lvt_permute4_batch": (_i, [_vp, _i, _i, _vp]),
    "lvt_rows_gather": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "lvt_vt_class_bias": (_i, [_vp, _ll, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "lvt_rows_add_group_bias": (_i, [_vp, _vp, _ll, _i, _i, _vp]),
    "lvt_colsum_groups_bf16": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "lvt_vt_class_grad": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _vp, _i, _i, _i, _vp]),
    "lvt_vq_argmin_nhwc": (_i, [_vp] * 7 + [_i] * 5 + [_vp]),
    "lvt_vq_gather_nhwc": (_i, [_vp] * 4 + [_i] * 5 + [_vp]),
    "lvt_vqvae_in_im2col": (_i, [_vp, _vp, _i, _f, _f, _vp]),
    "lvt_split3_bf16": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "lvt_vqvae_in_im2col_split": (_i, [_vp, _vp, _i, _f, _f, _vp]),
    "lvt_vqvae_out_col2im_tanh": (_i, [_vp, _vp, _vp, _i, _vp]),
    "lvt_vqvae_recon_loss": (_i, [_vp] * 5 + [_i, _f, _f, _f, _vp]),
    "lvt_vqvae_out_convt_g": (_i, [_vp, _vp, _i, _vp]),
    "lvt_vqvae_commit_loss": (_i, [_vp] * 5 + [_ll, _f, _vp]),
    "lvt_relu_bwd_add": (_i, [_vp] * 4 + [_ll, _vp]),
    "lvt_add_bf16_to_f32": (_i, [_vp, _vp, _ll, _vp]),
    "lvt_cast_relu_bf16": (_i, [_vp, _vp, _ll, _i, _vp]),
    "lvt_denorm_clamp": (_i, [_vp, _vp, _ll, _f, _f, _f, _f, _vp]),
}


def load():
    """Load the shared library (once). Raises LvtError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LvtError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C lvt_b200/csrc`). lvt_b200 has no CPU / PyTorch fallback.")
    override = os.environ.get("LVT_B200_LIB")  # A/B timing of an older build (tools/ only)
    lib = ctypes.CDLL(override or LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        if override and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().lvt_last_error()
        raise LvtError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def require_device():
    lib = load()
    if not torch.cuda.is_available():
        raise LvtError("lvt_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    check(lib.lvt_device_check(), "lvt_device_check")
    return lib


def launch_count():
    return int(load().lvt_launch_count())


def launch_count_reset():
    load().lvt_launch_count_reset()
