"""DSFVT execution engine: the whole VideoTransformer forward / backward as a fixed sequence of
C-ABI launches (include/lvt_b200.h) over pre-allocated HBM buffers.

Reference semantics: vidgen/modeling/autoregressive/videotransformer.py:11-248,
vt_attention.py:10-202, vt_utils.py:183-200 and the loss of meta_arch/vt.py:301-314.

Data layout in HBM (M = B * t*h*w tokens of the slice, token-major everywhere; the reference's
(B,C,T,H,W) <-> (B,thw,C) transposes around every layer, vt_attention.py:183-188, disappear):
  residual stream x        fp32 [M, d]
  GEMM operands            bf16 (LayerNorm outputs, qkv [M, 3*H*da], attention probabilities
                           P [B, H, L, L], head-concat o [M, H*da], FFN hidden a1 [M, d])
  parameters               one flat fp32 master buffer (+ flat fp32 gradient, + flat bf16 shadow
                           read by the GEMMs through TMA), laid out so that w_q|w_k|w_v of a layer
                           are contiguous ([3H, d, da] blocked MN-major B operand)
No torch op runs on this path; torch owns the device memory and the stream only.
"""
import ctypes
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ... import _lib, ops
from ..._lib import check, ptr, stream_ptr
from ...ops import Operand, gemm

LN_EPS = 1e-5
_SPLIT_ATTN = os.environ.get("LVT_SPLIT_ATTN", "0") == "1"
_SPLIT_BANK = os.environ.get("LVT_SPLIT_BANK", "0") == "1"  # A/B aid: bank gradient as its own kernel on the side stream
# A/B aid: LVT_ATTN_BWD=split keeps the round-1 attention backward (P stored by the forward, dV GEMM, fused dS + dQ
# kernel with dS stored, dK GEMM) instead of the single fused kernel that recomputes P from the saved log-sum-exp
_OLD_ATTN_BWD = os.environ.get("LVT_ATTN_BWD", "fused") == "split" or _SPLIT_ATTN
_ALIGN = 64  # elements; keeps every parameter 256 B (fp32) / 128 B (bf16) aligned for TMA


class VTSpec:
    """MODEL.AUTOREGRESSIVE.VT.* (reference config/defaults.py:36-53)."""

    def __init__(self, nc=4, nv=512, kernel=(7, 1, 1), stride=(16, 1, 1), de=128, d=512, da=128,
                 blocks_e=((1, 16, 16),) * 8, heads_e=(8,) * 8, blocks_d=((1, 16, 16),) * 8,
                 heads_d=(8,) * 8, pad_value=-1, ignore_index=-100, share_p=False,
                 share_embeddings=False, class_num=0):
        self.nc, self.nv = int(nc), int(nv)
        self.kernel, self.stride = tuple(kernel), tuple(stride)
        self.de, self.d, self.da = int(de), int(d), int(da)
        self.blocks_e = tuple(tuple(b) for b in blocks_e)
        self.blocks_d = tuple(tuple(b) for b in blocks_d)
        self.heads_e, self.heads_d = tuple(heads_e), tuple(heads_d)
        self.pad_value, self.ignore_index = int(pad_value), int(ignore_index)
        self.share_p = bool(share_p)  # one P for all channels (videotransformer.py:121-123,150-151)
        # SHARE_EMBEDDINGS: one P: d -> de for all channels, then the logits against channel k's embedding table
        # (videotransformer.py:124-125,152-154): logits_k = (P relu(u_k)) E_k^T
        self.share_embeddings = bool(share_embeddings)
        if self.share_p and self.share_embeddings:
            raise _lib.LvtError("SHARE_P and SHARE_EMBEDDINGS together do not make sense (videotransformer.py:122)")
        # CLASS_NUM > 0: a class embedding concatenated to every position before the encoder's projector
        # (videotransformer.py:29-33,54-57) = a per-sample bias W[:, de:] . E_class[class]
        self.class_num = int(class_num)
        heads = set(self.heads_e) | set(self.heads_d)
        if len(heads) != 1:
            raise _lib.LvtError("all attention layers must use the same number of heads")
        self.H = heads.pop()
        blocks = set(self.blocks_e) | set(self.blocks_d)
        if len(blocks) != 1:
            raise _lib.LvtError("all attention layers must use the same block shape")
        self.block = blocks.pop()
        if self.block[0] * self.block[1] * self.block[2] != 256:
            raise _lib.LvtError("attention blocks must hold 256 positions (all shipped configs do)")

    def param_shapes(self):
        """Names / shapes / order of the reference state_dict parameters
        (videotransformer.py:11-33,62-78,104-137; vt_attention.py:98-104,132-144)."""
        s = {}
        kt, kh, kw = self.kernel
        s["encoder.conv.weight"] = (self.de, self.nc * self.nv, kt, kh, kw)
        s["encoder.conv.bias"] = (self.de,)
        s["encoder.slice_embedding.weight"] = (self.stride[0] * self.stride[1] * self.stride[2], self.de)
        if self.class_num:
            s["encoder.class_embedding.weight"] = (self.class_num, self.de)
        s["encoder.linear_projector.weight"] = (self.d, self.de * (2 if self.class_num else 1), 1, 1, 1)

        def bla(prefix):
            t, h, w = self.block
            s[prefix + "dt_bank"] = (self.H, 2 * t - 1)
            s[prefix + "dh_bank"] = (self.H, 2 * h - 1)
            s[prefix + "dw_bank"] = (self.H, 2 * w - 1)
            for n in ("w_q", "w_k", "w_v"):  # contiguous on purpose (blocked [3H, d, da] operand)
                s[prefix + "mha." + n] = (self.H, self.d, self.da)
            s[prefix + "mha.layer_norm.weight"] = (self.d,)
            s[prefix + "mha.layer_norm.bias"] = (self.d,)
            s[prefix + "mha.proj.weight"] = (self.d, self.H * self.da)
            s[prefix + "ffn.0.weight"] = (self.d,)
            s[prefix + "ffn.0.bias"] = (self.d,)
            s[prefix + "ffn.1.weight"] = (self.d, self.d)
            s[prefix + "ffn.1.bias"] = (self.d,)
            s[prefix + "ffn.3.weight"] = (self.d, self.d)
            s[prefix + "ffn.3.bias"] = (self.d,)

        for i in range(len(self.blocks_e)):
            bla(f"encoder.block_local_attention.{i}.")
        for k in range(self.nc):
            s[f"decoder.ch_embedder.{k}.weight"] = (self.nv, self.de)
        s["decoder.conv.conv.weight"] = (self.d, self.de, 3, 3, 3)
        s["decoder.conv.conv.bias"] = (self.d,)
        s["decoder.linear_projector.weight"] = (self.d, self.d, 1, 1, 1)
        for i in range(len(self.blocks_d)):
            bla(f"decoder.block_local_attention.{i}.")
        s["ch_predictor.layer_norm.weight"] = (self.d,)
        s["ch_predictor.layer_norm.bias"] = (self.d,)
        for k in range(self.nc):
            s[f"ch_predictor.U.{k}.weight"] = (self.d, self.d + k * self.nv)
            s[f"ch_predictor.U.{k}.bias"] = (self.d,)
        n_out = self.de if self.share_embeddings else self.nv
        for k in range(1 if (self.share_p or self.share_embeddings) else self.nc):
            s[self.p_name(k) + ".weight"] = (n_out, self.d)
            s[self.p_name(k) + ".bias"] = (n_out,)
        return s

    def p_name(self, k):
        """state_dict prefix of channel k's output Linear: `ch_predictor.P` when SHARE_P (one nn.Linear,
        videotransformer.py:121-123), else `ch_predictor.P.<k>` (a ModuleList, :127-130)."""
        return "ch_predictor.P" if (self.share_p or self.share_embeddings) else f"ch_predictor.P.{k}"


class ParamStore:
    """Flat fp32 master / gradient / bf16 shadow buffers with named views."""

    def __init__(self, shapes: Dict[str, Tuple[int, ...]], device):
        self.shapes = dict(shapes)
        self.offsets = {}
        off = 0
        for name, shp in shapes.items():
            self.offsets[name] = off
            n = int(np.prod(shp))
            off += (n + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.device = device
        self.master = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device)
        self.shadow = torch.zeros(off, dtype=torch.bfloat16, device=device)
        self.p = {n: self._view(self.master, n) for n in shapes}
        self.g = {n: self._view(self.grad, n) for n in shapes}

    def _view(self, flat, name):
        o = self.offsets[name]
        n = int(np.prod(self.shapes[name]))
        return flat[o:o + n].view(self.shapes[name])

    # raw device addresses
    def pf(self, name):  # fp32 master
        return self.master.data_ptr() + 4 * self.offsets[name]

    def gf(self, name):  # fp32 grad
        return self.grad.data_ptr() + 4 * self.offsets[name]

    def pb(self, name):  # bf16 shadow
        return self.shadow.data_ptr() + 2 * self.offsets[name]

    def load(self, tensors: Dict[str, torch.Tensor]):
        for n, t in tensors.items():
            if n in self.p:
                self.p[n].copy_(t.to(self.device, torch.float32).view(self.shapes[n]))


def _vp(x):
    return ctypes.c_void_p(x) if isinstance(x, int) else ptr(x)


def _i3(v):
    return (ctypes.c_int * 3)(*[int(a) for a in v])


class _Layer:
    """Saved activations of one BlockLocalAttention layer (all needed by its backward)."""
    __slots__ = ("x", "mean1", "rstd1", "ln1", "qkv", "P", "lse", "o", "h", "mean2", "rstd2", "ln2", "a1", "y",
                 "y_bf16")


class VTWorkspace:
    """All activations / scratch for a batch of B slices (allocated once per B)."""

    def __init__(self, spec: VTSpec, B: int, slice_shape, ctx_shape, ntaps, device, train=True):
        f32, bf16 = torch.float32, torch.bfloat16
        self.B = B
        self.slice_shape = tuple(slice_shape)
        self.ctx_shape = tuple(ctx_shape)
        thw = slice_shape[0] * slice_shape[1] * slice_shape[2]
        self.thw, self.M = thw, B * thw
        M, d, H, da, de, nv, nc = self.M, spec.d, spec.H, spec.da, spec.de, spec.nv, spec.nc
        L = 256
        self.nseq = M // L
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=device)  # noqa: E731
        # static inputs (graph-capturable)
        self.context = torch.zeros((B, nc) + self.ctx_shape, dtype=torch.int64, device=device)
        self.slice = torch.zeros((B, nc) + self.slice_shape, dtype=torch.int64, device=device)
        self.slice_idx = torch.zeros((B,), dtype=torch.int64, device=device)
        if spec.class_num:
            self.class_idx = torch.zeros((B,), dtype=torch.int64, device=device)
            self.cbias = torch.zeros((B, spec.d), dtype=f32, device=device)   # W[:, de:] . E_class[class] per sample
            self.csum = torch.zeros((B, spec.d), dtype=f32, device=device)    # per-sample column sums of d(projector output)
        self.ignore = torch.zeros((B, thw), dtype=torch.uint8, device=device)
        self.loss = torch.zeros((1,), dtype=f32, device=device)
        self.count = torch.zeros((1,), dtype=torch.int32, device=device)
        n_layers = len(spec.blocks_e) + len(spec.blocks_d)
        self.layers: List[_Layer] = []
        self.e0 = e((M, de), bf16)
        self.x0 = e((M, d), f32)
        self.A0 = e((M, ntaps * de), bf16)
        self.y0 = e((M, d), f32)
        n_saved = n_layers if train else 1
        for i in range(n_saved):
            ly = _Layer()
            ly.mean1, ly.rstd1, ly.mean2, ly.rstd2 = (e((M,), f32) for _ in range(4))
            ly.ln1, ly.ln2, ly.a1 = e((M, d), bf16), e((M, d), bf16), e((M, d), bf16)
            ly.qkv = e((M, 3 * H * da), bf16)
            # the attention probabilities never reach HBM: the forward keeps the row log-sum-exp, the backward
            # recomputes P from it (csrc/attn_bwd.cu)
            ly.lse = e((self.nseq * H, L), f32)
            ly.P = e((self.nseq, H, L, L), bf16) if (_OLD_ATTN_BWD and train) else None
            ly.o = e((M, H * da), bf16)
            ly.h = e((M, d), f32)
            ly.y = e((M, d), f32)
            ly.y_bf16 = None
            self.layers.append(ly)
        if not train:  # ping-pong residual buffers for inference
            self.y_alt = e((M, d), f32)
        self.zl_bf16 = e((M, d), bf16)
        # general tiled BlockLocalAttention (vt_attention.py:189-200): the slice grid is a multiple of the attention
        # block.  The layers then run on the tokens in block-major order (every 256 consecutive rows = one block);
        # LayerNorm, projections and the FFN are per token, so ONE re-ordering after each front and one back after
        # each stack replaces the reference's split / stack / permute around every layer.
        blk = spec.block
        self.tiled = self.slice_shape != tuple(blk)
        if self.tiled:
            T_, H_, W_ = self.slice_shape
            t_, h_, w_ = blk
            r = torch.arange(B * thw, dtype=torch.int64).view(B, T_ // t_, t_, H_ // h_, h_, W_ // w_, w_)
            perm = r.permute(0, 1, 3, 5, 2, 4, 6).reshape(-1)          # block-major position -> raster row
            inv = torch.empty_like(perm)
            inv[perm] = torch.arange(perm.numel())
            self.perm, self.inv = perm.to(torch.int32).to(device), inv.to(torch.int32).to(device)
            self.x0p, self.y0p, self.yf_r = e((M, d), f32), e((M, d), f32), e((M, d), f32)
            self.zl_r = e((M, d), bf16)
            if train:
                self.tmp_f, self.tmp_b, self.tmp_b2 = e((M, d), f32), e((M, d), bf16), e((M, d), bf16)
        # predictor
        self.mean_p, self.rstd_p = e((M,), f32), e((M,), f32)
        self.ln_y = e((M, d), bf16)
        self.u = e((M, d), f32)
        self.a = e((nc, M, d), bf16)
        self.logits = e((nc, M, nv), f32)
        if spec.share_embeddings:  # P relu(u_k): kept per channel for the embedding tables' gradient
            self.pe = e((nc, M, de), bf16)
            if train:
                self.dpe = e((M, de), bf16)
        if train:
            self.dlogits = e((nc, M, nv), bf16)
            self.du = e((M, d), bf16)
            self.dln = e((M, d), f32)
            self.dln_bf16 = e((M, d), bf16)  # gradient wrt a LayerNorm output, as the dgrad GEMM writes it
            self.dy, self.dy_bf16 = e((M, d), f32), e((M, d), bf16)
            self.dh, self.dh_bf16 = e((M, d), f32), e((M, d), bf16)
            self.dz1 = e((M, d), bf16)
            self.do = e((M, H * da), bf16)
            self.delta = e((self.nseq, H, L), f32)
            self.dS = e((self.nseq, H, L, L), bf16) if _OLD_ATTN_BWD else None
            self.dqkv = e((M, 3 * H * da), bf16)
            self.dA0 = e((M, ntaps * de), f32)
            self.de0, self.de0_bf16 = e((M, de), f32), e((M, de), bf16)


def layer_cuts(n, parts):
    """layer boundaries of a stack of n layers split into `parts` backward segments, top down: (8, 2) -> [8, 4, 0]"""
    p = max(1, min(parts, n))
    return [n - (n * j) // p for j in range(p + 1)]


def bucket_ranges(offsets, numel, n_enc, n_dec, parts):
    """Flat-gradient ranges [lo, hi) of the backward segments of VTEngine.backward_plan, in execution order: the
    parameters are laid out in forward order (encoder front, encoder layers, decoder front, decoder layers,
    predictor), the backward completes them back to front, so every bucket is contiguous and the buckets tile
    [0, numel) from the top down."""
    def layer_off(stack, i):
        return offsets[f"{stack}.block_local_attention.{i}.dt_bank"]

    out, hi = [], numel
    cd, ce = layer_cuts(n_dec, parts), layer_cuts(n_enc, parts)
    for j in range(len(cd) - 1):
        lo = offsets["decoder.ch_embedder.0.weight"] if j == len(cd) - 2 else layer_off("decoder", cd[j + 1])
        out.append((lo, hi))
        hi = lo
    for j in range(len(ce) - 1):
        lo = 0 if j == len(ce) - 2 else layer_off("encoder", ce[j + 1])
        out.append((lo, hi))
        hi = lo
    return out


class VTEngine:
    def __init__(self, spec: VTSpec, device="cuda"):
        _lib.require_device()
        self.spec = spec
        self.device = torch.device(device)
        self.store = ParamStore(spec.param_shapes(), self.device)
        self.lib = _lib.load()
        self._ws: Dict[Tuple, VTWorkspace] = {}
        self._posenc: Dict[Tuple, torch.Tensor] = {}
        s = spec
        ktaps = s.kernel[0] * s.kernel[1] * s.kernel[2]
        f32 = torch.float32
        # re-laid-out small weights (refreshed from the master by refresh_shadows())
        self.enc_wt = torch.zeros((s.nc, ktaps, s.nv, s.de), dtype=f32, device=self.device)
        self.enc_dwt = torch.zeros_like(self.enc_wt)
        self.ut = [None] + [torch.zeros((k * s.nv, s.d), dtype=f32, device=self.device) for k in range(1, s.nc)]
        self.dut = [None] + [torch.zeros_like(self.ut[k]) for k in range(1, s.nc)]
        self._taps_cache = {}
        self.conv_wp = None  # packed masked-conv weight, depends on the slice shape (live taps)
        self._rec, self._pbatch = None, {}
        self._side = torch.cuda.Stream()  # parameter-gradient stream of the backward pass
        self._side_pending = False
        self.shadows_fresh = False

    # ------------------------------------------------------------------ parameters
    def load_state_dict(self, tensors):
        self.store.load(tensors)
        # MaskedConv3d keeps its masked taps at zero (vt_utils.py:198-199)
        self.store.p["decoder.conv.conv.weight"][:, :, -1, -1, 1:] = 0
        self.shadows_fresh = False

    def _live_taps(self, slice_shape):
        """MaskedConv3d(de, d, (3,3,3)) taps that can touch data (vt_utils.py:183-200):
        offsets (it-2, ih-2, iw-1); taps with it == ih == 2 and iw >= 1 are masked."""
        key = tuple(slice_shape)
        if key not in self._taps_cache:
            t, h, w = slice_shape
            taps = []
            for it in range(3):
                for ih in range(3):
                    for iw in range(3):
                        if it == 2 and ih == 2 and iw >= 1:
                            continue
                        dt, dh, dw = it - 2, ih - 2, iw - 1
                        if -dt > t - 1 or -dh > h - 1 or abs(dw) > w - 1:
                            continue
                        taps.append((it, ih, iw))
            offs = torch.tensor([[it - 2, ih - 2, iw - 1] for it, ih, iw in taps], dtype=torch.int32,
                                device=self.device).contiguous()
            ntaps = len(taps)
            wp = torch.zeros((self.spec.d, ntaps * self.spec.de), dtype=torch.bfloat16, device=self.device)
            dwp = torch.zeros((self.spec.d, ntaps * self.spec.de), dtype=torch.float32, device=self.device)
            self._taps_cache[key] = (taps, offs, wp, dwp)
            self.shadows_fresh = False
        return self._taps_cache[key]

    def _permute4(self, src, dst, bf16, acc, dims, istr, ostr):
        if self._rec is not None:  # inside _batched(): recorded, launched together
            self._rec.add(src, dst, bf16, acc, dims, istr, ostr)
            return
        check(self.lib.lvt_permute4(_vp(src), _vp(dst), int(bf16), int(acc), (ctypes.c_int * 4)(*dims),
                                    (ctypes.c_longlong * 4)(*istr), (ctypes.c_longlong * 4)(*ostr),
                                    stream_ptr()), "lvt_permute4")

    def _batched(self, key, fn):
        """The lvt_permute4 calls of fn() as ONE launch (lvt_permute4_batch): recorded on first use -- pointers, shapes
        and strides are fixed per key --, replayed afterwards."""
        b = self._pbatch.get(key)
        if b is None:
            b = self._pbatch[key] = _lib.PermuteBatch()
            self._rec = b
            try:
                fn()
            finally:
                self._rec = None
        b.run(self.device)

    def refresh_shadows(self):
        """master fp32 -> bf16 shadow + the re-laid-out small weights (after load / optimizer)."""
        s, st = self.spec, self.store
        check(self.lib.lvt_cast_bf16(ptr(st.master), ptr(st.shadow), st.numel, stream_ptr()), "lvt_cast_bf16")
        self._refresh_special()
        self.shadows_fresh = True

    def _refresh_special(self):
        self._batched(("refresh", len(self._taps_cache)), self._refresh_special_jobs)

    def _refresh_special_jobs(self):
        s, st = self.spec, self.store
        ktaps = s.kernel[0] * s.kernel[1] * s.kernel[2]
        # encoder.conv.weight [de, nc*nv, ktaps] -> enc_wt [nc, ktaps, nv, de]
        self._permute4(st.pf("encoder.conv.weight"), self.enc_wt, False, False,
                       (s.nc, ktaps, s.nv, s.de), (s.nv * ktaps, 1, ktaps, s.nc * s.nv * ktaps),
                       (ktaps * s.nv * s.de, s.nv * s.de, s.de, 1))
        # U[k].weight[:, d + r] -> ut[k][r, :]
        for k in range(1, s.nc):
            ld = s.d + k * s.nv
            self._permute4(st.pf(f"ch_predictor.U.{k}.weight") + 4 * s.d, self.ut[k], False, False,
                           (1, 1, k * s.nv, s.d), (0, 0, 1, ld), (0, 0, s.d, 1))
        # decoder.conv.conv.weight [d, de, 27] -> wp [d, ntaps*de] (bf16), live taps only
        for key, (taps, offs, wp, dwp) in self._taps_cache.items():
            ntaps = len(taps)
            for q, (it, ih, iw) in enumerate(taps):
                tap = (it * 3 + ih) * 3 + iw
                self._permute4(st.pf("decoder.conv.conv.weight") + 4 * tap, wp.data_ptr() + 2 * q * s.de,
                               True, False, (1, 1, s.d, s.de), (0, 0, s.de * 27, 27), (0, 0, ntaps * s.de, 1))

    def zero_grad(self):
        self.store.grad.zero_()

    def _zero_special_grads(self):
        self.enc_dwt.zero_()
        for k in range(1, self.spec.nc):
            self.dut[k].zero_()
        for key, (taps, offs, wp, dwp) in self._taps_cache.items():
            dwp.zero_()

    def _fold_special_grads_enc(self):
        """re-laid-out gradients -> master gradient layout (+=): encoder one-hot conv."""
        self._batched(("fold_enc",), self._fold_enc_jobs)

    def _fold_enc_jobs(self):
        s, st = self.spec, self.store
        ktaps = s.kernel[0] * s.kernel[1] * s.kernel[2]
        # (dims ordered so that the WRITE side -- a read-modify-write here -- is contiguous over the fastest two dims:
        # the strided side should be the read, whose 32-byte sectors are reused out of L2)
        self._permute4(self.enc_dwt, st.gf("encoder.conv.weight"), False, True,
                       (s.de, s.nc, s.nv, ktaps), (1, ktaps * s.nv * s.de, s.de, s.nv * s.de),
                       (s.nc * s.nv * ktaps, s.nv * ktaps, ktaps, 1))

    def _fold_special_grads_pred(self):
        """... one-hot half of the predictor's U[k] (complete after the predictor backward)"""
        self._batched(("fold_pred",), self._fold_pred_jobs)

    def _fold_pred_jobs(self):
        s, st = self.spec, self.store
        for k in range(1, s.nc):
            ld = s.d + k * s.nv
            self._permute4(self.dut[k], st.gf(f"ch_predictor.U.{k}.weight") + 4 * s.d, False, True,
                           (1, 1, s.d, k * s.nv), (0, 0, 1, s.d), (0, 0, ld, 1))   # contiguous writes (see _fold_enc_jobs)

    def _fold_special_grads_conv(self, slice_shape):
        """... live taps of the masked conv (complete after the decoder front)"""
        self._batched(("fold_conv", tuple(slice_shape)), lambda: self._fold_conv_jobs(slice_shape))

    def _fold_conv_jobs(self, slice_shape):
        s, st = self.spec, self.store
        taps, offs, wp, dwp = self._live_taps(slice_shape)
        ntaps = len(taps)
        for q, (it, ih, iw) in enumerate(taps):
            tap = (it * 3 + ih) * 3 + iw
            self._permute4(dwp.data_ptr() + 4 * q * s.de, st.gf("decoder.conv.conv.weight") + 4 * tap,
                           False, True, (1, 1, s.d, s.de), (0, 0, ntaps * s.de, 1), (0, 0, s.de * 27, 27))

    # ------------------------------------------------------------------ workspaces
    def workspace(self, B, slice_shape, ctx_shape, train=True) -> VTWorkspace:
        key = (B, tuple(slice_shape), tuple(ctx_shape), train)
        if key not in self._ws:
            if any(v % b for v, b in zip(slice_shape, self.spec.block)):
                raise _lib.LvtError(f"slice {tuple(slice_shape)} must be a multiple of the attention block "
                                    f"{self.spec.block} (vt_attention.py:178-180)")
            taps = self._live_taps(slice_shape)[0]
            ws = self._ws[key] = VTWorkspace(self.spec, B, slice_shape, ctx_shape, len(taps), self.device, train)
            # the positional table tiled over the samples: it enters the masked-conv GEMM as a residual tensor, which
            # keeps that GEMM on the TMA-store epilogue (a row-modulo bias is only in the staged generic epilogue)
            ws.posb = self.posenc_table(slice_shape).repeat(B, 1).contiguous()
        return self._ws[key]

    def posenc_table(self, slice_shape):
        """PositionalEncoding (vt_attention.py:10-50) as a constant [thw, d] table."""
        key = tuple(slice_shape)
        if key not in self._posenc:
            d = self.spec.d
            n = d // 6
            inc = np.log(1.0e4 / 1.0) / n
            inv = (1.0 * torch.exp(torch.arange(n).float() * -inc))
            tab = torch.zeros((d,) + key)
            for dim in range(3):
                pos = torch.arange(key[dim], dtype=torch.float)
                stime = pos.view(-1, 1) * inv.view(1, -1)
                sig = torch.cat([torch.sin(stime), torch.cos(stime)], 1).T
                view = [2 * n, 1, 1, 1]
                view[1 + dim] = key[dim]
                tab[dim * 2 * n:(dim + 1) * 2 * n] += sig.reshape(view)
            self._posenc[key] = tab.reshape(d, -1).t().contiguous().to(self.device)
        return self._posenc[key]

    # ------------------------------------------------------------------ small launch helpers
    def _ln_fwd(self, x, g, b, y, mean, rstd, M):
        check(self.lib.lvt_layernorm_fwd(_vp(x), _vp(g), _vp(b), _vp(y), _vp(mean), _vp(rstd), M, self.spec.d,
                                         LN_EPS, stream_ptr()), "lvt_layernorm_fwd")

    def _ln_bwd(self, dy, x, mean, rstd, g, dres, dx, dxb, dg, db, M, dx_colsum=None):
        """dx_colsum: gradient buffer of the bias of the Linear that produced x (+= column sums of dx)"""
        is_bf16 = int(getattr(dy, "dtype", None) == torch.bfloat16)
        check(self.lib.lvt_layernorm_bwd_ex(_vp(dy), is_bf16, _vp(x), _vp(mean), _vp(rstd), _vp(g), _vp(dres), _vp(dx),
                                            _vp(dxb), _vp(dg), _vp(db), _vp(dx_colsum) if dx_colsum is not None else None,
                                            M, self.spec.d, stream_ptr()), "lvt_layernorm_bwd_ex")

    def _reorder(self, src, dst, idx):
        """dst[r] = src[idx[r]] over the M token rows (ws.perm: raster -> block-major, ws.inv: back)."""
        check(self.lib.lvt_rows_gather(ptr(src), ptr(dst), ptr(idx), src.shape[0], src.shape[1] * src.element_size(),
                                       stream_ptr()), "lvt_rows_gather")

    def _colsum(self, x, out, M, N, ld=None):
        check(self.lib.lvt_colsum_bf16(_vp(x), _vp(out), M, N, ld or N, stream_ptr()), "lvt_colsum_bf16")

    @staticmethod
    def _splits(m, n, k):
        """split-K factor of a weight-gradient GEMM: chosen by the library (fills the SMs / SM pairs once with
        its tile shape, at least 4 k-blocks of 64 per split; lvt_gemm_bf16, splits < 0)."""
        return -1

    def _wgrad(self, dy_ptr, ld_dy, x_ptr, ld_x, out: Operand, n_out, n_in, tokens):
        """dW[n_out, n_in] += dY[tokens, n_out]^T X[tokens, n_in] (split-K, fp32 red.add)."""
        gemm(n_out, n_in, tokens, Operand(dy_ptr, ld_dy, mn_major=True), Operand(x_ptr, ld_x, mn_major=True),
             out, out_f32=out.data, splits=self._splits(n_out, n_in, tokens), flags=ops.GEMM_ATOMIC)

    # ------------------------------------------------------------------ one attention layer
    def _qkv_op(self, buf_ptr, which, mn, L):
        s = self.spec
        ld = 3 * s.H * s.da
        return Operand(buf_ptr + 2 * which * s.H * s.da, ld, mn_major=mn, cin=s.da, zdiv=s.H, s_zlo=s.da,
                       s_zhi=L * ld)

    def _layer_fwd(self, prefix, ws: VTWorkspace, ly: _Layer, x, y, causal, y_bf16=None, keep_p=True):
        s, st = self.spec, self.store
        M, d, H, da, L = ws.M, s.d, s.H, s.da, 256
        nz = ws.nseq * H
        # pre-LN + fused QKV projection (vt_attention.py:120-124)
        self._ln_fwd(x, st.pf(prefix + "mha.layer_norm.weight"), st.pf(prefix + "mha.layer_norm.bias"),
                     ly.ln1, ly.mean1, ly.rstd1, M)
        gemm(M, 3 * H * da, d, Operand(ly.ln1.data_ptr(), d),
             Operand(st.pb(prefix + "mha.w_q"), da, mn_major=True, cin=da, s_blk=d * da),
             Operand(ly.qkv.data_ptr(), 3 * H * da), out_bf16=ly.qkv)
        # P = softmax(QK^T/sqrt(da) + B [causal -1e4]) (vt_attention.py:63-79) and O = P V (:80) in ONE kernel: P goes
        # to the second MMA through shared memory; it is written to HBM only when a backward will read it
        banks = (st.pf(prefix + "dt_bank"), st.pf(prefix + "dh_bank"), st.pf(prefix + "dw_bank"))
        if _SPLIT_ATTN:  # A/B timing aid: the two-kernel form (softmax GEMM, then P V GEMM)
            gemm(L, L, da, self._qkv_op(ly.qkv.data_ptr(), 0, False, L), self._qkv_op(ly.qkv.data_ptr(), 1, False, L),
                 Operand(ly.P.data_ptr(), L, zdiv=1, s_zhi=L * L), out_bf16=ly.P, batch=nz,
                 alpha=1.0 / math.sqrt(da), mode=ops.EPI_SOFTMAX, flags=ops.GEMM_CAUSAL if causal else 0,
                 banks=banks, block=s.block, heads=H)
            gemm(L, da, L, Operand(ly.P.data_ptr(), L, zdiv=1, s_zhi=L * L), self._qkv_op(ly.qkv.data_ptr(), 2, True, L),
                 Operand(ly.o.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da), out_bf16=ly.o,
                 batch=nz)
        else:
            gemm(L, L, da, self._qkv_op(ly.qkv.data_ptr(), 0, False, L), self._qkv_op(ly.qkv.data_ptr(), 1, False, L),
                 Operand(ly.P.data_ptr() if ly.P is not None else ly.o.data_ptr(), L, zdiv=1, s_zhi=L * L),
                 out_bf16=ly.P if (keep_p and ly.P is not None) else None, batch=nz,
                 alpha=1.0 / math.sqrt(da), mode=ops.EPI_SOFTMAX, flags=ops.GEMM_CAUSAL if causal else 0,
                 lse=ly.lse if keep_p else None,
                 banks=banks, block=s.block, heads=H, v=self._qkv_op(ly.qkv.data_ptr(), 2, True, L),
                 o2=Operand(ly.o.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da), o2_n=da)
        # output projection + residual (:127-128)
        gemm(M, d, H * da, Operand(ly.o.data_ptr(), H * da), Operand(st.pb(prefix + "mha.proj.weight"), H * da),
             Operand(ly.h.data_ptr(), d), out_f32=ly.h, res=x)
        # FFN: LN -> Linear -> ReLU -> Linear, + residual (vt_attention.py:138,186)
        self._ln_fwd(ly.h, st.pf(prefix + "ffn.0.weight"), st.pf(prefix + "ffn.0.bias"), ly.ln2, ly.mean2,
                     ly.rstd2, M)
        gemm(M, d, d, Operand(ly.ln2.data_ptr(), d), Operand(st.pb(prefix + "ffn.1.weight"), d),
             Operand(ly.a1.data_ptr(), d), out_bf16=ly.a1, bias=st.pf(prefix + "ffn.1.bias"), flags=ops.GEMM_RELU)
        gemm(M, d, d, Operand(ly.a1.data_ptr(), d), Operand(st.pb(prefix + "ffn.3.weight"), d),
             Operand(_vp(y).value, d), out_f32=y, bias=st.pf(prefix + "ffn.3.bias"), res=ly.h)
        if y_bf16 is not None:  # (a second GEMM output would put it on the staged generic epilogue: 58 us against 27 + 8)
            check(self.lib.lvt_cast_bf16(_vp(y), _vp(y_bf16), M * d, stream_ptr()), "lvt_cast_bf16")

    # Weight / bias / bank gradients are leaves of the backward graph: they go to a second stream so that
    # their CTAs fill the tails of the dgrad chain's kernels (and vice versa) instead of queueing behind them.
    def _side_begin(self):
        """side stream waits for everything issued so far on the main stream"""
        self._side.wait_stream(torch.cuda.current_stream())
        self._side_pending = True
        return torch.cuda.stream(self._side)

    def _side_join(self):
        # (only when the side stream holds work forked from this stream: during CUDA-graph capture a wait on a
        # stream outside the capture would cross the capture boundary)
        if self._side_pending:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_pending = False

    def _layer_bwd(self, prefix, ws: VTWorkspace, ly: _Layer, x, dy, dy_bf16, dx, dx_bf16, causal=False,
                   bias3_done=False, dx_colsum=None):
        """dy (fp32 + bf16 copy) = gradient wrt the layer output; writes dx (fp32 + bf16).
        Main stream: the data-gradient chain.  Side stream: parameter gradients (they only meet again in the
        optimizer); the scratch buffers they read (dy_bf16, dz1, dh_bf16, dS, dqkv) are protected by a join at
        the start of the next layer and before this layer's final LayerNorm backward overwrites dy_bf16."""
        s, st = self.spec, self.store
        M, d, H, da, L = ws.M, s.d, s.H, s.da, 256
        nz = ws.nseq * H
        scale = 1.0 / math.sqrt(da)
        dyb = _vp(dy_bf16).value
        self._side_join()  # previous layer's parameter gradients no longer read the scratch buffers
        # ---- FFN
        # (bias3_done: the LayerNorm backward that produced dy already accumulated its column sums = the ffn.3 bias
        # gradient; dx_colsum: where this layer's last LayerNorm backward puts the column sums of dx, the bias
        # gradient of the Linear that produced x)
        with self._side_begin():
            if not bias3_done:
                self._colsum(dyb, st.gf(prefix + "ffn.3.bias"), M, d)
            self._wgrad(dyb, d, ly.a1.data_ptr(), d, Operand(st.gf(prefix + "ffn.3.weight"), d), d, d, M)
            ev_dy_read = torch.cuda.Event()
            ev_dy_read.record()
        gemm(M, d, d, Operand(dyb, d), Operand(st.pb(prefix + "ffn.3.weight"), d, mn_major=True),
             Operand(ws.dz1.data_ptr(), d), out_bf16=ws.dz1, aux=ly.a1, flags=ops.GEMM_MASK)
        with self._side_begin():
            self._colsum(ws.dz1, st.gf(prefix + "ffn.1.bias"), M, d)
            self._wgrad(ws.dz1.data_ptr(), d, ly.ln2.data_ptr(), d, Operand(st.gf(prefix + "ffn.1.weight"), d), d, d, M)
        gemm(M, d, d, Operand(ws.dz1.data_ptr(), d), Operand(st.pb(prefix + "ffn.1.weight"), d, mn_major=True),
             Operand(ws.dln_bf16.data_ptr(), d), out_bf16=ws.dln_bf16)
        if dyb == ws.dh_bf16.data_ptr():
            # last encoder layer: the incoming gradient lives in ws.dh / ws.dh_bf16, which the LayerNorm backward
            # below overwrites -- the side-stream ffn.3 bias / weight gradients must have read it first
            torch.cuda.current_stream().wait_event(ev_dy_read)
        self._ln_bwd(ws.dln_bf16, ly.h, ly.mean2, ly.rstd2, st.pf(prefix + "ffn.0.weight"), dy, ws.dh, ws.dh_bf16,
                     st.gf(prefix + "ffn.0.weight"), st.gf(prefix + "ffn.0.bias"), M)
        # ---- attention output projection
        dhb = ws.dh_bf16.data_ptr()
        with self._side_begin():
            self._wgrad(dhb, d, ly.o.data_ptr(), H * da, Operand(st.gf(prefix + "mha.proj.weight"), H * da), d, H * da, M)
        # dO = dh Wproj; its epilogue also emits delta[seq, head, i] = rowsum(dO * O) (softmax backward row term)
        gemm(M, H * da, d, Operand(dhb, d), Operand(st.pb(prefix + "mha.proj.weight"), H * da, mn_major=True),
             Operand(ws.do.data_ptr(), H * da), out_bf16=ws.do, aux=ly.o, rowdot=ws.delta, rd_block=da, rd_L=L)
        # ---- softmax attention backward
        qkv, dqkv = ly.qkv.data_ptr(), ws.dqkv.data_ptr()
        gbanks = (st.gf(prefix + "dt_bank"), st.gf(prefix + "dh_bank"), st.gf(prefix + "dw_bank"))
        if not _OLD_ATTN_BWD:
            # ONE kernel: P recomputed from the saved log-sum-exp, dV = P^T dO, dS = P * (dO V^T - delta),
            # dK = scale * dS^T Q, dQ = scale * dS K and the dt/dh/dw_bank gradients; neither P nor dS reaches HBM
            banks = (st.pf(prefix + "dt_bank"), st.pf(prefix + "dh_bank"), st.pf(prefix + "dw_bank"))
            ops.attn_bwd(qkv, ws.do.data_ptr(), dqkv, ly.lse, ws.delta, banks, gbanks, ws.nseq, H, s.block, causal,
                         scale, qkv_ld=3 * H * da, do_ld=H * da)
        else:
            self._attn_bwd_split(prefix, ws, ly, qkv, dqkv, gbanks, scale)
        # ---- QKV projection
        with self._side_begin():
            gemm(d, 3 * H * da, M, Operand(ly.ln1.data_ptr(), d, mn_major=True),
                 Operand(dqkv, 3 * H * da, mn_major=True),
                 Operand(st.gf(prefix + "mha.w_q"), da, cin=da, s_blk=d * da), out_f32=st.gf(prefix + "mha.w_q"),
                 splits=self._splits(d, 3 * H * da, M), flags=ops.GEMM_ATOMIC)
        gemm(M, d, 3 * H * da, Operand(dqkv, 3 * H * da),
             Operand(st.pb(prefix + "mha.w_q"), da, mn_major=False, cin=da, s_blk=d * da),
             Operand(ws.dln_bf16.data_ptr(), d), out_bf16=ws.dln_bf16)
        torch.cuda.current_stream().wait_event(ev_dy_read)  # dx_bf16 may alias dy_bf16
        self._ln_bwd(ws.dln_bf16, x, ly.mean1, ly.rstd1, st.pf(prefix + "mha.layer_norm.weight"), ws.dh, dx, dx_bf16,
                     st.gf(prefix + "mha.layer_norm.weight"), st.gf(prefix + "mha.layer_norm.bias"), M,
                     dx_colsum=dx_colsum)

    def _attn_bwd_split(self, prefix, ws, ly, qkv, dqkv, gbanks, scale):
        """Round-1 attention backward (A/B aid, LVT_ATTN_BWD=split): P read from HBM, dS written to HBM."""
        s, st = self.spec, self.store
        M, d, H, da, L = ws.M, s.d, s.H, s.da, 256
        nz = ws.nseq * H
        P_k = Operand(ly.P.data_ptr(), L, zdiv=1, s_zhi=L * L)
        P_mn = Operand(ly.P.data_ptr(), L, mn_major=True, zdiv=1, s_zhi=L * L)
        do_k = Operand(ws.do.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da)
        do_mn = Operand(ws.do.data_ptr(), H * da, mn_major=True, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da)
        dS_k = Operand(ws.dS.data_ptr(), L, zdiv=1, s_zhi=L * L)
        dS_mn = Operand(ws.dS.data_ptr(), L, mn_major=True, zdiv=1, s_zhi=L * L)
        out_blk = lambda which: self._qkv_op(dqkv, which, False, L)  # noqa: E731
        # dV = P^T dO
        gemm(L, da, L, P_mn, do_mn, out_blk(2), out_bf16=out_blk(2).data, batch=nz)
        # dS = P * (dO V^T - delta) and dQ = scale * dS K in ONE kernel (dS reaches the second MMA through shared
        # memory; it is still written out for the dK GEMM and the bank gradient)
        # (for block (1,16,16) the same epilogue also accumulates the dt/dh/dw_bank gradients from dS)
        fused_bank = tuple(s.block) == (1, 16, 16) and not _SPLIT_BANK
        gemm(L, L, da, do_k, self._qkv_op(qkv, 2, False, L), dS_k, out_bf16=ws.dS, batch=nz, mode=ops.EPI_DS,
             aux=ly.P, delta=ws.delta, alpha=scale, v=self._qkv_op(qkv, 1, True, L), o2=out_blk(0), o2_n=da,
             banks=gbanks if fused_bank else None, block=s.block, heads=H)
        if not fused_bank:
            with self._side_begin():
                check(self.lib.lvt_relpos_bank_grad(ptr(ws.dS), _vp(gbanks[0]), _vp(gbanks[1]), _vp(gbanks[2]),
                                                    ws.nseq, H, s.block[0], s.block[1], s.block[2], stream_ptr()),
                      "lvt_relpos_bank_grad")
        # dK = scale * dS^T Q
        gemm(L, da, L, dS_mn, self._qkv_op(qkv, 0, True, L), out_blk(1), out_bf16=out_blk(1).data, batch=nz,
             alpha=scale)

    # ------------------------------------------------------------------ whole network
    def set_inputs(self, ws: VTWorkspace, context, slc, slice_idx, ignore_mask=None, class_idx=None):
        """Stage one batch into the static input buffers (H2D copies when given host tensors)."""
        if self.spec.class_num:
            if class_idx is None:
                raise _lib.LvtError("CLASS_NUM > 0: class_idx (b,) is required (videotransformer.py:54-57)")
            ws.class_idx.copy_(torch.as_tensor(class_idx).reshape(-1), non_blocking=True)
        ws.context.copy_(context.reshape(ws.context.shape), non_blocking=True)
        ws.slice.copy_(slc.reshape(ws.slice.shape), non_blocking=True)
        ws.slice_idx.copy_(slice_idx.reshape(-1), non_blocking=True)
        if ignore_mask is not None:
            ws.ignore.copy_(ignore_mask.reshape(ws.ignore.shape).to(torch.uint8), non_blocking=True)
        else:
            ws.ignore.zero_()

    def prefetch_inputs(self, ws: VTWorkspace, context, slc, slice_idx, ignore_mask=None):
        """Start the H2D copy of the NEXT batch (pinned host tensors) into a staging copy of the input buffers on a
        copy stream, while the current step still runs on the buffers the captured graphs read; commit_inputs() then
        moves it into place (device-to-device, a few microseconds).  What a DataLoader with pin_memory does for the
        reference (data/build.py)."""
        if getattr(ws, "_stage_in", None) is None:
            ws._stage_in = [torch.empty_like(t) for t in (ws.context, ws.slice, ws.slice_idx, ws.ignore)]
            ws._copy_stream = torch.cuda.Stream()
            ws._copy_done = torch.cuda.Event()
            ws._commit_done = None
        st = ws._stage_in
        # the previous commit has read the staging buffers: wait for IT, not for whatever else has been queued on the
        # compute stream since (the caller may already have launched the step that overlaps this copy)
        if ws._commit_done is not None:
            ws._copy_stream.wait_event(ws._commit_done)
        else:
            ws._copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(ws._copy_stream):
            st[0].copy_(context.reshape(st[0].shape), non_blocking=True)
            st[1].copy_(slc.reshape(st[1].shape), non_blocking=True)
            st[2].copy_(slice_idx.reshape(-1), non_blocking=True)
            if ignore_mask is not None:
                st[3].copy_(ignore_mask.reshape(st[3].shape).to(torch.uint8), non_blocking=True)
            else:
                st[3].zero_()
            ws._copy_done.record()

    def commit_inputs(self, ws: VTWorkspace):
        """Prefetched batch -> the static input buffers (ordered after the H2D copy and after the previous step)."""
        torch.cuda.current_stream().wait_event(ws._copy_done)
        for dst, src in zip((ws.context, ws.slice, ws.slice_idx, ws.ignore), ws._stage_in):
            dst.copy_(src, non_blocking=True)
        if ws._commit_done is None:
            ws._commit_done = torch.cuda.Event()
        ws._commit_done.record()

    def set_inputs_from_videos(self, ws: VTWorkspace, videos, abc, n_prime=1):
        """Device-side input construction: latent videos (B, T, nc, H, W) already in HBM + slice offsets (B, 3) ->
        the workspace's context / slice / slice_idx / ignore buffers (lvt_b200.data.prepare_slices_batched: the
        reference builds them per sample in DataLoader workers, dataset_mapper.py:113-149)."""
        from ...data.slices import prepare_slices_batched
        s = self.spec
        d = prepare_slices_batched(videos, abc, s.kernel, s.stride, n_prime, s.pad_value)
        self.set_inputs(ws, d["context"], d["slice"], d["slice_idx"], d["ignore_mask"])

    def _layer_ws(self, ws, i, train):
        return ws.layers[i] if train else ws.layers[0]

    def encoder_forward(self, ws: VTWorkspace, train=True):
        """VTEncoder.forward (videotransformer.py:35-59): ws.context, ws.slice_idx -> ws.zl_bf16."""
        s, st = self.spec, self.store
        if not self.shadows_fresh:
            self.refresh_shadows()
        M, d, de, nv, nc = ws.M, s.d, s.de, s.nv, s.nc
        nE = len(s.blocks_e)
        check(self.lib.lvt_vt_enc_front_fwd(ptr(ws.context), ptr(ws.slice_idx), ptr(self.enc_wt),
                                            _vp(st.pf("encoder.conv.bias")),
                                            _vp(st.pf("encoder.slice_embedding.weight")), ptr(ws.e0), ws.B, nc, nv,
                                            de, _i3(ws.ctx_shape), _i3(s.kernel), _i3(s.stride), s.pad_value,
                                            stream_ptr()), "lvt_vt_enc_front_fwd")
        ldw = de * (2 if s.class_num else 1)   # with classes the projector weight is (d, 2de): [W1 | W2]
        gemm(M, d, de, Operand(ws.e0.data_ptr(), de), Operand(st.pb("encoder.linear_projector.weight"), ldw),
             Operand(ws.x0.data_ptr(), d), out_f32=ws.x0)
        if s.class_num:
            w2 = ctypes.c_void_p(st.pf("encoder.linear_projector.weight") + 4 * de)
            check(self.lib.lvt_vt_class_bias(w2, ldw, _vp(st.pf("encoder.class_embedding.weight")), ptr(ws.class_idx),
                                             ptr(ws.cbias), ws.B, d, de, stream_ptr()), "lvt_vt_class_bias")
            check(self.lib.lvt_rows_add_group_bias(ptr(ws.x0), ptr(ws.cbias), M, d, ws.thw, stream_ptr()),
                  "lvt_rows_add_group_bias")
        x = ws.x0
        if ws.tiled:
            self._reorder(ws.x0, ws.x0p, ws.perm)
            x = ws.x0p
        for i in range(nE):
            ly = self._layer_ws(ws, i, train)
            y = ly.y if train else (ws.y_alt if x is ly.y else ly.y)
            self._layer_fwd(f"encoder.block_local_attention.{i}.", ws, ly, x, y, causal=False,
                            y_bf16=ws.zl_bf16 if i == nE - 1 else None, keep_p=train)
            x = y
        if ws.tiled:  # the decoder front adds W_lp zl per token of the raster grid
            self._reorder(ws.zl_bf16, ws.zl_r, ws.inv)

    def decoder_forward(self, ws: VTWorkspace, train=True):
        """VTDecoder.forward (videotransformer.py:91-101): ws.slice, ws.zl_bf16 -> ws.y_final, ws.ln_y."""
        s, st = self.spec, self.store
        M, d, de, nv, nc = ws.M, s.d, s.de, s.nv, s.nc
        t, h, w = ws.slice_shape
        taps, offs, wp, dwp = self._live_taps(ws.slice_shape)
        if not self.shadows_fresh:
            self.refresh_shadows()
        ntaps = len(taps)
        nE, nD = len(s.blocks_e), len(s.blocks_d)
        check(self.lib.lvt_vt_dec_front_fwd(ptr(ws.slice), _vp(st.pf("decoder.ch_embedder.0.weight")), ptr(offs),
                                            ptr(ws.A0), ws.B, nc, nv, de, t, h, w, ntaps, stream_ptr()),
              "lvt_vt_dec_front_fwd")
        gemm(M, d, ntaps * de, Operand(ws.A0.data_ptr(), ntaps * de), Operand(wp.data_ptr(), ntaps * de),
             Operand(ws.y0.data_ptr(), d), out_f32=ws.y0, res=ws.posb)
        zl = ws.zl_r if ws.tiled else ws.zl_bf16
        gemm(M, d, d, Operand(zl.data_ptr(), d), Operand(st.pb("decoder.linear_projector.weight"), d),
             Operand(ws.y0.data_ptr(), d), out_f32=ws.y0, res=ws.y0, bias=st.pf("decoder.conv.conv.bias"))
        x = ws.y0
        if ws.tiled:
            self._reorder(ws.y0, ws.y0p, ws.perm)
            x = ws.y0p
        for i in range(nD):
            ly = self._layer_ws(ws, nE + i, train)
            y = ly.y if train else (ws.y_alt if x is ly.y else ly.y)
            self._layer_fwd(f"decoder.block_local_attention.{i}.", ws, ly, x, y, causal=True, keep_p=train)
            x = y
        if ws.tiled:  # the predictor and the loss index tokens in raster order (ws.slice, ws.ignore)
            self._reorder(x, ws.yf_r, ws.inv)
            x = ws.yf_r
        ws.y_final = x
        self._ln_fwd(x, st.pf("ch_predictor.layer_norm.weight"), st.pf("ch_predictor.layer_norm.bias"), ws.ln_y,
                     ws.mean_p, ws.rstd_p, M)

    def predictor_forward(self, ws: VTWorkspace, channels=None):
        """ChannelPredictor (videotransformer.py:144-160) for the given channels: ws.ln_y, ws.slice -> ws.logits[k]."""
        s, st = self.spec, self.store
        M, d, nv, nc = ws.M, s.d, s.nv, s.nc
        for k in (range(nc) if channels is None else channels):
            ld = d + k * nv
            gemm(M, d, d, Operand(ws.ln_y.data_ptr(), d), Operand(st.pb(f"ch_predictor.U.{k}.weight"), ld),
                 Operand(ws.u.data_ptr(), d), out_f32=ws.u, bias=st.pf(f"ch_predictor.U.{k}.bias"))
            check(self.lib.lvt_chpred_combine_fwd(ptr(ws.u), ptr(self.ut[k]) if k else None, ptr(ws.slice),
                                                  ptr(ws.a[k]), M, nc, nv, d, ws.thw, k, stream_ptr()),
                  "lvt_chpred_combine_fwd")
            if s.share_embeddings:
                de = s.de
                gemm(M, de, d, Operand(ws.a[k].data_ptr(), d), Operand(st.pb("ch_predictor.P.weight"), d),
                     Operand(ws.pe[k].data_ptr(), de), out_bf16=ws.pe[k], bias=st.pf("ch_predictor.P.bias"))
                gemm(M, nv, de, Operand(ws.pe[k].data_ptr(), de), Operand(st.pb(f"decoder.ch_embedder.{k}.weight"), de),
                     Operand(ws.logits[k].data_ptr(), nv), out_f32=ws.logits[k])
                continue
            gemm(M, nv, d, Operand(ws.a[k].data_ptr(), d), Operand(st.pb(s.p_name(k) + ".weight"), d),
                 Operand(ws.logits[k].data_ptr(), nv), out_f32=ws.logits[k], bias=st.pf(s.p_name(k) + ".bias"))

    def forward(self, ws: VTWorkspace, train=True, want_loss=True):
        """VideoTransformer.forward(mode="logits") (videotransformer.py:232-239) [+ the CE loss of
        meta_arch/vt.py:305-312].  Results: ws.logits [nc, M, nv] fp32, ws.loss."""
        s = self.spec
        self.encoder_forward(ws, train)
        self.decoder_forward(ws, train)
        self.predictor_forward(ws)
        if want_loss:
            check(self.lib.lvt_cross_entropy(ptr(ws.logits), ptr(ws.slice), ptr(ws.ignore),
                                             ptr(ws.dlogits) if train else None, ptr(ws.loss), ptr(ws.count),
                                             ws.B, s.nc, s.nv, ws.thw, stream_ptr()), "lvt_cross_entropy")
        return ws.loss

    def backward(self, ws: VTWorkspace):
        """Gradient of ws.loss wrt every parameter, ACCUMULATED into the flat gradient buffer."""
        self.backward_decoder(ws)
        self.backward_encoder(ws)

    def grad_buckets(self):
        """The flat gradient as two contiguous buckets in the order the backward completes them:
        [decoder + channel predictor] (final after backward_decoder), [encoder] (final after backward_encoder)."""
        off = self.store.offsets["decoder.ch_embedder.0.weight"]
        return self.store.grad[off:], self.store.grad[:off]

    def backward_plan(self, ws: VTWorkspace, parts=2):
        """The backward as 2 * parts segments in execution order, [(callable, lo, hi)]: after segment i has run, the
        slice [lo, hi) of the flat gradient is final (the parameters are laid out in forward order, so the buckets
        are contiguous and complete back to front) and may be all-reduced while the next segments run."""
        s = self.spec
        nE, nD = len(s.blocks_e), len(s.blocks_d)
        ranges = bucket_ranges(self.store.offsets, self.store.numel, nE, nD, parts)
        cd, ce = layer_cuts(nD, parts), layer_cuts(nE, parts)
        segs = []
        for j in range(len(cd) - 1):
            def seg(top=cd[j], bot=cd[j + 1], first=j == 0, last=j == len(cd) - 2):
                if first:
                    self._bwd_predictor(ws)
                self._bwd_dec_layers(ws, top, bot)
                if last:
                    self._bwd_dec_front(ws)
            segs.append(seg)
        for j in range(len(ce) - 1):
            def seg(top=ce[j], bot=ce[j + 1], last=j == len(ce) - 2):
                self._bwd_enc_layers(ws, top, bot)
                if last:
                    self._bwd_enc_front(ws)
            segs.append(seg)
        return [(fn, lo, hi) for fn, (lo, hi) in zip(segs, ranges)]

    def backward_decoder(self, ws: VTWorkspace):
        """Channel predictor, decoder stack, decoder front; leaves d(loss)/d(zl) in ws.dh / ws.dh_bf16."""
        self._bwd_predictor(ws)
        self._bwd_dec_layers(ws, len(self.spec.blocks_d), 0)
        self._bwd_dec_front(ws)

    def backward_encoder(self, ws: VTWorkspace):
        """Encoder stack and encoder front (gradient entering through ws.dh / ws.dh_bf16)."""
        self._bwd_enc_layers(ws, len(self.spec.blocks_e), 0)
        self._bwd_enc_front(ws)

    def _bwd_predictor(self, ws: VTWorkspace):
        s, st = self.spec, self.store
        M, d, nv, nc = ws.M, s.d, s.nv, s.nc
        nD = len(s.blocks_d)
        self._zero_special_grads()
        for k in range(nc):
            ld = d + k * nv
            dl = ws.dlogits[k].data_ptr()
            if s.share_embeddings:
                # logits_k = pe_k E_k^T, pe_k = a_k P^T + b: dE_k += dl^T pe_k (next to what the decoder front adds to the
                # same table), dpe = dl E_k, dP += dpe^T a_k (all channels into one buffer), da = (dpe P) * [a_k > 0]
                de = s.de
                ek = f"decoder.ch_embedder.{k}.weight"
                self._wgrad(dl, nv, ws.pe[k].data_ptr(), de, Operand(st.gf(ek), de), nv, de, M)
                gemm(M, de, nv, Operand(dl, nv), Operand(st.pb(ek), de, mn_major=True),
                     Operand(ws.dpe.data_ptr(), de), out_bf16=ws.dpe)
                self._colsum(ws.dpe, st.gf("ch_predictor.P.bias"), M, de)
                self._wgrad(ws.dpe.data_ptr(), de, ws.a[k].data_ptr(), d, Operand(st.gf("ch_predictor.P.weight"), d), de, d, M)
                gemm(M, d, de, Operand(ws.dpe.data_ptr(), de), Operand(st.pb("ch_predictor.P.weight"), d, mn_major=True),
                     Operand(ws.du.data_ptr(), d), out_bf16=ws.du, aux=ws.a[k], flags=ops.GEMM_MASK)
            else:
                # (with SHARE_P the four channels add into the same gradient: both kernels accumulate with red.add)
                self._colsum(dl, st.gf(s.p_name(k) + ".bias"), M, nv)
                self._wgrad(dl, nv, ws.a[k].data_ptr(), d, Operand(st.gf(s.p_name(k) + ".weight"), d), nv, d, M)
                gemm(M, d, nv, Operand(dl, nv), Operand(st.pb(s.p_name(k) + ".weight"), d, mn_major=True),
                     Operand(ws.du.data_ptr(), d), out_bf16=ws.du, aux=ws.a[k], flags=ops.GEMM_MASK)
            self._colsum(ws.du, st.gf(f"ch_predictor.U.{k}.bias"), M, d)
            self._wgrad(ws.du.data_ptr(), d, ws.ln_y.data_ptr(), d,
                        Operand(st.gf(f"ch_predictor.U.{k}.weight"), ld), d, d, M)
            check(self.lib.lvt_chpred_combine_bwd(ptr(ws.du), ptr(ws.slice), ptr(self.dut[k]) if k else None, M, nc,
                                                  nv, d, ws.thw, k, stream_ptr()), "lvt_chpred_combine_bwd")
            gemm(M, d, d, Operand(ws.du.data_ptr(), d),
                 Operand(st.pb(f"ch_predictor.U.{k}.weight"), ld, mn_major=True),
                 Operand(ws.dln.data_ptr(), d), out_f32=ws.dln, res=ws.dln if k else None)
        self._fold_special_grads_pred()
        dy, dyb = (ws.tmp_f, ws.tmp_b) if ws.tiled else (ws.dy, ws.dy_bf16)
        self._ln_bwd(ws.dln, ws.y_final, ws.mean_p, ws.rstd_p, st.pf("ch_predictor.layer_norm.weight"), None,
                     dy, dyb, st.gf("ch_predictor.layer_norm.weight"),
                     st.gf("ch_predictor.layer_norm.bias"), M,
                     dx_colsum=st.gf(f"decoder.block_local_attention.{nD - 1}.ffn.3.bias"))
        if ws.tiled:  # gradient wrt the decoder stack's output, back in block-major order
            self._reorder(dy, ws.dy, ws.perm)
            self._reorder(dyb, ws.dy_bf16, ws.perm)

    def _bwd_dec_layers(self, ws: VTWorkspace, top, bot):
        """decoder layers top-1 .. bot"""
        st = self.store
        nE = len(self.spec.blocks_e)
        for i in reversed(range(bot, top)):
            ly = ws.layers[nE + i]
            x = ws.layers[nE + i - 1].y if i > 0 else (ws.y0p if ws.tiled else ws.y0)
            # every LayerNorm backward also emits the column sums of its dx: the bias gradient of whatever Linear
            # produced its input (ffn.3 of the layer below, or the masked conv for layer 0)
            below = st.gf(f"decoder.block_local_attention.{i - 1}.ffn.3.bias") if i > 0 else st.gf("decoder.conv.conv.bias")
            self._layer_bwd(f"decoder.block_local_attention.{i}.", ws, ly, x, ws.dy, ws.dy_bf16, ws.dy, ws.dy_bf16,
                            causal=True, bias3_done=True, dx_colsum=below)
        self._side_join()

    def _bwd_dec_front(self, ws: VTWorkspace):
        """decoder front: y0 = conv(emb) + posenc + bias + zl Wlp^T"""
        s, st = self.spec, self.store
        M, d, de, nv, nc = ws.M, s.d, s.de, s.nv, s.nc
        t, h, w = ws.slice_shape
        taps, offs, wp, dwp = self._live_taps(ws.slice_shape)
        ntaps = len(taps)
        if ws.tiled:  # dy0 arrives in block-major order; everything below is on the raster grid
            self._reorder(ws.dy_bf16, ws.tmp_b, ws.inv)
        dyb = (ws.tmp_b if ws.tiled else ws.dy_bf16).data_ptr()
        zl = ws.zl_r if ws.tiled else ws.zl_bf16
        self._wgrad(dyb, d, zl.data_ptr(), d, Operand(st.gf("decoder.linear_projector.weight"), d), d, d, M)
        self._wgrad(dyb, d, ws.A0.data_ptr(), ntaps * de, Operand(dwp.data_ptr(), ntaps * de), d, ntaps * de, M)
        gemm(M, ntaps * de, d, Operand(dyb, d), Operand(wp.data_ptr(), ntaps * de, mn_major=True),
             Operand(ws.dA0.data_ptr(), ntaps * de), out_f32=ws.dA0)
        check(self.lib.lvt_vt_dec_front_bwd(ptr(ws.slice), ptr(ws.dA0), ptr(offs),
                                            _vp(st.gf("decoder.ch_embedder.0.weight")), ws.B, nc, nv, de, t, h, w,
                                            ntaps, stream_ptr()), "lvt_vt_dec_front_bwd")
        # gradient entering the encoder stack: dzl = dy0 Wlp_d
        # (written to the dh buffers: the GEMM may not overwrite its own A operand)
        dh, dhb = (ws.tmp_f, ws.tmp_b2) if ws.tiled else (ws.dh, ws.dh_bf16)
        # (fp32 through the TMA-store epilogue, then one cast: a GEMM with two outputs runs the staged generic epilogue)
        gemm(M, d, d, Operand(dyb, d), Operand(st.pb("decoder.linear_projector.weight"), d, mn_major=True),
             Operand(dh.data_ptr(), d), out_f32=dh)
        check(self.lib.lvt_cast_bf16(ptr(dh), ptr(dhb), dh.numel(), stream_ptr()), "lvt_cast_bf16")
        if ws.tiled:  # the encoder stack's output is in block-major order
            self._reorder(dh, ws.dh, ws.perm)
            self._reorder(dhb, ws.dh_bf16, ws.perm)
        self._fold_special_grads_conv(ws.slice_shape)

    def _bwd_enc_layers(self, ws: VTWorkspace, top, bot):
        """encoder layers top-1 .. bot"""
        st = self.store
        nE = len(self.spec.blocks_e)
        for i in reversed(range(bot, top)):
            ly = ws.layers[i]
            x = ws.layers[i - 1].y if i > 0 else (ws.x0p if ws.tiled else ws.x0)
            first = i == nE - 1
            below = st.gf(f"encoder.block_local_attention.{i - 1}.ffn.3.bias") if i > 0 else None
            self._layer_bwd(f"encoder.block_local_attention.{i}.", ws, ly, x, ws.dh if first else ws.dy,
                            ws.dh_bf16 if first else ws.dy_bf16, ws.dy, ws.dy_bf16, bias3_done=not first,
                            dx_colsum=below)
        self._side_join()

    def _bwd_enc_front(self, ws: VTWorkspace):
        """encoder front: x0 = e0 Wlp_e^T ; e0 = gather-sum + bias + slice_emb"""
        s, st = self.spec, self.store
        M, d, de, nv, nc = ws.M, s.d, s.de, s.nv, s.nc
        if ws.tiled:
            self._reorder(ws.dy_bf16, ws.tmp_b, ws.inv)
        dxb = (ws.tmp_b if ws.tiled else ws.dy_bf16).data_ptr()
        ldw = de * (2 if s.class_num else 1)
        self._wgrad(dxb, d, ws.e0.data_ptr(), de, Operand(st.gf("encoder.linear_projector.weight"), ldw), d, de, M)
        if s.class_num:  # the class term: per-sample column sums of dx0, then two tiny contractions
            w2 = ctypes.c_void_p(st.pf("encoder.linear_projector.weight") + 4 * de)
            dw2 = ctypes.c_void_p(st.gf("encoder.linear_projector.weight") + 4 * de)
            check(self.lib.lvt_colsum_groups_bf16(ctypes.c_void_p(dxb), ptr(ws.csum), ws.B, ws.thw, d, stream_ptr()),
                  "lvt_colsum_groups_bf16")
            check(self.lib.lvt_vt_class_grad(ptr(ws.csum), _vp(st.pf("encoder.class_embedding.weight")),
                                             ptr(ws.class_idx), w2, dw2, ldw,
                                             _vp(st.gf("encoder.class_embedding.weight")), ws.B, d, de, stream_ptr()),
                  "lvt_vt_class_grad")
        gemm(M, de, d, Operand(dxb, d), Operand(st.pb("encoder.linear_projector.weight"), ldw, mn_major=True),
             Operand(ws.de0.data_ptr(), de), out_f32=ws.de0)
        check(self.lib.lvt_cast_bf16(ptr(ws.de0), ptr(ws.de0_bf16), ws.de0.numel(), stream_ptr()), "lvt_cast_bf16")
        self._colsum(ws.de0_bf16, st.gf("encoder.conv.bias"), M, de)
        check(self.lib.lvt_vt_enc_front_bwd(ptr(ws.context), ptr(ws.slice_idx), ptr(ws.de0), ptr(self.enc_dwt),
                                            _vp(st.gf("encoder.slice_embedding.weight")), ws.B, nc, nv, de,
                                            _i3(ws.ctx_shape), _i3(s.kernel), _i3(s.stride), s.pad_value,
                                            stream_ptr()), "lvt_vt_enc_front_bwd")
        self._fold_special_grads_enc()

    # ------------------------------------------------------------------ optimizer
    def init_optimizer(self, name="rmsprop", lr=2e-5, alpha=0.95, momentum=0.9, eps=1e-8, betas=(0.9, 0.9)):
        """torch.optim.RMSprop / Adam state over the flat buffers (reference solver/build.py:62-72)."""
        self.opt = dict(name=name, lr=lr, alpha=alpha, momentum=momentum, eps=eps, betas=betas, step=0)
        self.opt_s1 = torch.zeros_like(self.store.master)
        self.opt_s2 = torch.zeros_like(self.store.master)

    def optimizer_step(self, grad_scale=1.0):
        self.optimizer_kernel(grad_scale)
        self._refresh_special()
        self.shadows_fresh = True

    def optimizer_kernel(self, grad_scale=1.0):
        """The fused multi-tensor update alone.  lr and Adam's bias corrections are passed BY VALUE, so this launch
        must stay outside CUDA graphs (a captured launch would freeze the schedule and the step count)."""
        self.opt["step"] += 1
        self.optimizer_slice(0, self.store.numel, grad_scale)

    def optimizer_slice(self, lo, hi, grad_scale=1.0):
        """The update of the parameters [lo, hi) of the flat buffers (the update is elementwise, so a gradient bucket
        can be applied as soon as it has been reduced); the caller advances opt["step"] once per step."""
        o, st = self.opt, self.store
        at = lambda t, esize: ctypes.c_void_p(t.data_ptr() + esize * lo)  # noqa: E731
        if o["name"] == "rmsprop":
            check(self.lib.lvt_rmsprop_step(at(st.master, 4), at(st.grad, 4), at(self.opt_s1, 4), at(self.opt_s2, 4),
                                            at(st.shadow, 2), hi - lo, o["lr"], o["alpha"], o["momentum"], o["eps"],
                                            grad_scale, stream_ptr()), "lvt_rmsprop_step")
        else:
            check(self.lib.lvt_adam_step(at(st.master, 4), at(st.grad, 4), at(self.opt_s1, 4), at(self.opt_s2, 4),
                                         at(st.shadow, 2), hi - lo, o["lr"], o["betas"][0], o["betas"][1], o["eps"],
                                         o["step"], grad_scale, stream_ptr()), "lvt_adam_step")

    def train_step(self, ws: VTWorkspace, grad_hook=None, grad_scale=1.0):
        """forward + backward + optimizer on the batch staged in `ws` (trainer.py:79-87)."""
        self.zero_grad()
        self.forward(ws, train=True)
        self.backward(ws)
        if grad_hook is not None:
            grad_hook(self.store.grad)  # data-parallel all-reduce of the flat gradient
        self.optimizer_step(grad_scale)
        return ws.loss


class GraphedTrainStep:
    """The DSFVT train step (zero_grad + forward + backward [+ gradient all-reduce] + optimizer)
    captured once into CUDA graphs and replayed: one host launch per step instead of ~700.
    With world_size > 1 the flat fp32 gradient is summed across ranks (reference: DDP over self.model,
    meta_arch/vt.py:61-63) in 2 * parts buckets, back to front as the backward completes them
    (VTEngine.backward_plan): bucket i is all-reduced on a communication stream WHILE the segments after it
    run; only the last bucket (the bottom encoder layers) is exposed.  The overlapped segments are captured with
    an SM budget of (all - comm_sms) for the persistent GEMM / attention kernels (lvt_set_sm_limit): NCCL's CTAs
    hold their SMs for the whole collective, and a persistent grid sized for every SM would run its last CTAs as
    a second wave.  The optimizer kernel's grad_scale averages."""

    def __init__(self, engine: VTEngine, ws: VTWorkspace, world_size=1, allreduce=None, overlap=True, parts=None,
                 comm_sms=None):
        self.engine, self.ws = engine, ws
        self.world_size = world_size
        self.allreduce = allreduce
        self.overlap = overlap  # False: one all-reduce of the whole flat gradient after the backward
        self.parts = int(os.environ.get("LVT_GRAD_PARTS", "8")) if parts is None else parts
        self.comm_sms = int(os.environ.get("LVT_COMM_SMS", "8")) if comm_sms is None else comm_sms
        self.graphs = []     # [zero-grad + forward + segment 0, segment 1, ...]
        self.plan = None
        self.g_opt = None
        # one GPU: LVT_OPT_OVERLAP=1 uses the same segmentation to run the optimizer update of a bucket on the side
        # stream while the backward segments below it are still going; measured neutral (9.58-9.72 against 9.61 ms:
        # the update is an HBM pass that competes with the backward), so one graph + one optimizer launch stays
        self.segmented = allreduce is not None or (overlap and os.environ.get("LVT_OPT_OVERLAP", "0") == "1")
        if allreduce is None and parts is None:
            self.parts = int(os.environ.get("LVT_GRAD_PARTS", "2"))
        self.comm = torch.cuda.Stream() if self.segmented else None
        self.launches_per_step = 0

    def _fwd_bwd(self):
        self.engine.zero_grad()
        self.engine.forward(self.ws, train=True)
        self.engine.backward(self.ws)

    def _opt(self):
        """eager: lr / bias corrections are by-value kernel arguments (LR schedules, Adam step count)"""
        self.engine.optimizer_kernel(grad_scale=1.0 / self.world_size)

    def _refresh(self):
        self.engine._refresh_special()
        self.engine.shadows_fresh = True

    def _snapshot(self):
        eng = self.engine
        return (eng.store.master.clone(), eng.opt_s1.clone(), eng.opt_s2.clone(), eng.opt["step"])

    def _restore(self, snap):
        eng = self.engine
        eng.store.master.copy_(snap[0]); eng.opt_s1.copy_(snap[1]); eng.opt_s2.copy_(snap[2])
        eng.opt["step"] = snap[3]
        eng.refresh_shadows()

    def _run(self, segments):
        """segments[i](): [zero-grad + forward +] backward segment i; the all-reduce of bucket i goes to the
        communication stream as soon as the segment has been issued, and the optimizer update of that bucket's
        parameters (elementwise, eager launch) follows it there: when the backward ends only the last bucket's
        reduction and update are still outstanding"""
        eng = self.engine
        grad = eng.store.grad
        eng.opt["step"] += 1
        if not self.overlap:
            for seg in segments:
                seg()
            if self.allreduce is not None:
                self.allreduce(grad)
            eng.optimizer_slice(0, eng.store.numel, 1.0 / self.world_size)
            return
        main = torch.cuda.current_stream()
        for seg, (_, lo, hi) in zip(segments, self.plan):
            seg()
            self.comm.wait_stream(main)
            with torch.cuda.stream(self.comm):
                if self.allreduce is not None:
                    self.allreduce(grad[lo:hi])
                eng.optimizer_slice(lo, hi, 1.0 / self.world_size)
        main.wait_stream(self.comm)

    def capture(self, warmup=2):
        """Warm-up (kernel attributes, TMA maps, NCCL channels) runs real steps on whatever is staged in the
        workspace; parameters and optimizer state are snapshotted before and restored after, so capturing
        does not train."""
        eng, lib = self.engine, self.engine.lib
        dist = self.segmented
        if dist:
            self.plan = eng.backward_plan(self.ws, self.parts)

            def first():
                eng.zero_grad()
                eng.forward(self.ws, train=True)
                self.plan[0][0]()
            eager = [first] + [p[0] for p in self.plan[1:]]
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):  # eager warm-up: sets kernel attributes, builds TMA maps
                if dist:
                    self._run(eager)
                else:
                    self._fwd_bwd()
                    self._opt()
                self._refresh()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(snap)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        self.graphs = []
        if dist:
            sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
            for i, fn in enumerate(eager):
                # segments after the first run next to the all-reduce of the bucket before them
                lib.lvt_set_sm_limit(sms - self.comm_sms if (i > 0 and self.overlap and self.comm_sms > 0 and
                                                             self.allreduce is not None) else 0)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                self.graphs.append(g)
            lib.lvt_set_sm_limit(0)
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._fwd_bwd()
            self.graphs.append(g)
        self.g_opt = torch.cuda.CUDAGraph()  # weight re-layouts after the (eager) optimizer kernel
        with torch.cuda.graph(self.g_opt):
            self._refresh()
        # + the eager optimizer launches (one per gradient bucket when the update follows each all-reduce)
        self.launches_per_step = _lib.launch_count() - n0 + (len(self.plan) if (dist and self.overlap) else 1)
        torch.cuda.synchronize()

    def step(self):
        if self.segmented:
            self._run([g.replay for g in self.graphs])
        else:
            self.graphs[0].replay()
            self._opt()
        self.g_opt.replay()
        return self.ws.loss
