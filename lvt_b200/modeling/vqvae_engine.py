"""VQ-VAE execution engine (PR-DVQVAE2 / K-DVQVAE): ResEncoder -> DVQ codebook (EMA) -> ResDecoder,
forward / backward / Adam as C-ABI launches over pre-allocated HBM buffers.

Reference semantics: vidgen/modeling/encoder/resencoder.py:10-76, generator/resdecoder.py:10-75,
vq/vq_embedding.py:9-99, vq/vq_utils.py:5-65, meta_arch/ae.py:34-36,120-168, meta_arch/vqvae.py:66-106,
loss/loss.py:5-20.

Layout: activations are channels-last bf16; the 32x32 feature maps are stored "phase-major"
([hp][wp][n][16][16][C]) so that the stride-2 convolutions / transposed convolutions become plain
multi-tap implicit GEMMs over 16x16 grids (per-tap shifted TMA boxes, zero-filled padding); z_e is
fp32 [n*256, 256] for the exact codebook search.  Conv weights keep the reference layout in the
fp32 master buffer and are re-packed per step ([co][tap][ci] bf16, and [ci][tap][co] for the data
gradient) with lvt_permute4.
"""
import ctypes
from typing import Dict, Tuple

import numpy as np
import torch

from .. import _lib, ops
from .._lib import check, ptr, stream_ptr
from ..ops import ConvSpec, Operand, gemm
from .autoregressive.vt_engine import ParamStore, _vp

K4S2 = {0: (1, -1), 1: (0, 0), 2: (1, 0), 3: (0, 1)}  # Conv2d(k4,s2,p1): kh -> (input parity, shift)
TAPS3 = [(kh - 1, kw - 1, 0) for kh in range(3) for kw in range(3)]
TAPS3_T = [(1 - kh, 1 - kw, 0) for kh in range(3) for kw in range(3)]
TAPS_K4S2 = [(K4S2[kh][1], K4S2[kw][1], K4S2[kh][0] * 2 + K4S2[kw][0]) for kh in range(4) for kw in range(4)]


def phase_taps(ph, pw):
    """ConvTranspose2d(k4,s2,p1) output parity (ph,pw) == data gradient of Conv2d(k4,s2,p1) wrt input
    parity (ph,pw): tap ids (kh*4+kw) and their (dh, dw) shifts on the 16x16 grid."""
    khs = [kh for kh in range(4) if (kh - ph - 1) % 2 == 0]
    kws = [kw for kw in range(4) if (kw - pw - 1) % 2 == 0]
    ids = [kh * 4 + kw for kh in khs for kw in kws]
    shifts = [((ph + 1 - kh) // 2, (pw + 1 - kw) // 2, 0) for kh in khs for kw in kws]
    return ids, shifts


class VQVAESpec:
    """configs/vqvae/{Base-VQVAE,PR-DVQVAE2,K-DVQVAE}.yaml values."""

    def __init__(self, in_channels=3, nf=256, res_channels=128, n_layers=2, codebook_num=4, codebook_size=512,
                 codebook_dim=256, beta=1.0, ema=True, ema_decay=0.99, ema_eps=1e-5, pixel_lambda=1.0,
                 pixel_mean=0.5, pixel_std=0.5, out_activation="tanh"):
        if (in_channels, nf, res_channels, codebook_dim, out_activation) != (3, 256, 128, 256, "tanh"):
            raise _lib.LvtError("lvt_b200 implements the shipped DVQ-VAE configs: 3->128->256 channels, "
                                "RES_CHANNELS 128, CODEBOOK.DIM 256, tanh output")
        # MODEL.CODEBOOK.EMA (every shipped config: True).  False: the codebook is a trained parameter -- extra loss
        # mse(z_q, sg[z_e]) (vqvae.py:84-85), its gradient through index_select, Adam with the generator's settings
        self.ema = bool(ema)
        self.in_channels, self.nf, self.rc, self.n_layers = in_channels, nf, res_channels, n_layers
        self.num, self.K, self.D = codebook_num, codebook_size, codebook_dim // codebook_num
        self.beta, self.ema_decay, self.ema_eps, self.pixel_lambda = beta, ema_decay, ema_eps, pixel_lambda
        self.mean, self.std = pixel_mean, pixel_std

    def param_shapes(self):
        """netE / netG parameter names (prefix E. / G.) and shapes (resencoder.py:46-62, resdecoder.py:48-57)."""
        nf, rc, L = self.nf, self.rc, self.n_layers
        s = {"E.layers.0.weight": (nf // 2, 3, 4, 4), "E.layers.0.bias": (nf // 2,),
             "E.layers.2.weight": (nf, nf // 2, 4, 4), "E.layers.2.bias": (nf,),
             "E.layers.4.weight": (nf, nf, 3, 3), "E.layers.4.bias": (nf,)}
        for i in range(L):
            p = f"E.layers.{5 + i}.block."
            s[p + "1.weight"], s[p + "1.bias"] = (rc, nf, 3, 3), (rc,)
            s[p + "3.weight"], s[p + "3.bias"] = (nf, rc, 1, 1), (nf,)
        s["G.layers.0.weight"], s["G.layers.0.bias"] = (nf, nf, 3, 3), (nf,)
        for i in range(L):
            p = f"G.layers.{1 + i}.block."
            s[p + "1.weight"], s[p + "1.bias"] = (rc, nf, 3, 3), (rc,)
            s[p + "3.weight"], s[p + "3.bias"] = (nf, rc, 1, 1), (nf,)
        k = 1 + L
        s[f"G.layers.{k + 1}.weight"], s[f"G.layers.{k + 1}.bias"] = (nf, nf // 2, 4, 4), (nf // 2,)
        s[f"G.layers.{k + 3}.weight"], s[f"G.layers.{k + 3}.bias"] = (nf // 2, 3, 4, 4), (3,)
        return s


class VQVAEEngine:
    def __init__(self, spec: VQVAESpec, device="cuda"):
        _lib.require_device()
        self.spec, self.device, self.lib = spec, torch.device(device), _lib.load()
        self.store = ParamStore(spec.param_shapes(), self.device)
        s = spec
        f32, bf16 = torch.float32, torch.bfloat16
        z = lambda shape, dt=f32: torch.zeros(shape, dtype=dt, device=self.device)  # noqa: E731
        self.codebook = z((s.num, s.K, s.D))
        self.running_size = z((s.num, s.K))
        self.running_sum = z((s.num, s.K, s.D))
        self.cb_grad = z((s.num, s.K, s.D))  # codebook gradient (CODEBOOK.EMA False only)
        L, nf, rc = s.n_layers, s.nf, s.rc
        self.kG = 1 + L
        # packed weights: name -> (fwd [co][tap*ci] bf16, dgrad [ci][tap*co] bf16, fwd-layout fp32 grad)
        self.pk: Dict[str, dict] = {}

        def conv_entry(name, co, ci, taps, dgrad=True):
            self.pk[name] = dict(kind="conv", co=co, ci=ci, T=taps, fwd=z((co, taps * ci), bf16),
                                 dg=z((ci, taps * co), bf16) if dgrad else None, grad=z((co, taps * ci)))
        self.w1p = z((nf // 2, 64), bf16)
        self.dw1p = z((nf // 2, 64))
        conv_entry("E.layers.2.weight", nf, nf // 2, 16, dgrad=False)
        self.e2_dg = z((4, nf // 2, 4 * nf), bf16)  # per input parity: [ci][4 taps * co]
        conv_entry("E.layers.4.weight", nf, nf, 9)
        for i in range(L):
            conv_entry(f"E.layers.{5 + i}.block.1.weight", rc, nf, 9)
        conv_entry("G.layers.0.weight", nf, nf, 9)
        for i in range(L):
            conv_entry(f"G.layers.{1 + i}.block.1.weight", rc, nf, 9)
        # ConvTranspose2d(nf -> nf/2): weight [ci=nf][co=nf/2][4][4]
        self.ct1_fwd = z((4, nf // 2, 4 * nf), bf16)      # per output parity: [co][4 taps * ci]
        self.ct1_grad = z((4, nf // 2, 4 * nf))
        self.ct1_dg = z((nf, 16 * (nf // 2)), bf16)       # as Conv2d(k4,s2,p1): [ci][16 taps * co]
        self.dw_out = z((nf // 2, 64))                    # output ConvT weight gradient scratch [c][48 (+16)]
        # output ConvTranspose2d(nf/2 -> 3) weight [c][co][kh][kw] as GEMM operands (rows beyond 48 stay zero)
        self.w_out_fwd = z((64, nf // 2), bf16)           # [(kh*4+kw)*3 + co][c]
        self.w_out_dg = z((nf // 2, 64), bf16)            # [c][co*16 + kh*4 + kw]
        # the packed fp32 weight gradients live in ONE flat buffer, so that zeroing them is one launch per step, not ten
        pg = [("dw1p", None), ("ct1_grad", None), ("dw_out", None)] + [("pk", k) for k in self.pk]
        get = lambda a, k: getattr(self, a) if k is None else self.pk[k]["grad"]  # noqa: E731
        sizes = [(get(a, k).numel() + 3) // 4 * 4 for a, k in pg]
        self._pg_flat = z((sum(sizes),))
        off = 0
        for (a, k), n in zip(pg, sizes):
            t = get(a, k)
            view = self._pg_flat[off:off + t.numel()].view(t.shape)
            if k is None:
                setattr(self, a, view)
            else:
                self.pk[k]["grad"] = view
            off += n
        self._ws = {}
        self.shadows_fresh = False
        self.opt = None
        self._rec, self._pbatch = None, {}
        self._pp = None             # [hi | hi | lo] split packs of the encoder weights (high-precision encode)
        self._pp_fresh = False

    # ------------------------------------------------------------------ parameters
    def load_state_dict(self, netE=None, netG=None, codebook=None, running_size=None, running_sum=None):
        if netE:
            self.store.load({"E." + k: v for k, v in netE.items()})
        if netG:
            self.store.load({"G." + k: v for k, v in netG.items()})
        if codebook is not None:
            self.codebook.copy_(codebook)
            # GPU semantics of the reference: running_sum starts as a COPY of the codebook
            self.running_sum.copy_(codebook if running_sum is None else running_sum)
            if running_size is not None:
                self.running_size.copy_(running_size)
        self.shadows_fresh = False

    def _permute4(self, src, dst, bf16, acc, dims, istr, ostr):
        if self._rec is not None:  # inside _batched(): recorded, launched together
            self._rec.add(src, dst, bf16, acc, dims, istr, ostr)
            return
        check(self.lib.lvt_permute4(_vp(src), _vp(dst), int(bf16), int(acc), (ctypes.c_int * 4)(*dims),
                                    (ctypes.c_longlong * 4)(*istr), (ctypes.c_longlong * 4)(*ostr),
                                    stream_ptr()), "lvt_permute4")

    def _batched(self, key, fn):
        """The lvt_permute4 calls of fn() as ONE launch (lvt_permute4_batch): recorded on first use (pointers, shapes
        and strides are fixed), replayed afterwards -- the per-step weight packs and gradient folds were 74 launches."""
        b = self._pbatch.get(key)
        if b is None:
            b = self._pbatch[key] = _lib.PermuteBatch()
            self._rec = b
            try:
                fn()
            finally:
                self._rec = None
        b.run(self.device)

    def refresh_shadows(self, cast=True):
        self._pp_fresh = False
        st = self.store
        if cast:
            check(self.lib.lvt_cast_bf16(ptr(st.master), ptr(st.shadow), st.numel, stream_ptr()), "lvt_cast_bf16")
        self._batched(("refresh",), self._refresh_jobs)
        self.shadows_fresh = True

    def _refresh_jobs(self):
        s, st = self.spec, self.store
        nf = s.nf
        # conv1 [128][3][16] -> [128][(tap, c)] padded to 64 columns
        self._permute4(st.pf("E.layers.0.weight"), self.w1p, True, False, (nf // 2, 16, 3, 1), (48, 1, 16, 0),
                       (64, 3, 1, 0))
        for name, e in self.pk.items():
            co, ci, T = e["co"], e["ci"], e["T"]
            self._permute4(st.pf(name), e["fwd"], True, False, (co, T, ci, 1), (ci * T, 1, T, 0), (T * ci, ci, 1, 0))
            if e["dg"] is not None:
                self._permute4(st.pf(name), e["dg"], True, False, (ci, T, co, 1), (T, 1, ci * T, 0), (T * co, co, 1, 0))
        # conv2 data-gradient packs per input parity: [ci][j][co] = W[co][ci][tap_j]
        co, ci = nf, nf // 2
        for ph in range(2):
            for pw in range(2):
                ids, _ = phase_taps(ph, pw)
                for j, t in enumerate(ids):
                    self._permute4(st.pf("E.layers.2.weight") + 4 * t, self.e2_dg[ph * 2 + pw].data_ptr() + 2 * j * co,
                                   True, False, (ci, co, 1, 1), (16, ci * 16, 0, 0), (4 * co, 1, 0, 0))
        # ConvTranspose2d weight [ci=nf][co=nf/2][16]
        wct = f"G.layers.{self.kG + 1}.weight"
        ci, co = nf, nf // 2
        for ph in range(2):
            for pw in range(2):
                ids, _ = phase_taps(ph, pw)
                for j, t in enumerate(ids):
                    self._permute4(st.pf(wct) + 4 * t, self.ct1_fwd[ph * 2 + pw].data_ptr() + 2 * j * ci, True, False,
                                   (co, ci, 1, 1), (16, co * 16, 0, 0), (4 * ci, 1, 0, 0))
        self._permute4(st.pf(wct), self.ct1_dg, True, False, (ci, 16, co, 1), (co * 16, 1, 16, 0), (16 * co, co, 1, 0))
        wo = st.pf(f"G.layers.{self.kG + 3}.weight")      # [c][3][16]
        c2 = nf // 2
        self._permute4(wo, self.w_out_fwd, True, False, (16, 3, c2, 1), (1, 16, 48, 0), (3 * c2, c2, 1, 0))
        self._permute4(wo, self.w_out_dg, True, False, (c2, 48, 1, 1), (48, 1, 0, 0), (64, 1, 0, 0))

    def _fold_packed_grads(self):
        """packed-layout weight gradients -> reference layout in the flat gradient (+=)."""
        self._batched(("fold",), self._fold_jobs)

    def _fold_jobs(self):
        s, st = self.spec, self.store
        nf = s.nf
        self._permute4(self.dw1p, st.gf("E.layers.0.weight"), False, True, (nf // 2, 16, 3, 1), (64, 3, 1, 0),
                       (48, 1, 16, 0))
        for name, e in self.pk.items():
            co, ci, T = e["co"], e["ci"], e["T"]
            # (fastest dim = the taps, over which the destination of this read-modify-write is contiguous)
            self._permute4(e["grad"], st.gf(name), False, True, (1, co, ci, T), (0, T * ci, 1, ci), (0, ci * T, T, 1))
        wct = f"G.layers.{self.kG + 1}.weight"
        ci, co = nf, nf // 2
        for ph in range(2):
            for pw in range(2):
                ids, _ = phase_taps(ph, pw)
                for j, t in enumerate(ids):
                    self._permute4(self.ct1_grad[ph * 2 + pw].data_ptr() + 4 * j * ci, st.gf(wct) + 4 * t, False, True,
                                   (co, ci, 1, 1), (4 * ci, 1, 0, 0), (16, co * 16, 0, 0))
        # output ConvT: scratch [c][64] (48 valid) -> [c][3][4][4]
        self._permute4(self.dw_out, st.gf(f"G.layers.{self.kG + 3}.weight"), False, True, (nf // 2, 48, 1, 1),
                       (64, 1, 0, 0), (48, 1, 0, 0))

    def _zero_packed_grads(self):
        self._pg_flat.zero_()

    # ------------------------------------------------------------------ workspace
    def workspace(self, n, train=True):
        key = (n, train)
        if key in self._ws:
            return self._ws[key]
        s = self.spec
        f32, bf16 = torch.float32, torch.bfloat16
        e = lambda shape, dt=bf16: torch.empty(shape, dtype=dt, device=self.device)  # noqa: E731
        M, nf, rc, L = n * 256, s.nf, s.rc, s.n_layers
        w = type("VQVAEWorkspace", (), {})()
        w.n, w.M = n, M
        w.x = torch.zeros((n, 3, 64, 64), dtype=f32, device=self.device)
        w.A1 = e((4 * M, 64))
        w.act1 = e((4 * M, nf // 2))
        w.act2 = e((M, nf))
        w.er = [e((M, nf)) for _ in range(L + 1)]       # r_0 .. r_L inputs of the encoder blocks (ReLU'd)
        w.eh = [e((M, rc)) for _ in range(L)]
        w.z_e = e((M, nf), f32)
        w.idx = torch.empty((n, s.num, 16, 16), dtype=torch.int64, device=self.device)
        w.zq_st = e((M, nf))
        w.zq_bar = e((M, nf), f32)
        w.counts = torch.zeros((s.num, s.K), dtype=f32, device=self.device)
        w.sums = torch.zeros((s.num, s.K, s.D), dtype=f32, device=self.device)
        w.gr = [e((M, nf)) for _ in range(L + 1)]       # decoder block inputs; gr[L] = ReLU'd decoder trunk output
        w.gh = [e((M, rc)) for _ in range(L)]
        w.act32 = e((4 * M, nf // 2))
        w.Y = torch.empty((4 * M, 64), dtype=f32, device=self.device)  # output ConvT partial sums per input pixel
        w.x_tilde = e((n, 3, 64, 64), f32)
        w.recon = e((n, 3, 64, 64), f32)
        # [reconstruction, commitment] (+ [2] = mse(z_q, sg[z_e]), the `loss_dict` entry of vqvae.py:84-85, without EMA)
        w.loss = torch.zeros((2 if s.ema else 3,), dtype=f32, device=self.device)
        w.loss_scratch = torch.zeros((1,), dtype=f32, device=self.device)
        if train:
            w.dpre = e((n, 3, 64, 64), f32)
            w.dact32 = e((4 * M, nf // 2))
            w.G = e((4 * M, 64))
            w.d_a = e((M, nf))
            w.d_b = e((M, nf))
            w.dh = e((M, rc))
            w.dz_st = e((M, nf), f32)
            w.dact1 = e((4 * M, nf // 2))
        self._ws[key] = w
        return w

    # ------------------------------------------------------------------ helpers
    def _conv(self, x, C, n, wpk, K, N, out, bias, taps, relu=True, P=1, s_phase=0, aux=None, flags=0, f32=False):
        """out[M, N] = epi(conv(x) @ wpk^T + bias)."""
        M = n * 256
        gemm(M, N, K, Operand(_vp(x).value, C), Operand(_vp(wpk).value, K), Operand(_vp(out).value, N),
             out_f32=out if f32 else None, out_bf16=None if f32 else out, bias=bias, aux=aux,
             flags=flags | (ops.GEMM_RELU if relu else 0), conv=ConvSpec("a", C, 16, 16, n, taps, P=P, s_phase=s_phase))

    def _colsum(self, x, out, rows, N):
        check(self.lib.lvt_colsum_bf16(_vp(x), _vp(out), rows, N, N, stream_ptr()), "lvt_colsum_bf16")

    @staticmethod
    def _splits(m, n, k):
        """split-K factor of a weight-gradient GEMM: chosen by the library (fills the SMs / SM pairs once with
        its tile shape, at least 4 k-blocks of 64 per split; lvt_gemm_bf16, splits < 0)."""
        return -1

    def _wgrad_conv(self, dy, co, x, C, n, taps, grad, P=1, s_phase=0):
        """grad[co][tap*C + c] += sum_m dy[m, co] * x[m + tap, c]."""
        M = n * 256
        N = len(taps) * C
        gemm(co, N, M, Operand(_vp(dy).value, co, mn_major=True), Operand(_vp(x).value, C, mn_major=True),
             Operand(_vp(grad).value, N), out_f32=grad, splits=self._splits(co, N, M), flags=ops.GEMM_ATOMIC,
             conv=ConvSpec("b", C, 16, 16, n, taps, P=P, s_phase=s_phase))

    def _wgrad_plain(self, dy, co, x, ci, rows, grad, ld_grad=None):
        gemm(co, ci, rows, Operand(_vp(dy).value, co, mn_major=True), Operand(_vp(x).value, ci, mn_major=True),
             Operand(_vp(grad).value, ld_grad or ci), out_f32=grad, splits=self._splits(co, ci, rows),
             flags=ops.GEMM_ATOMIC)

    def _block_fwd(self, pre, i, r, h, out, n, last_f32=False, relu_out=True):
        """ResBlock (resencoder.py:10-21): out = r + conv1x1(relu(conv3x3(r))), r already ReLU'd."""
        st, s = self.store, self.spec
        wA, wB = f"{pre}.block.1.weight", f"{pre}.block.3.weight"
        self._conv(r, s.nf, n, self.pk[wA]["fwd"], 9 * s.nf, s.rc, h, st.pf(f"{pre}.block.1.bias"), TAPS3)
        M = n * 256
        if last_f32 and not relu_out:
            # fp32 output (z_e) with a bf16 skip: the TMA-store epilogue has no such pair (the staged generic epilogue took
            # 145 us for it at 512 frames), so the 1x1 convolution stores fp32 through TMA and one pass adds the skip
            gemm(M, s.nf, s.rc, Operand(h.data_ptr(), s.rc), Operand(st.pb(wB), s.rc), Operand(out.data_ptr(), s.nf),
                 out_f32=out, bias=st.pf(f"{pre}.block.3.bias"))
            check(self.lib.lvt_add_bf16_to_f32(ptr(out), ptr(r), M * s.nf, stream_ptr()), "lvt_add_bf16_to_f32")
            return
        gemm(M, s.nf, s.rc, Operand(h.data_ptr(), s.rc), Operand(st.pb(wB), s.rc), Operand(out.data_ptr(), s.nf),
             out_f32=out if last_f32 else None, out_bf16=None if last_f32 else out, bias=st.pf(f"{pre}.block.3.bias"),
             aux=r, flags=ops.GEMM_AUX_ADD | (ops.GEMM_RELU if relu_out else 0))

    def _block_bwd(self, pre, r, h, d_out, d_prev, w, n):
        """d_out: gradient wrt the block output (bf16); writes d_prev = gradient wrt the pre-ReLU input."""
        st, s = self.store, self.spec
        M, nf, rc = n * 256, s.nf, s.rc
        wA, wB = f"{pre}.block.1.weight", f"{pre}.block.3.weight"
        self._colsum(d_out, st.gf(f"{pre}.block.3.bias"), M, nf)
        self._wgrad_plain(d_out, nf, h, rc, M, st.gf(wB))
        gemm(M, rc, nf, Operand(d_out.data_ptr(), nf), Operand(st.pb(wB), rc, mn_major=True), Operand(w.dh.data_ptr(), rc),
             out_bf16=w.dh, aux=h, flags=ops.GEMM_MASK)
        self._colsum(w.dh, st.gf(f"{pre}.block.1.bias"), M, rc)
        self._wgrad_conv(w.dh, rc, r, nf, n, TAPS3, self.pk[wA]["grad"])
        # dr = conv-dgrad(dh) + d_out (skip), then ReLU mask of r
        self._conv(w.dh, rc, n, self.pk[wA]["dg"], 9 * rc, nf, d_prev, None, TAPS3_T, relu=False, aux=d_out,
                   flags=ops.GEMM_AUX_ADD)
        check(self.lib.lvt_relu_bwd_add(ptr(d_prev), None, ptr(r), ptr(d_prev), M * nf, stream_ptr()), "lvt_relu_bwd_add")

    # ------------------------------------------------------------------ forward pieces
    def encode(self, w):
        """ResEncoder.forward + DVQ indices (vqvae.py:93-101).  w.x -> w.z_e, w.idx (+ w.zq_st)."""
        s, st = self.spec, self.store
        if not self.shadows_fresh:
            self.refresh_shadows()
        n, M, nf, L = w.n, w.M, s.nf, s.n_layers
        check(self.lib.lvt_vqvae_in_im2col(ptr(w.x), ptr(w.A1), n, s.mean, s.std, stream_ptr()), "lvt_vqvae_in_im2col")
        gemm(4 * M, nf // 2, 64, Operand(w.A1.data_ptr(), 64), Operand(self.w1p.data_ptr(), 64),
             Operand(w.act1.data_ptr(), nf // 2), out_bf16=w.act1, bias=st.pf("E.layers.0.bias"), flags=ops.GEMM_RELU)
        self._conv(w.act1, nf // 2, n, self.pk["E.layers.2.weight"]["fwd"], 16 * (nf // 2), nf, w.act2,
                   st.pf("E.layers.2.bias"), TAPS_K4S2, P=4, s_phase=M * (nf // 2))
        if L == 0:
            self._conv(w.act2, nf, n, self.pk["E.layers.4.weight"]["fwd"], 9 * nf, nf, w.z_e, st.pf("E.layers.4.bias"),
                       TAPS3, relu=False, f32=True)
        else:
            self._conv(w.act2, nf, n, self.pk["E.layers.4.weight"]["fwd"], 9 * nf, nf, w.er[0], st.pf("E.layers.4.bias"),
                       TAPS3)
            for i in range(L):
                last = i == L - 1
                self._block_fwd(f"E.layers.{5 + i}", i, w.er[i], w.eh[i], w.z_e if last else w.er[i + 1], n,
                                last_f32=last, relu_out=not last)

    # ------------------------------------------------------------------ high-precision encoder (inference)
    def _split3(self, src, add, out, out_f32, rows, C, relu, weight):
        check(self.lib.lvt_split3_bf16(_vp(src), _vp(add) if add is not None else None, _vp(out),
                                       _vp(out_f32) if out_f32 is not None else None, rows, C, int(relu), int(weight),
                                       stream_ptr()), "lvt_split3_bf16")

    def _refresh_precise(self):
        """Encoder weights as [hi | hi | lo] bf16 triples per (output channel, tap), built from the fp32 master."""
        s, st = self.spec, self.store
        nf, rc, L = s.nf, s.rc, s.n_layers
        f32, bf16 = torch.float32, torch.bfloat16
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=self.device)  # noqa: E731
        if self._pp is None:
            pp = {"w1": z((nf // 2, 192), bf16), "scratch": z((nf * 16 * nf,), f32)}
            for name in ["E.layers.2.weight", "E.layers.4.weight"] + [f"E.layers.{5 + i}.block.1.weight" for i in range(L)]:
                e = self.pk[name]
                pp[name] = z((e["co"], e["T"] * 3 * e["ci"]), bf16)
            for i in range(L):
                pp[f"E.layers.{5 + i}.block.3.weight"] = z((nf, 3 * rc), bf16)
            self._pp = pp
        pp, sc = self._pp, self._pp["scratch"]
        # conv1: [128][3][16] -> [128][(tap, c)] padded to 64 columns, then split
        sc[:(nf // 2) * 64].zero_()
        self._permute4(st.pf("E.layers.0.weight"), sc, False, False, (nf // 2, 16, 3, 1), (48, 1, 16, 0), (64, 3, 1, 0))
        self._split3(sc, None, pp["w1"], None, nf // 2, 64, False, True)
        for name, dst in pp.items():
            if name in ("w1", "scratch"):
                continue
            if name.endswith("block.3.weight"):      # 1x1 conv: the master layout [nf][rc] is already K-major
                self._split3(st.pf(name), None, dst, None, nf, rc, False, True)
                continue
            e = self.pk[name]
            co, ci, T = e["co"], e["ci"], e["T"]
            self._permute4(st.pf(name), sc, False, False, (co, T, ci, 1), (ci * T, 1, T, 0), (T * ci, ci, 1, 0))
            self._split3(sc, None, dst, None, co * T, ci, False, True)
        self._pp_fresh = True

    def _precise_ws(self, w):
        if getattr(w, "pz", None) is None:
            s = self.spec
            M, nf, rc = w.M, s.nf, s.rc
            f32, bf16 = torch.float32, torch.bfloat16
            e = lambda shape, dt: torch.empty(shape, dtype=dt, device=self.device)  # noqa: E731
            w.pz = dict(A1=e((4 * M, 192), bf16), t1=e((4 * M, nf // 2), f32), act1=e((4 * M, 3 * (nf // 2)), bf16),
                        t2=e((M, nf), f32), act2=e((M, 3 * nf), bf16), r32=e((M, nf), f32), r=e((M, 3 * nf), bf16),
                        th=e((M, rc), f32), h=e((M, 3 * rc), bf16))
        return w.pz

    def encode_precise(self, w):
        """ResEncoder.forward with every product carried as a 3-term bf16 split (lvt_split3_bf16): fp32 accumulation
        of a_hi w_hi + a_lo w_hi + a_hi w_lo, i.e. ~2^-17 relative per product instead of bf16's 2^-9.  Used for
        encode / inference, where the code indices are the product (CodesExtractor -> VT training data); the
        training step keeps the plain bf16 convolutions.  w.x -> w.z_e."""
        s, st = self.spec, self.store
        if not self.shadows_fresh:
            self.refresh_shadows()
        if not self._pp_fresh:
            self._refresh_precise()
        n, M, nf, rc, L = w.n, w.M, s.nf, s.rc, s.n_layers
        pz, pp = self._precise_ws(w), self._pp
        c1 = nf // 2
        check(self.lib.lvt_vqvae_in_im2col_split(ptr(w.x), ptr(pz["A1"]), n, s.mean, s.std, stream_ptr()),
              "lvt_vqvae_in_im2col_split")
        gemm(4 * M, c1, 192, Operand(pz["A1"].data_ptr(), 192), Operand(pp["w1"].data_ptr(), 192),
             Operand(pz["t1"].data_ptr(), c1), out_f32=pz["t1"], bias=st.pf("E.layers.0.bias"))
        self._split3(pz["t1"], None, pz["act1"], None, 4 * M, c1, True, False)
        self._conv(pz["act1"], 3 * c1, n, pp["E.layers.2.weight"], 16 * 3 * c1, nf, pz["t2"], st.pf("E.layers.2.bias"),
                   TAPS_K4S2, relu=False, P=4, s_phase=M * 3 * c1, f32=True)
        self._split3(pz["t2"], None, pz["act2"], None, M, nf, True, False)
        out0 = w.z_e if L == 0 else pz["t2"]
        self._conv(pz["act2"], 3 * nf, n, pp["E.layers.4.weight"], 9 * 3 * nf, nf, out0, st.pf("E.layers.4.bias"),
                   TAPS3, relu=False, f32=True)
        if L == 0:
            return
        self._split3(pz["t2"], None, pz["r"], pz["r32"], M, nf, True, False)        # r_0 = relu(conv3(...))
        for i in range(L):
            pre = f"E.layers.{5 + i}"
            last = i == L - 1
            self._conv(pz["r"], 3 * nf, n, pp[f"{pre}.block.1.weight"], 9 * 3 * nf, rc, pz["th"],
                       st.pf(f"{pre}.block.1.bias"), TAPS3, relu=False, f32=True)
            self._split3(pz["th"], None, pz["h"], None, M, rc, True, False)
            out = w.z_e if last else pz["t2"]
            # out = r + conv1x1(h) + bias (the skip adds the ReLU'd block input, resencoder.py:10-21)
            gemm(M, nf, 3 * rc, Operand(pz["h"].data_ptr(), 3 * rc), Operand(pp[f"{pre}.block.3.weight"].data_ptr(), 3 * rc),
                 Operand(out.data_ptr(), nf), out_f32=out, bias=st.pf(f"{pre}.block.3.bias"), res=pz["r32"])
            if not last:
                self._split3(pz["t2"], None, pz["r"], pz["r32"], M, nf, True, False)

    def quantize(self, w, train):
        s = self.spec
        n = w.n
        if train:
            w.counts.zero_()
            w.sums.zero_()
        check(self.lib.lvt_vq_argmin_nhwc(ptr(w.z_e), ptr(self.codebook), ptr(w.idx), None, ptr(w.zq_st),
                                          ptr(w.counts) if train else None, ptr(w.sums) if train else None, n, s.num,
                                          s.K, s.D, 256, stream_ptr()), "lvt_vq_argmin_nhwc")

    def ema_update(self, w):
        """vq_embedding.py:48-64 (after the cross-rank sum of counts / sums) + z_q_bar from the NEW codebook."""
        s = self.spec
        check(self.lib.lvt_vq_ema_update(ptr(self.codebook), ptr(self.running_size), ptr(self.running_sum),
                                         ptr(w.counts), ptr(w.sums), s.num, s.K, s.D, float(s.ema_decay),
                                         float(s.ema_eps), stream_ptr()), "lvt_vq_ema_update")
        check(self.lib.lvt_vq_gather_nhwc(ptr(w.idx), ptr(self.codebook), ptr(w.zq_bar), None, w.n, s.num, s.K, s.D,
                                          256, stream_ptr()), "lvt_vq_gather_nhwc")

    def decode(self, w, zq=None):
        """ResDecoder.forward (resdecoder.py:48-57,66-69): zq bf16 [M, 256] channels-last -> w.x_tilde."""
        s, st = self.spec, self.store
        if not self.shadows_fresh:
            self.refresh_shadows()
        n, M, nf, L = w.n, w.M, s.nf, s.n_layers
        zq = w.zq_st if zq is None else zq
        self._conv(zq, nf, n, self.pk["G.layers.0.weight"]["fwd"], 9 * nf, nf, w.gr[0], st.pf("G.layers.0.bias"), TAPS3)
        for i in range(L):
            self._block_fwd(f"G.layers.{1 + i}", i, w.gr[i], w.gh[i], w.gr[i + 1], n)
        k = self.kG
        for ph in range(2):
            for pw in range(2):
                _, shifts = phase_taps(ph, pw)
                p = ph * 2 + pw
                self._conv(w.gr[L], nf, n, self.ct1_fwd[p], 4 * nf, nf // 2, w.act32[p * M:(p + 1) * M],
                           st.pf(f"G.layers.{k + 1}.bias"), shifts)
        # output ConvTranspose2d(nf/2 -> 3) + tanh: channel contraction on the tensor cores, then the 2x2 tap gather
        c2 = nf // 2
        gemm(4 * M, 64, c2, Operand(w.act32.data_ptr(), c2), Operand(self.w_out_fwd.data_ptr(), c2),
             Operand(w.Y.data_ptr(), 64), out_f32=w.Y)
        check(self.lib.lvt_vqvae_out_col2im_tanh(ptr(w.Y), _vp(st.pf(f"G.layers.{k + 3}.bias")), ptr(w.x_tilde), n,
                                                 stream_ptr()), "lvt_vqvae_out_col2im_tanh")

    def inference(self, w, precise=False):
        """AutoEncoderModel.forward(mode='inference') (ae.py:120-147): x in [0,1] -> recon in [0,1], latent.
        precise: the encoder runs its convolutions as 3-term bf16 splits (encode_precise)."""
        s = self.spec
        if precise:
            self.encode_precise(w)
        else:
            self.encode(w)
        self.quantize(w, train=False)
        self.decode(w)
        check(self.lib.lvt_denorm_clamp(ptr(w.x_tilde), ptr(w.recon), w.x_tilde.numel(), s.mean, s.std, 0.0, 1.0,
                                        stream_ptr()), "lvt_denorm_clamp")
        return w.recon, w.idx

    def decode_indices(self, w, idx):
        """VQVAEModel.decode (vqvae.py:103-106): codes [n, num, 16, 16] -> x_tilde."""
        s = self.spec
        check(self.lib.lvt_vq_gather_nhwc(ptr(idx), ptr(self.codebook), None, ptr(w.zq_st), w.n, s.num, s.K, s.D, 256,
                                          stream_ptr()), "lvt_vq_gather_nhwc")
        self.decode(w)
        return w.x_tilde

    # ------------------------------------------------------------------ training
    def forward_train(self, w, allreduce=None):
        """compute_supervised_loss (vqvae.py:66-91): losses in w.loss = [reconstruction, commitment(after bwd)]."""
        self._forward_train_a(w)
        if allreduce is not None and self.spec.ema:  # without EMA the statistics only feed this rank's codebook gradient
            allreduce(w.counts)
            allreduce(w.sums)
        self._forward_train_b(w)

    def _forward_train_a(self, w):
        """encoder + codebook search with the OLD codebook; per-code counts / sums of this rank"""
        self.encode(w)
        self.quantize(w, train=True)

    def _forward_train_b(self, w):
        """EMA codebook update (after the cross-rank sum), decoder, losses"""
        s = self.spec
        if s.ema:
            self.ema_update(w)
        else:  # z_q_bar = index_select from the (trained, unchanged here) codebook (vq_embedding.py:61-64)
            check(self.lib.lvt_vq_gather_nhwc(ptr(w.idx), ptr(self.codebook), ptr(w.zq_bar), None, w.n, s.num, s.K,
                                              s.D, 256, stream_ptr()), "lvt_vq_gather_nhwc")
        self.decode(w)
        w.loss.zero_()
        k = self.kG
        check(self.lib.lvt_vqvae_recon_loss(ptr(w.x_tilde), ptr(w.x), ptr(w.dpre), ptr(w.loss),
                                            _vp(self.store.gf(f"G.layers.{k + 3}.bias")), w.n, s.mean, s.std,
                                            s.pixel_lambda, stream_ptr()), "lvt_vqvae_recon_loss")
        check(self.lib.lvt_vqvae_commit_loss(ptr(w.z_e), ptr(w.zq_bar), None, None, _vp(w.loss.data_ptr() + 4),
                                             w.M * s.nf, s.beta, stream_ptr()), "lvt_vqvae_commit_loss")
        if not s.ema:  # the vector-quantisation objective mse(z_q, sg[z_e]): the same mean square, without beta
            check(self.lib.lvt_vqvae_commit_loss(ptr(w.z_e), ptr(w.zq_bar), None, None, _vp(w.loss.data_ptr() + 8),
                                                 w.M * s.nf, 1.0, stream_ptr()), "lvt_vqvae_commit_loss")

    def backward(self, w):
        s, st = self.spec, self.store
        n, M, nf, rc, L = w.n, w.M, s.nf, s.rc, s.n_layers
        k = self.kG
        self._zero_packed_grads()
        # ---- output ConvTranspose2d + tanh
        check(self.lib.lvt_vqvae_out_convt_g(ptr(w.dpre), ptr(w.G), n, stream_ptr()), "lvt_vqvae_out_convt_g")
        gemm(4 * M, nf // 2, 64, Operand(w.G.data_ptr(), 64), Operand(self.w_out_dg.data_ptr(), 64),
             Operand(w.dact32.data_ptr(), nf // 2), out_bf16=w.dact32, aux=w.act32, flags=ops.GEMM_MASK)
        self._wgrad_plain(w.act32, nf // 2, w.G, 64, 4 * M, self.dw_out)
        # ---- ConvTranspose2d(nf -> nf/2)
        self._colsum(w.dact32, st.gf(f"G.layers.{k + 1}.bias"), 4 * M, nf // 2)
        for ph in range(2):
            for pw in range(2):
                _, shifts = phase_taps(ph, pw)
                p = ph * 2 + pw
                self._wgrad_conv(w.dact32[p * M:(p + 1) * M], nf // 2, w.gr[L], nf, n, shifts, self.ct1_grad[p])
        d_out, d_prev = w.d_a, w.d_b
        self._conv(w.dact32, nf // 2, n, self.ct1_dg, 16 * (nf // 2), nf, d_out, None, TAPS_K4S2, relu=False, P=4,
                   s_phase=M * (nf // 2), aux=w.gr[L], flags=ops.GEMM_MASK)
        # ---- decoder blocks
        for i in reversed(range(L)):
            self._block_bwd(f"G.layers.{1 + i}", w.gr[i], w.gh[i], d_out, d_prev, w, n)
            d_out, d_prev = d_prev, d_out
        if L == 0:
            pass  # d_out already masked by gr[0] > 0
        # ---- decoder conv0: gr[0] = relu(conv(zq_st))
        self._colsum(d_out, st.gf("G.layers.0.bias"), M, nf)
        self._wgrad_conv(d_out, nf, w.zq_st, nf, n, TAPS3, self.pk["G.layers.0.weight"]["grad"])
        self._conv(d_out, nf, n, self.pk["G.layers.0.weight"]["dg"], 9 * nf, nf, w.dz_st, None, TAPS3_T, relu=False,
                   f32=True)
        # ---- straight-through + commitment -> gradient wrt z_e
        check(self.lib.lvt_vqvae_commit_loss(ptr(w.z_e), ptr(w.zq_bar), ptr(w.dz_st), ptr(d_out),
                                             ptr(w.loss_scratch), M * nf, s.beta, stream_ptr()),
              "lvt_vqvae_commit_loss")
        # ---- encoder blocks (z_e is the un-ReLU'd output of the last block)
        if L == 0:
            d_conv3 = d_out
        else:
            for i in reversed(range(L)):
                self._block_bwd(f"E.layers.{5 + i}", w.er[i], w.eh[i], d_out, d_prev, w, n)
                d_out, d_prev = d_prev, d_out
            d_conv3 = d_out
        # ---- conv3 (256->256 k3): er[0] = relu(conv3(act2))
        self._colsum(d_conv3, st.gf("E.layers.4.bias"), M, nf)
        self._wgrad_conv(d_conv3, nf, w.act2, nf, n, TAPS3, self.pk["E.layers.4.weight"]["grad"])
        d_act2 = d_prev
        self._conv(d_conv3, nf, n, self.pk["E.layers.4.weight"]["dg"], 9 * nf, nf, d_act2, None, TAPS3_T, relu=False,
                   aux=w.act2, flags=ops.GEMM_MASK)
        # ---- conv2 (128->256 k4 s2) over the phase-major act1
        self._colsum(d_act2, st.gf("E.layers.2.bias"), M, nf)
        self._wgrad_conv(d_act2, nf, w.act1, nf // 2, n, TAPS_K4S2, self.pk["E.layers.2.weight"]["grad"], P=4,
                         s_phase=M * (nf // 2))
        for ph in range(2):
            for pw in range(2):
                _, shifts = phase_taps(ph, pw)
                p = ph * 2 + pw
                self._conv(d_act2, nf, n, self.e2_dg[p], 4 * nf, nf // 2, w.dact1[p * M:(p + 1) * M], None, shifts,
                           relu=False, aux=w.act1[p * M:(p + 1) * M], flags=ops.GEMM_MASK)
        # ---- conv1 (3->128 k4 s2): weight gradient only
        self._colsum(w.dact1, st.gf("E.layers.0.bias"), 4 * M, nf // 2)
        self._wgrad_plain(w.dact1, nf // 2, w.A1, 64, 4 * M, self.dw1p)
        self._fold_packed_grads()
        if not s.ema:  # d mse(z_q, sg[z_e]) / d codebook from this rank's per-code counts / sums
            check(self.lib.lvt_vq_codebook_grad(ptr(w.counts), ptr(w.sums), ptr(self.codebook), ptr(self.cb_grad),
                                                2.0 / (M * nf), s.num * s.K, s.D, stream_ptr()), "lvt_vq_codebook_grad")

    def init_optimizer(self, lr=3e-4, betas=(0.9, 0.9), eps=1e-8):
        """torch.optim.Adam(lr=LR_G, betas=(BETA1_G, BETA2_G)) (config/defaults.py:105-114; solver/build.py:62-66);
        netE and netG share hyper-parameters and step count, so one flat update covers both."""
        self.opt = dict(lr=lr, betas=betas, eps=eps, step=0)
        self.opt_m = torch.zeros_like(self.store.master)
        self.opt_v = torch.zeros_like(self.store.master)
        if not self.spec.ema:  # optimizer_c of vqvae.py:108-116: same builder, same hyper-parameters, stepped together
            self.cb_m = torch.zeros_like(self.codebook)
            self.cb_v = torch.zeros_like(self.codebook)

    def optimizer_step(self, grad_scale=1.0):
        o, st = self.opt, self.store
        o["step"] += 1
        check(self.lib.lvt_adam_step(ptr(st.master), ptr(st.grad), ptr(self.opt_m), ptr(self.opt_v), ptr(st.shadow),
                                     st.numel, o["lr"], o["betas"][0], o["betas"][1], o["eps"], o["step"], grad_scale,
                                     stream_ptr()), "lvt_adam_step")
        if not self.spec.ema:
            check(self.lib.lvt_adam_step(ptr(self.codebook), ptr(self.cb_grad), ptr(self.cb_m), ptr(self.cb_v), None,
                                         self.codebook.numel(), o["lr"], o["betas"][0], o["betas"][1], o["eps"],
                                         o["step"], grad_scale, stream_ptr()), "lvt_adam_step")
        self.refresh_shadows(cast=False)

    def train_step(self, w, allreduce=None, world_size=1):
        self.store.grad.zero_()
        self.forward_train(w, allreduce)
        self.backward(w)
        if allreduce is not None:
            allreduce(self.store.grad)
            if not self.spec.ema:
                allreduce(self.cb_grad)
        self.optimizer_step(1.0 / world_size)
        return w.loss


class GraphedVQVAEStep:
    """The VQ-VAE train step (compute_supervised_loss + backward + Adam, trainer.py:79-87) with its ~250 launches
    captured in CUDA graphs: [zero-grad, encoder, codebook search] (all-reduce of the EMA counts / sums when
    world_size > 1, vq_embedding.py:44-59) [EMA update, decoder, losses, backward] (all-reduce of the flat gradient)
    Adam (eager: its bias correction depends on the step count) [weight re-layout]."""

    def __init__(self, engine: "VQVAEEngine", w, world_size=1, allreduce=None):
        if not engine.spec.ema:
            raise _lib.LvtError("GraphedVQVAEStep covers the EMA codebook (every shipped config); "
                                "CODEBOOK.EMA False runs through VQVAEEngine.train_step")
        self.engine, self.w, self.world_size, self.allreduce = engine, w, world_size, allreduce
        self.graphs = None

    def _seg_a(self):
        self.engine.store.grad.zero_()
        self.engine._forward_train_a(self.w)

    def _seg_b(self):
        self.engine._forward_train_b(self.w)
        self.engine.backward(self.w)

    def _adam(self):
        eng = self.engine
        o, st = eng.opt, eng.store
        o["step"] += 1
        check(eng.lib.lvt_adam_step(ptr(st.master), ptr(st.grad), ptr(eng.opt_m), ptr(eng.opt_v), ptr(st.shadow),
                                    st.numel, o["lr"], o["betas"][0], o["betas"][1], o["eps"], o["step"],
                                    1.0 / self.world_size, stream_ptr()), "lvt_adam_step")

    def _refresh(self):
        self.engine.refresh_shadows(cast=False)

    def _eager(self):
        self._seg_a()
        if self.allreduce is not None:
            self.allreduce(self.w.counts)
            self.allreduce(self.w.sums)
        self._seg_b()
        if self.allreduce is not None:
            self.allreduce(self.engine.store.grad)
        self._adam()
        self._refresh()

    def capture(self, warmup=2):
        """The eager warm-up runs real steps; parameters, Adam state and the EMA codebook state are restored
        afterwards so that capturing does not train."""
        eng = self.engine
        snap = [t.clone() for t in (eng.store.master, eng.opt_m, eng.opt_v, eng.codebook, eng.running_size,
                                    eng.running_sum)]
        step0 = eng.opt["step"]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):  # eager warm-up: kernel attributes, TMA maps
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for dst, src in zip((eng.store.master, eng.opt_m, eng.opt_v, eng.codebook, eng.running_size,
                             eng.running_sum), snap):
            dst.copy_(src)
        eng.opt["step"] = step0
        eng.refresh_shadows()
        torch.cuda.synchronize()
        segs = [self._seg_a, self._seg_b] if self.allreduce is not None else [lambda: (self._seg_a(), self._seg_b())]
        self.graphs = []
        for seg in segs + [self._refresh]:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                seg()
            self.graphs.append(g)
        torch.cuda.synchronize()

    def step(self):
        if self.allreduce is not None:
            self.graphs[0].replay()
            self.allreduce(self.w.counts)
            self.allreduce(self.w.sums)
            self.graphs[1].replay()
            self.allreduce(self.engine.store.grad)
        else:
            self.graphs[0].replay()
        self._adam()
        self.graphs[-1].replay()
        return self.w.loss
