"""Incremental (K/V-cached) decoding of one slice, position by position.

The reference samples a slice by running the whole decoder for every position (meta_arch/vt.py:107-134 around
videotransformer.py:240-246): 256 full passes over 256 tokens.  Row p of every masked decoder layer only depends on rows
<= p, so this module keeps the keys / values of the rows already decoded and pushes ONE row per sequence through the
decoder stack: per position ~60 skinny launches (csrc/sampler.cu) instead of ~90 tensor-core GEMM launches on 256-row
tiles.  Activations that are bf16 in the full pass are rounded at the same places, so the logits of the two paths differ
by summation order only (checked in tests/test_api_gpu.py)."""
import ctypes
import math
import os

import torch

from ... import _lib
from ..._lib import LvtDecodeStep, LvtRowsLinear, check, ptr, stream_ptr
from ...ops import Operand, gemm
from .vt_engine import LN_EPS, _vp

MAX_ROWS = 16  # sequences per call (csrc/sampler.cu keeps the rows in registers / shared memory)


class IncrementalDecoder:
    """Bound to one engine workspace (static buffers), so its per-position step can be captured in a CUDA graph."""

    def __init__(self, engine, ws):
        s = engine.spec
        assert ws.B <= MAX_ROWS
        if ws.tiled or s.share_embeddings:
            raise _lib.LvtError("IncrementalDecoder: the slice must be one attention block and SHARE_EMBEDDINGS off "
                                "(use the full decoder pass)")
        self.eng, self.ws = engine, ws
        dev = ws.slice.device
        B, d, H, da, L, nv = ws.B, s.d, s.H, s.da, ws.thw, s.nv
        nD = len(s.blocks_d)
        f32, bf16 = torch.float32, torch.bfloat16
        e = lambda shape, dt=f32: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
        self.kc = [e((B, H, L, da), bf16) for _ in range(nD)]
        self.vc = [e((B, H, L, da), bf16) for _ in range(nD)]
        self.xa, self.xb, self.h, self.a1 = e((B, d)), e((B, d)), e((B, d)), e((B, d))
        self.q, self.o = e((B, H * da)), e((B, H * da))
        self.a = e((B, d))
        self.logits = e((B, nv))
        self.q_exp = e((B, nv))
        self.y0s = e((ws.M, d))  # positional encoding + conv bias + zl Wlp^T: the part of y0 that does not depend on the slice
        self.pos = torch.zeros(1, dtype=torch.int64, device=dev)
        # the whole per-position step as ONE persistent kernel (csrc/decode_step.cu) instead of ~58 launches: measured
        # 0.44 against 0.50 ms per position for one sequence, slower from two sequences on (32 CTAs, 2.5 us per grid
        # barrier), so it is the default for B == 1 only; LVT_SAMPLER_FUSED=1 / 0 forces it on / off (same arithmetic,
        # same sampled codes either way)
        env = os.environ.get("LVT_SAMPLER_FUSED", "")
        self.fused = (env == "1" or (env != "0" and B == 1)) and nD <= 8 and s.nc <= 4
        self.q_exp4 = e((s.nc, B, nv))
        self._barrier = torch.zeros(2, dtype=torch.int32, device=dev)
        self._desc = {}

    # ------------------------------------------------------------------ helpers
    def _rows(self, x, K, w, w_ld, N, out, *, x_ldb=None, x_pos_mul=0, x_bf16=False, ln=None, round_in=False, bias=None,
              res=None, res_ldb=0, res_pos_mul=0, gtab=None, g_count=0, relu=False, round_out=False):
        s, ws = self.eng.spec, self.ws
        a = LvtRowsLinear()
        a.B, a.N, a.K = ws.B, N, K
        a.x, a.x_ldb, a.x_pos_mul, a.x_bf16 = _vp(x), (K if x_ldb is None else x_ldb), x_pos_mul, int(x_bf16)
        if ln is not None:
            a.ln_gamma, a.ln_beta, a.ln_eps = _vp(ln[0]), _vp(ln[1]), LN_EPS
        a.round_in = int(round_in)
        a.w_bf16, a.w_ld = _vp(w), w_ld
        a.bias = _vp(bias)
        a.res, a.res_ldb, a.res_pos_mul = _vp(res), res_ldb, res_pos_mul
        if g_count:
            a.gtab, a.slice, a.g_count, a.nv, a.nc, a.thw = _vp(gtab), _vp(ws.slice), g_count, s.nv, s.nc, ws.thw
        a.relu, a.round_out = int(relu), int(round_out)
        a.out, a.out_ldb = _vp(out), N
        a.pos = _vp(self.pos)
        check(self.eng.lib.lvt_rows_linear(ctypes.byref(a), stream_ptr()), "lvt_rows_linear")

    # ------------------------------------------------------------------ per slice
    def begin_slice(self):
        """After encoder_forward: y0s = zl Wlp^T + positional encoding + conv bias (videotransformer.py:96-99)."""
        eng, ws, st = self.eng, self.ws, self.eng.store
        d = eng.spec.d
        gemm(ws.M, d, d, Operand(ws.zl_bf16.data_ptr(), d), Operand(st.pb("decoder.linear_projector.weight"), d),
             Operand(self.y0s.data_ptr(), d), out_f32=self.y0s, bias=eng.posenc_table(ws.slice_shape), bias_mod=ws.thw)
        self.y0s.add_(st.p["decoder.conv.conv.bias"])

    # ------------------------------------------------------------------ per position (graph-capturable)
    def decode_row(self):
        """Row *pos of the decoder stack -> self.xa (= y_final[b, pos, :]); K/V caches of every layer updated."""
        eng, ws, st, s = self.eng, self.ws, self.eng.store, self.eng.spec
        d, H, da, de, nc, nv = s.d, s.H, s.da, s.de, s.nc, s.nv
        t, h, w = ws.slice_shape
        taps, offs, wp, _ = eng._live_taps(ws.slice_shape)
        ntaps = len(taps)
        lib = eng.lib
        # embed-sum + causal im2col of the current slice content (all positions: one small launch), then the row's conv
        check(lib.lvt_vt_dec_front_fwd(ptr(ws.slice), _vp(st.pf("decoder.ch_embedder.0.weight")), ptr(offs), ptr(ws.A0),
                                       ws.B, nc, nv, de, t, h, w, ntaps, stream_ptr()), "lvt_vt_dec_front_fwd")
        K0 = ntaps * de
        self._rows(ws.A0, K0, wp, K0, d, self.xa, x_ldb=ws.thw * K0, x_pos_mul=K0, x_bf16=True,
                   res=self.y0s, res_ldb=ws.thw * d, res_pos_mul=d)
        nE = len(s.blocks_e)
        scale = 1.0 / math.sqrt(da)
        x, y = self.xa, self.xb
        for i in range(len(s.blocks_d)):
            pre = f"decoder.block_local_attention.{i}."
            check(lib.lvt_rows_qkv(ptr(x), _vp(st.pf(pre + "mha.layer_norm.weight")), _vp(st.pf(pre + "mha.layer_norm.bias")),
                                   LN_EPS, _vp(st.pb(pre + "mha.w_q")), ptr(self.q), ptr(self.kc[i]), ptr(self.vc[i]),
                                   ptr(self.pos), ws.B, H, d, da, ws.thw, stream_ptr()), "lvt_rows_qkv")
            check(lib.lvt_attn_row(ptr(self.q), ptr(self.kc[i]), ptr(self.vc[i]), _vp(st.pf(pre + "dt_bank")),
                                   _vp(st.pf(pre + "dh_bank")), _vp(st.pf(pre + "dw_bank")), s.block[0], s.block[1],
                                   s.block[2], ptr(self.pos), scale, ptr(self.o), ws.B, H, ws.thw, da, stream_ptr()),
                  "lvt_attn_row")
            self._rows(self.o, H * da, st.pb(pre + "mha.proj.weight"), H * da, d, self.h, res=x, res_ldb=d)
            self._rows(self.h, d, st.pb(pre + "ffn.1.weight"), d, d, self.a1,
                       ln=(st.pf(pre + "ffn.0.weight"), st.pf(pre + "ffn.0.bias")), round_in=True,
                       bias=st.pf(pre + "ffn.1.bias"), relu=True, round_out=True)
            self._rows(self.a1, d, st.pb(pre + "ffn.3.weight"), d, d, y, bias=st.pf(pre + "ffn.3.bias"), res=self.h,
                       res_ldb=d)
            x, y = y, x
        self.y_final = x
        del nE

    # ------------------------------------------------------------------ fused per-position step
    def _fused_desc(self, do_sample, temp):
        key = (bool(do_sample), float(temp))
        if key in self._desc:
            return self._desc[key]
        eng, ws, st, s = self.eng, self.ws, self.eng.store, self.eng.spec
        taps, offs, wp, _ = eng._live_taps(ws.slice_shape)
        t, h, w = ws.slice_shape
        p = LvtDecodeStep()
        p.B, p.d, p.H, p.da, p.L, p.nc, p.nv, p.de = ws.B, s.d, s.H, s.da, ws.thw, s.nc, s.nv, s.de
        p.ntaps, p.n_layers = len(taps), len(s.blocks_d)
        p.bt, p.bh, p.bw = s.block
        p.t, p.h, p.w = t, h, w
        p.scale, p.ln_eps, p.temp, p.do_sample = 1.0 / math.sqrt(s.da), LN_EPS, float(temp), int(do_sample)
        p.pos, p.slice = self.pos.data_ptr(), ws.slice.data_ptr()
        p.emb, p.taps = st.pf("decoder.ch_embedder.0.weight"), offs.data_ptr()
        p.conv_w, p.y0s = wp.data_ptr(), self.y0s.data_ptr()
        for i in range(p.n_layers):
            pre, ly = f"decoder.block_local_attention.{i}.", p.layer[i]
            ly.ln1_g, ly.ln1_b = st.pf(pre + "mha.layer_norm.weight"), st.pf(pre + "mha.layer_norm.bias")
            ly.w_qkv = st.pb(pre + "mha.w_q")
            ly.k_cache, ly.v_cache = self.kc[i].data_ptr(), self.vc[i].data_ptr()
            ly.bank_t, ly.bank_h, ly.bank_w = st.pf(pre + "dt_bank"), st.pf(pre + "dh_bank"), st.pf(pre + "dw_bank")
            ly.w_proj = st.pb(pre + "mha.proj.weight")
            ly.ln2_g, ly.ln2_b = st.pf(pre + "ffn.0.weight"), st.pf(pre + "ffn.0.bias")
            ly.w_ffn1, ly.b_ffn1 = st.pb(pre + "ffn.1.weight"), st.pf(pre + "ffn.1.bias")
            ly.w_ffn3, ly.b_ffn3 = st.pb(pre + "ffn.3.weight"), st.pf(pre + "ffn.3.bias")
        p.lnp_g, p.lnp_b = st.pf("ch_predictor.layer_norm.weight"), st.pf("ch_predictor.layer_norm.bias")
        for k in range(s.nc):
            p.U[k], p.U_ld[k] = st.pb(f"ch_predictor.U.{k}.weight"), s.d + k * s.nv
            p.U_bias[k] = st.pf(f"ch_predictor.U.{k}.bias")
            p.gtab[k] = eng.ut[k].data_ptr() if k else None
            p.P[k], p.P_bias[k] = st.pb(s.p_name(k) + ".weight"), st.pf(s.p_name(k) + ".bias")
        p.q_exp = self.q_exp4.data_ptr()
        p.xa, p.xb, p.hbuf, p.a1 = self.xa.data_ptr(), self.xb.data_ptr(), self.h.data_ptr(), self.a1.data_ptr()
        p.q, p.o, p.abuf, p.logits = self.q.data_ptr(), self.o.data_ptr(), self.a.data_ptr(), self.logits.data_ptr()
        p.barrier = self._barrier.data_ptr()
        self._desc[key] = p
        return p

    def decode_row_fused(self):
        """decode_row as one kernel (primed positions: only the K/V caches of the row are needed)."""
        check(self.eng.lib.lvt_vt_decode_step(ctypes.byref(self._fused_desc(False, 1.0)), stream_ptr()), "lvt_vt_decode_step")

    def sample_row_fused(self, temp=1.0):
        """sample_row as one kernel; the Exp(1) noise of the four draws comes from torch's generator in the order of
        the reference's torch.multinomial calls, before the launch."""
        for k in range(self.eng.spec.nc):
            self.q_exp4[k].exponential_(1)
        check(self.eng.lib.lvt_vt_decode_step(ctypes.byref(self._fused_desc(True, temp)), stream_ptr()), "lvt_vt_decode_step")

    def channel_logits(self, k):
        """ChannelPredictor for channel k at *pos (videotransformer.py:144-160): self.logits [B, nv]."""
        eng, ws, st, s = self.eng, self.ws, self.eng.store, self.eng.spec
        d, nv = s.d, s.nv
        self._rows(self.y_final, d, st.pb(f"ch_predictor.U.{k}.weight"), d + k * nv, d, self.a,
                   ln=(st.pf("ch_predictor.layer_norm.weight"), st.pf("ch_predictor.layer_norm.bias")), round_in=True,
                   bias=st.pf(f"ch_predictor.U.{k}.bias"), gtab=eng.ut[k] if k else None, g_count=k, relu=True,
                   round_out=True)
        self._rows(self.a, d, st.pb(s.p_name(k) + ".weight"), d, nv, self.logits, bias=st.pf(s.p_name(k) + ".bias"))

    def sample_row(self, temp=1.0):
        """decode_row + channel-by-channel categorical draw into ws.slice[:, :, *pos] (videotransformer.py:161-185)."""
        eng, ws, s = self.eng, self.ws, self.eng.spec
        self.decode_row()
        for k in range(s.nc):
            self.channel_logits(k)
            self.q_exp.exponential_(1)
            check(eng.lib.lvt_vt_sample_pixel(ptr(self.logits), ptr(self.q_exp), ptr(self.pos), ptr(ws.slice), ws.B,
                                              ws.thw, s.nv, s.nc, k, float(temp), 1, stream_ptr()), "lvt_vt_sample_pixel")
