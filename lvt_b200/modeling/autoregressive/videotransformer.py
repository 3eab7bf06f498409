"""`VideoTransformer` with the reference's constructor / forward surface
(vidgen/modeling/autoregressive/videotransformer.py:190-248, build.py:6-29), executing on VTEngine.
Parameters / buffers carry the reference's state_dict names and shapes."""
import logging

import numpy as np
import torch

from ..._lib import check, ptr, stream_ptr
from torch import nn

from ...utils.registry import Registry
from ..param_tree import ParamTree, attach_store
from .vt_engine import VTEngine, VTSpec

AUTOREGRESSIVE_REGISTRY = Registry("AUTOREGRESSIVE")


class Autoregressive(ParamTree):
    """Base class of autoregressive models (reference autoregressive/autoregressive.py:8-27)."""


def build_autoregressive(cfg, **kwargs):
    """`cfg.MODEL.AUTOREGRESSIVE.NAME` -> instance (reference autoregressive/build.py:16-29)."""
    model = AUTOREGRESSIVE_REGISTRY.get(cfg.MODEL.AUTOREGRESSIVE.NAME).from_config(cfg, **kwargs)
    assert isinstance(model, Autoregressive)
    logging.getLogger(__name__).info(
        "#params in autoregressive: {}M".format(sum(p.numel() for p in model.parameters()) / 1e6))
    return model


@AUTOREGRESSIVE_REGISTRY.register()
class VideoTransformer(Autoregressive):
    @classmethod
    def from_config(cls, cfg, **kwargs):
        vt = cfg.MODEL.AUTOREGRESSIVE.VT
        return cls(nc=vt.NC, nv=vt.NV, kernel_size=vt.KERNEL, stride=vt.STRIDE, d=vt.D, da=vt.DA, de=vt.DE,
                   blocks_e=vt.BLOCKS_E, n_head_e=vt.N_HEAD_E, blocks_d=vt.BLOCKS_D, n_head_d=vt.N_HEAD_D,
                   pad_value=vt.PAD_VALUE, share_p=vt.SHARE_P, share_embeddings=vt.SHARE_EMBEDDINGS,
                   class_num=vt.CLASS_NUM, device=cfg.MODEL.DEVICE)

    def __init__(self, nc, nv, da, de, d, blocks_e, n_head_e, kernel_size, stride, blocks_d, n_head_d, pad_value,
                 share_p, share_embeddings, class_num, device="cuda"):
        super().__init__()
        self.nv = nv
        spec = VTSpec(nc=nc, nv=nv, kernel=kernel_size, stride=stride, de=de, d=d, da=da, blocks_e=blocks_e,
                      heads_e=n_head_e, blocks_d=blocks_d, heads_d=n_head_d, pad_value=pad_value, share_p=share_p,
                      share_embeddings=share_embeddings, class_num=class_num)
        object.__setattr__(self, "engine", VTEngine(spec, device))  # not a sub-module
        attach_store(self, self.engine.store)
        self._register_reference_buffers(spec)
        self._reset_parameters(spec)

    # buffers the reference keeps in its state_dict (vt_attention.py:24,146-167; not used by the kernels,
    # which derive the same quantities from the block shape)
    def _register_reference_buffers(self, spec):
        dev = self.engine.device
        t, h, w = spec.block
        L = t * h * w
        idx = torch.arange(L)
        comp = {"dt": idx // (h * w), "dh": (idx // w) % h, "dw": idx % w}
        for side, n in (("encoder", len(spec.blocks_e)), ("decoder", len(spec.blocks_d))):
            n_ts = (spec.de if side == "encoder" else spec.d) // 6
            inc = np.log(1.0e4) / n_ts
            self.add_buffer(f"{side}.positional_encoder.inv_timescales",
                            torch.exp(torch.arange(n_ts).float() * -inc).to(dev))
            for i in range(n):
                for name, c in comp.items():
                    diff = c[:, None] - c[None, :]
                    self.add_buffer(f"{side}.block_local_attention.{i}.{name}", (diff - diff.min()).reshape(-1).to(dev))
                if side == "decoder":
                    self.add_buffer(f"{side}.block_local_attention.{i}.mask",
                                    torch.triu(torch.ones(1, 1, L, L), diagonal=1).to(dev))

    @torch.no_grad()
    def _reset_parameters(self, spec):
        """Default initialisation of the reference modules (nn.Conv3d / nn.Linear / nn.Embedding / LayerNorm
        defaults, xavier_normal_ for w_q/w_k/w_v/proj, zero banks, MaskedConv3d weight = ones:
        vt_attention.py:107-111,142-144, vt_utils.py:192-193); VideoTransformerModel.init_weights then applies
        MODEL.INIT_TYPE on top, as in the reference."""
        for name, p in self.named_parameters():
            if name.endswith("_bank"):
                p.zero_()
            elif "layer_norm.weight" in name or name.endswith("ffn.0.weight"):
                p.fill_(1.0)
            elif "layer_norm.bias" in name or name.endswith("ffn.0.bias"):
                p.zero_()
            elif ".mha.w_" in name or name.endswith("mha.proj.weight"):
                nn.init.xavier_normal_(p)
            elif "embed" in name:
                p.normal_()
            elif name == "decoder.conv.conv.weight":
                p.fill_(1.0)
            elif p.dim() > 1:
                nn.init.kaiming_uniform_(p.view(p.shape[0], -1), a=5 ** 0.5)
            else:
                p.uniform_(-0.05, 0.05)
        self.engine.store.p["decoder.conv.conv.weight"][:, :, -1, -1, 1:] = 0
        self.engine.shadows_fresh = False

    def load_state_dict(self, state_dict, strict=True):
        out = super().load_state_dict(state_dict, strict=strict)
        with torch.no_grad():
            self.engine.store.p["decoder.conv.conv.weight"][:, :, -1, -1, 1:] = 0
        self.engine.shadows_fresh = False
        return out

    def _stage(self, context, slc, slice_idx, ignore_mask, train, class_idx=None):
        eng = self.engine
        B = context.shape[0]
        ws = eng.workspace(B, tuple(slc.shape[2:]), tuple(context.shape[2:]), train=train)
        eng.set_inputs(ws, context, slc, slice_idx, ignore_mask, class_idx=class_idx)
        return ws

    def forward(self, context, slice, slice_idx, mode="logits", pixel=None, zl=None, temp=1.0, drop_mask=None,
                class_idx=None):
        """context (b, nc, T, H, W), slice (b, nc, t, h, w), slice_idx (b,) int64 — as in the reference.
        mode "logits": list of nc tensors (b, nv, t, h, w).  mode "sample_pixel": (codes (b, nc), zl)."""
        eng, spec = self.engine, self.engine.spec
        b = context.shape[0]
        t, h, w = slice.shape[2:]
        if mode == "logits":
            ws = self._stage(context, slice, slice_idx, None, train=False, class_idx=class_idx)
            eng.forward(ws, train=False, want_loss=False)
            lg = ws.logits.view(spec.nc, b, t, h, w, spec.nv)
            return [lg[k].permute(0, 4, 1, 2, 3).contiguous() for k in range(spec.nc)]
        if mode == "sample_pixel":
            ws = self._stage(context, slice, slice_idx, None, train=False, class_idx=class_idx)
            if zl is None:
                eng.encoder_forward(ws, train=False)  # cached across the pixels of one slice, like `zl`
                zl = ws
            eng.decoder_forward(ws, train=False)
            ti, hi, wi = pixel
            pos = (ti * h + hi) * w + wi
            out = torch.zeros(b, spec.nc, dtype=torch.int64, device=ws.slice.device)
            for k in range(spec.nc):
                eng.predictor_forward(ws, channels=[k])
                logits = ws.logits[k].view(b, t * h * w, spec.nv)[:, pos]
                prob = torch.softmax(logits / temp, 1)
                sample = torch.multinomial(prob, 1).squeeze(-1)
                out[:, k] = sample
                ws.slice.view(b, spec.nc, -1)[:, k, pos] = sample  # feeds the one-hot half of U[k+1]
            return out, zl
        raise ValueError("|mode| is invalid")

    @torch.no_grad()
    def sample_slice(self, context, slice, slice_idx, prime_mask=None, temp=1.0, use_graph=True, incremental=None,
                     class_idx=None):
        """Every non-primed position of one slice in raster order — the inner loops of
        VideoTransformerModel.sample_video (meta_arch/vt.py:107-134) around mode "sample_pixel"
        (videotransformer.py:161-185, 240-246): encoder once per slice, then per position the masked decoder and the
        channel-by-channel multinomial draw.  The per-position step reads its position from a device scalar, so it is
        captured ONCE in a CUDA graph and replayed.  incremental (default for <= 16 sequences): one ROW per sequence
        goes through the decoder against cached keys / values (IncrementalDecoder) instead of the full 256-token pass.
        Returns the completed slice (b, nc, t, h, w) int64."""
        from .incremental import MAX_ROWS, IncrementalDecoder
        eng, spec = self.engine, self.engine.spec
        b = context.shape[0]
        t, h, w = slice.shape[2:]
        thw = t * h * w
        primed = torch.zeros(thw, dtype=torch.bool) if prime_mask is None else prime_mask.reshape(-1).cpu()
        todo = [p for p in range(thw) if not bool(primed[p])]
        if not todo:  # a fully given slice: nothing to sample, no encoder pass either
            return slice.clone()
        ws = self._stage(context, slice, slice_idx, None, train=False, class_idx=class_idx)
        eng.encoder_forward(ws, train=False)
        if incremental is None:
            incremental = b <= 8  # measured: 0.50 vs 0.61 ms/position at b = 1, break-even near b = 8
        if ws.tiled or spec.share_embeddings:
            # slice larger than the attention block (the K/V-cached row decoder assumes one block per slice), or the
            # two-stage output projection of SHARE_EMBEDDINGS: the full decoder pass per position
            incremental = False
        assert not incremental or b <= MAX_ROWS
        cache = self.__dict__.setdefault("_sample_graphs", {})
        key = (id(ws), float(temp), bool(incremental))
        if key not in cache:
            if incremental:
                dec = IncrementalDecoder(eng, ws)
                pos_t = dec.pos
                if dec.fused:   # one persistent kernel per position (csrc/decode_step.cu)
                    steps = {"sample": lambda: dec.sample_row_fused(temp), "fill": dec.decode_row_fused}
                else:
                    steps = {"sample": lambda: dec.sample_row(temp), "fill": dec.decode_row}
            else:
                dec = None
                pos_t = torch.zeros(1, dtype=torch.int64, device=ws.slice.device)
                q_exp = torch.empty((b, spec.nv), dtype=torch.float32, device=ws.slice.device)

                def full_step():
                    eng.decoder_forward(ws, train=False)
                    for k in range(spec.nc):
                        eng.predictor_forward(ws, channels=[k])
                        # torch.multinomial(softmax(logits / temp), 1) == argmax(probs / q), q ~ Exp(1): q from torch's
                        # generator (same random stream as the reference's call), the rest in one kernel that writes
                        # the code into the slice buffer (it feeds the one-hot half of U[k+1] and the next positions)
                        q_exp.exponential_(1)
                        check(eng.lib.lvt_vt_sample_pixel(ptr(ws.logits[k]), ptr(q_exp), ptr(pos_t), ptr(ws.slice), b, thw,
                                                          spec.nv, spec.nc, k, float(temp), thw, stream_ptr()),
                              "lvt_vt_sample_pixel")
                steps = {"sample": full_step}
            cache[key] = {"dec": dec, "pos": pos_t, "steps": steps, "graphs": {}}
        entry = cache[key]
        dec, pos_t = entry["dec"], entry["pos"]
        if dec is not None:
            dec.begin_slice()

        def run(kind, p):
            pos_t.fill_(p)
            if not use_graph:
                entry["steps"][kind]()
                return
            if kind not in entry["graphs"]:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    entry["steps"][kind]()  # eager warm-up (kernel attributes, TMA maps, cached tables); redone below
                torch.cuda.current_stream().wait_stream(side)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    entry["steps"][kind]()
                entry["graphs"][kind] = g
            entry["graphs"][kind].replay()

        for p in range(todo[-1] + 1):
            if bool(primed[p]):
                if dec is not None:
                    run("fill", p)  # given position: only its keys / values are needed by the later rows
            else:
                run("sample", p)
        return ws.slice.view(b, spec.nc, t, h, w).clone()
