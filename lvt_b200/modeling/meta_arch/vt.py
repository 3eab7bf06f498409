"""VideoTransformerModel: the reference's meta-architecture surface (vidgen/modeling/meta_arch/vt.py:21-328)
on top of the B200 engine: forward(data: list[dict], mode), sample_video(s), calculate_logits_for_entire_video,
configure_optimizers_and_checkpointers, wrap_parallel."""
import os

import numpy as np
import torch
from torch import nn

from ...data.slices import slice_mask, ss_shift, subscale_order, visible_abc_mask
from ...solver import build_lr_scheduler, build_optimizer
from ...utils import comm
from ...utils.events import get_event_storage
from ..autoregressive import build_autoregressive
from .build import META_ARCH_REGISTRY


class _SupervisedLoss(torch.autograd.Function):
    """Bridges the engine into autograd: forward = whole network + cross-entropy, backward = engine.backward
    (gradients land in the flat buffer every Parameter.grad is a view of).  loss.backward() is expected to be
    called with the default upstream gradient of 1 (Trainer.run_step, trainer.py:80-82)."""

    @staticmethod
    def forward(ctx, anchor, engine, ws):
        ctx.engine, ctx.ws = engine, ws
        engine.forward(ws, train=True)
        return ws.loss[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.engine.backward(ctx.ws)
        if ctx.engine.grad_hook is not None:
            ctx.engine.grad_hook(ctx.engine.store.grad)
        return None, None, None


class _GraphedSupervisedLoss(torch.autograd.Function):
    """The same bridge with the ~400 launches of forward + backward replayed as ONE CUDA graph at forward time
    (Trainer.run_step always calls loss.backward() right after model(data, 'supervised'), trainer.py:79-82, and the
    gradients are accumulated into the flat buffer exactly as engine.backward does); backward() is left with the
    data-parallel gradient hook."""

    @staticmethod
    def forward(ctx, anchor, engine, ws, graph):
        ctx.engine = engine
        graph.replay()
        return ws.loss[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.engine.grad_hook is not None:
            ctx.engine.grad_hook(ctx.engine.store.grad)
        return None, None, None, None


@META_ARCH_REGISTRY.register()
class VideoTransformerModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.device = torch.device(cfg.MODEL.DEVICE)
        self.model = build_autoregressive(cfg)
        self.init_weights(self.model, cfg.MODEL.INIT_TYPE)
        self.vis_period = cfg.VIS_PERIOD
        # LVT_SAMPLER_GRAPH=0: the reference's per-pixel Python loop (eager launches, eager RNG stream);
        # sampler_graph = "eager": the fused per-position step without graph capture (eager RNG stream)
        self.sampler_graph = os.environ.get("LVT_SAMPLER_GRAPH", "1") != "0"
        self.model.engine.grad_hook = None
        self._anchor = torch.zeros(1, device=self.device, requires_grad=True)
        self._graphed, self._graphs = False, {}

    def enable_graphed_step(self, on=True):
        """Trainer fast path: model(data, 'supervised') stages the batch into the engine's static buffers and
        replays forward + backward as one CUDA graph (captured on first use per batch shape) instead of launching
        ~400 kernels from the host.  Only valid when every supervised forward is followed by loss.backward(), which
        is what Trainer.run_step does."""
        self._graphed = bool(on)

    def _graph_for(self, eng, ws):
        if not eng.shadows_fresh:            # (inside a graph this host-side check would be frozen)
            eng.refresh_shadows()
        g = self._graphs.get(id(ws))
        if g is None:
            grad0 = eng.store.grad.clone()   # a capture must not leave warm-up gradients behind
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):           # eager warm-up: kernel attributes, TMA maps, scratch allocations
                    eng.forward(ws, train=True)
                    eng.backward(ws)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            eng.store.grad.copy_(grad0)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                eng.forward(ws, train=True)
                eng.backward(ws)
            self._graphs[id(ws)] = g
        return g

    @staticmethod
    @torch.no_grad()
    def init_weights(module, init_type="normal", slope=0.2):
        """vt.py:34-57: the reference re-initialises modules whose class name contains Conv / Linear
        (weights `normal` or `xavier_uniform`, biases 0).  Its second loop only visits the DIRECT children of
        VideoTransformer (encoder, decoder, ch_predictor), none of which defines init_weights, so
        MultiHeadAttention.init_weights is never re-run: w_q/w_k/w_v keep the xavier_normal_ draw of their
        constructor (vt_attention.py:106-112) and mha.proj.weight (an nn.Linear) keeps the INIT_TYPE init."""
        conv_or_linear = ("encoder.conv.", "encoder.linear_projector.", "decoder.conv.conv.", "decoder.linear_projector.",
                          ".mha.proj.", ".ffn.1.", ".ffn.3.", "ch_predictor.U.", "ch_predictor.P.")
        for name, p in module.named_parameters():
            if not any(tag in name for tag in conv_or_linear):
                continue
            if name.endswith("weight"):
                if init_type == "normal":
                    std = 1 / np.sqrt((1 + slope ** 2) * np.prod(p.shape[:-1]))
                    p.normal_(std=std)
                elif init_type == "xavier_uniform":
                    nn.init.xavier_uniform_(p)
                else:
                    raise ValueError
            elif name.endswith("bias"):
                p.zero_()
        eng = module.engine
        eng.store.p["decoder.conv.conv.weight"][:, :, -1, -1, 1:] = 0
        eng.shadows_fresh = False

    def train(self, mode=True):
        self.training = mode
        self.model.train(mode)
        return self

    def wrap_parallel(self, device_ids, broadcast_buffers):
        """Reference: DistributedDataParallel(self.model) (vt.py:61-63).  Here: broadcast rank 0's flat
        parameter buffer once, then ONE all-reduce of the flat fp32 gradient per backward, averaged."""
        eng = self.model.engine
        if comm.get_world_size() > 1:
            torch.distributed.broadcast(eng.store.master, src=0)
            eng.shadows_fresh = False
            world = comm.get_world_size()

            def hook(flat_grad):
                torch.distributed.all_reduce(flat_grad)
                flat_grad.div_(world)
            eng.grad_hook = hook

    # ------------------------------------------------------------------ data
    def preprocess_data(self, data):
        """vt.py:284-299 (host tensors are staged into the engine's static device buffers)."""
        context = torch.stack([x["context"] for x in data], 0)
        slc = torch.stack([x["slice"] for x in data], 0)
        slice_idx = torch.stack([x["slice_idx"] for x in data], 0)
        ignore_mask = torch.stack([x["ignore_mask"] for x in data], 0)
        class_idx = torch.stack([x["class"] for x in data], 0) if "class" in data[0] else None
        return context, slc, slice_idx, ignore_mask, class_idx

    def forward(self, data, mode="inference"):
        if mode == "supervised":
            context, slc, slice_idx, ignore_mask, class_idx = self.preprocess_data(data)
            get_event_storage()  # the reference requires an active EventStorage in training modes (vt.py:186)
            return self.compute_supervised_loss(context, slc, slice_idx, ignore_mask, None, class_idx)
        if mode == "inference":
            output = [{} for _ in range(len(data))]
            if "BitsEvaluator" in self.cfg.TEST.EVALUATORS:
                output = self.calculate_logits_for_entire_video(data, output)
            if "VTSampler" in self.cfg.TEST.EVALUATORS:
                output = self.sample_videos(data, output, n_prime=self.cfg.TEST.VT_SAMPLER.N_PRIME,
                                            num_samples=self.cfg.TEST.VT_SAMPLER.NUM_SAMPLES)
            assert len(output[0]) > 0
            return output
        raise ValueError("|mode| is invalid")

    def compute_supervised_loss(self, context, slc, slice_idx, ignore_mask, iter, class_idx):
        """vt.py:301-314: mean over channels of CE(pred_k, target_k), ignore_index = MODEL.IGNORE_INDEX."""
        eng = self.model.engine
        ws = eng.workspace(context.shape[0], tuple(slc.shape[2:]), tuple(context.shape[2:]), train=True)
        eng.set_inputs(ws, context, slc, slice_idx, ignore_mask, class_idx=class_idx)
        if self._graphed:
            return {"loss_cross_entropy": _GraphedSupervisedLoss.apply(self._anchor, eng, ws, self._graph_for(eng, ws))}
        return {"loss_cross_entropy": _SupervisedLoss.apply(self._anchor, eng, ws)}

    # ------------------------------------------------------------------ evaluation / sampling
    def _slices(self):
        vt = self.cfg.MODEL.AUTOREGRESSIVE.VT
        return tuple(vt.STRIDE), tuple(vt.KERNEL), vt.PAD_VALUE

    @torch.no_grad()
    def calculate_logits_for_entire_video(self, data, output):
        """vt.py:230-282: teacher-forced logits of every slice, scattered back to (nc, nv, T, H, W)."""
        video = torch.stack([torch.as_tensor(x["image_sequence"]) for x in data], 0).to(self.device)
        B, T, nc, H, W = video.shape
        video = video.transpose(1, 2).contiguous()
        (st, sh, sw), kernel, pad_value = self._slices()
        n_prime, nv = self.cfg.MODEL.AUTOREGRESSIVE.VT.N_PRIME, self.cfg.MODEL.AUTOREGRESSIVE.VT.NV
        idx2abc, _ = subscale_order(st, sh, sw)
        logits = torch.zeros(B, nc, nv, T, H, W, device=self.device)
        for slice_idx, (a, b, c) in enumerate(idx2abc):
            slc = video[:, :, a::st, b::sh, c::sw].contiguous()
            vmask = visible_abc_mask(a, b, c, st, sh, sw, T, H, W, dtype=torch.bool, device=self.device)
            context = ss_shift(video.masked_fill(~vmask, pad_value), a, b, c, st, sh, sw, T, H, W, *kernel,
                               pad_value=pad_value)
            sidx = torch.full((B,), slice_idx, dtype=torch.int64, device=self.device)
            pred = self.model(context, slc, sidx, mode="logits")
            for k in range(nc):
                logits[:, k, :, a::st, b::sh, c::sw] = pred[k]
        ignore_mask = torch.zeros(1, T, H, W, dtype=torch.bool, device=self.device)
        if n_prime > 0:
            ignore_mask[:, :n_prime] = True
        for i in range(B):
            output[i]["ignore_mask"] = ignore_mask
            output[i]["logits"] = logits[i]
        return output

    @torch.no_grad()
    def sample_video(self, video, temp=1.0, n_prime=1, class_idx=None):
        """vt.py:81-136: slice by slice, position by position, channel by channel (torch.multinomial)."""
        video = video.to(self.device)
        (st, sh, sw), kernel, pad_value = self._slices()
        idx2abc, _ = subscale_order(st, sh, sw)
        B, nc, T, H, W = video.shape
        t, h, w = T // st, H // sh, W // sw
        prime = torch.zeros(T, H, W, dtype=torch.bool)
        if n_prime > 0:
            prime[:n_prime] = True
        for slice_idx, (a, b, c) in enumerate(idx2abc):
            slc = video[:, :, a::st, b::sh, c::sw].contiguous()
            prime_slice = prime[a::st, b::sh, c::sw]
            vmask = visible_abc_mask(a, b, c, st, sh, sw, T, H, W, dtype=torch.bool, device=self.device)
            context = ss_shift(video.masked_fill(~vmask, pad_value), a, b, c, st, sh, sw, T, H, W, *kernel,
                               pad_value=pad_value)
            sidx = torch.full((B,), slice_idx, dtype=torch.int64, device=self.device)
            if self.sampler_graph:
                # same loops, one CUDA-graph replay per position (VideoTransformer.sample_slice)
                slc = self.model.sample_slice(context, slc, sidx, prime_slice, temp=temp,
                                              use_graph=self.sampler_graph != "eager",
                                              incremental=getattr(self.model, "sample_incremental", None),
                                              class_idx=class_idx)
            else:
                zl = None
                for ti in range(t):
                    for hi in range(h):
                        for wi in range(w):
                            if bool(prime_slice[ti, hi, wi]):
                                continue
                            pred, zl = self.model(context, slc, sidx, mode="sample_pixel", pixel=(ti, hi, wi), zl=zl,
                                                  temp=temp, class_idx=class_idx)
                            slc[:, :, ti, hi, wi] = pred
            video[:, :, a::st, b::sh, c::sw] = slc
        return video

    @torch.no_grad()
    def sample_videos(self, data, output, n_prime=5, num_samples=1):
        """vt.py:210-228."""
        video = torch.stack([torch.as_tensor(x["image_sequence"]) for x in data], 0).to(self.device)
        video = video.transpose(1, 2).contiguous()
        video[:, :, n_prime:] = 0
        samples = [self.sample_video(video.clone(), n_prime=n_prime) for _ in range(num_samples)]
        for i in range(video.size(0)):
            output[i]["samples"] = [s[i] for s in samples]
        return output

    # ------------------------------------------------------------------ optimisation
    def configure_optimizers_and_checkpointers(self):
        """vt.py:316-328."""
        from ...engine.checkpoint import Checkpointer
        import os
        optimizer_g = build_optimizer(self.model, self.cfg, suffix="_G")
        scheduler_g = build_lr_scheduler(self.cfg, optimizer_g)
        os.makedirs(os.path.join(self.cfg.OUTPUT_DIR, "netG"), exist_ok=True)
        c = [{"checkpointer": Checkpointer(self.model, os.path.join(self.cfg.OUTPUT_DIR, "netG")),
              "pretrained": self.cfg.MODEL.GENERATOR.WEIGHTS}]
        o = [{"optimizer": optimizer_g, "scheduler": scheduler_g, "type": "generator"}]
        return o, c
