"""VQVAEModel: the reference's meta-architecture surface (vidgen/modeling/meta_arch/ae.py:21-244,
vqvae.py:16-124) on the B200 VQ-VAE engine."""
import os

import torch
from torch import nn

from ...solver import build_lr_scheduler, build_optimizer
from ...utils import comm
from ..vqvae_engine import VQVAEEngine, VQVAESpec
from ..vqvae_modules import DVQEmbedding, VQEmbedding, build_encoder, build_generator
from .build import META_ARCH_REGISTRY


class _SupervisedLoss(torch.autograd.Function):
    """forward = encoder + DVQ/EMA + decoder + both losses; backward = engine.backward (gradients are written to
    the flat buffer the Parameter.grad views point into).  Returns [loss_reconstruction, loss_commitment]."""

    @staticmethod
    def forward(ctx, anchor, model, ws):
        ctx.model, ctx.ws = model, ws
        eng = model.engine
        eng.forward_train(ws, allreduce=comm.all_reduce_sum_ if comm.get_world_size() > 1 else None)
        return ws.loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.model.engine
        eng.backward(ctx.ws)
        if comm.get_world_size() > 1:
            comm.all_reduce_sum_(eng.store.grad)
            eng.store.grad.div_(comm.get_world_size())
            if not eng.spec.ema:  # vqvae.py:46-50: the codebook is DDP-wrapped too when it is trained
                comm.all_reduce_sum_(eng.cb_grad)
                eng.cb_grad.div_(comm.get_world_size())
        return None, None, None


class _GraphedSupervisedLoss(torch.autograd.Function):
    """Trainer fast path: [encoder, codebook search] and [EMA update, decoder, losses, backward] replayed as two
    CUDA graphs at forward time (Trainer.run_step calls loss.backward() right after, trainer.py:79-82); the
    cross-rank sums (EMA statistics between the graphs, the flat gradient in backward()) stay eager."""

    @staticmethod
    def forward(ctx, anchor, model, ws, graphs):
        ctx.model = model
        graphs[0].replay()
        if comm.get_world_size() > 1:
            comm.all_reduce_sum_(ws.counts)
            comm.all_reduce_sum_(ws.sums)
        graphs[1].replay()
        return ws.loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.model.engine
        if comm.get_world_size() > 1:
            comm.all_reduce_sum_(eng.store.grad)
            eng.store.grad.div_(comm.get_world_size())
        return None, None, None, None


@META_ARCH_REGISTRY.register()
class VQVAEModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.device = torch.device(cfg.MODEL.DEVICE)
        m = cfg.MODEL
        assert len(m.PIXEL_MEAN) == len(m.PIXEL_STD)
        assert len(set(m.PIXEL_MEAN)) == 1 and len(set(m.PIXEL_STD)) == 1, "per-channel identical mean/std"
        spec = VQVAESpec(in_channels=m.ENCODER.IN_CHANNELS, nf=m.ENCODER.NF, res_channels=m.ENCODER.RES_CHANNELS,
                         n_layers=m.ENCODER.N_LAYERS, codebook_num=m.CODEBOOK.NUM, codebook_size=m.CODEBOOK.SIZE,
                         codebook_dim=m.CODEBOOK.DIM, beta=m.CODEBOOK.BETA, ema=m.CODEBOOK.EMA,
                         pixel_lambda=cfg.LOSS.PIXEL.LAMBDA, pixel_mean=m.PIXEL_MEAN[0], pixel_std=m.PIXEL_STD[0],
                         out_activation=m.GENERATOR.OUT_ACTIVATION)
        assert m.GENERATOR.N_LAYERS == m.ENCODER.N_LAYERS and cfg.LOSS.PIXEL.MODE == "l2"
        object.__setattr__(self, "engine", VQVAEEngine(spec, cfg.MODEL.DEVICE))
        self.encoder = build_encoder(cfg, engine=self.engine)
        self.generator = build_generator(cfg, engine=self.engine)
        self.init_weights(self.encoder, cfg.MODEL.INIT_TYPE)
        self.init_weights(self.generator, cfg.MODEL.INIT_TYPE)
        self.use_codebook_ema = cfg.MODEL.CODEBOOK.EMA
        # vqvae.py:26-31: one VQEmbedding (Base-VQVAE.yaml, CODEBOOK.NUM 1: latents (n, h, w)) or the DVQ stack
        if m.CODEBOOK.NUM == 1:
            self.codebook = VQEmbedding(self.engine, 0, self.use_codebook_ema, standalone=True)
        else:
            self.codebook = DVQEmbedding(self.engine, self.use_codebook_ema)
        self._single = m.CODEBOOK.NUM == 1
        self.beta = cfg.MODEL.CODEBOOK.BETA
        self.vis_period = cfg.VIS_PERIOD
        self._anchor = torch.zeros(1, device=self.device, requires_grad=True)
        self._graphed, self._graphs = False, {}
        # encode / inference run the encoder in the high-precision split form (VQVAEEngine.encode_precise): the code
        # indices are the wire format between the two models; LVT_VQVAE_PRECISE=0 selects the plain bf16 encoder
        self.precise_latents = os.environ.get("LVT_VQVAE_PRECISE", "1") != "0"
        self.back_normalizer = lambda y: y * spec.std + spec.mean
        self.normalizer = lambda x: (x - spec.mean) / spec.std

    @staticmethod
    @torch.no_grad()
    def init_weights(module, init_type="normal", slope=0.2):
        """ae.py:41-61: conv weights `normal` / `xavier_uniform`, biases 0."""
        for name, p in module.named_parameters():
            if name.endswith("weight"):
                if init_type == "normal":
                    p.normal_(std=1 / ((1 + slope ** 2) * float(torch.tensor(p.shape[:-1]).prod())) ** 0.5)
                elif init_type == "xavier_uniform":
                    nn.init.xavier_uniform_(p)
                else:
                    raise ValueError
            else:
                p.zero_()
        module.engine.shadows_fresh = False

    def train(self, mode=True):
        self.training = mode
        return self

    def enable_graphed_step(self, on=True):
        """Trainer fast path (see _GraphedSupervisedLoss): valid when every supervised forward is followed by
        loss.backward(), which is what Trainer.run_step does."""
        self._graphed = bool(on)

    def _graphs_for(self, w):
        eng = self.engine
        if not eng.shadows_fresh:            # (inside a graph this host-side check would be frozen)
            eng.refresh_shadows()
        g = self._graphs.get(id(w))
        if g is None:
            # the eager warm-up runs real forward / backward passes: the gradient buffer and the EMA codebook
            # state they touch are restored afterwards
            keep = (eng.store.grad, eng.codebook, eng.running_size, eng.running_sum)
            snap = [t.clone() for t in keep]
            allreduce = comm.all_reduce_sum_ if comm.get_world_size() > 1 else None
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    eng.forward_train(w, allreduce=allreduce)
                    eng.backward(w)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for dst, src in zip(keep, snap):
                dst.copy_(src)
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga):
                eng._forward_train_a(w)
            with torch.cuda.graph(gb):
                eng._forward_train_b(w)
                eng.backward(w)
            g = self._graphs[id(w)] = (ga, gb)
        return g

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.engine.shadows_fresh = False
        return out

    def wrap_parallel(self, device_ids, broadcast_buffers):
        """ae.py:69-73 wraps encoder and generator in DDP; here rank 0's flat parameters are broadcast once and
        the flat gradient is all-reduced once per backward (see _SupervisedLoss)."""
        if comm.get_world_size() > 1:
            torch.distributed.broadcast(self.engine.store.master, src=0)
            torch.distributed.broadcast(self.engine.codebook, src=0)
            self.engine.running_sum.copy_(self.engine.codebook)
            self.engine.shadows_fresh = False

    # ------------------------------------------------------------------ data
    def preprocess_data(self, data):
        """ae.py:151-168 without the normalisation (done inside the first kernel): (n, 3, 64, 64) in [0, 1] plus
        the (b, t) shape of sequences."""
        key = "image" if "image" in data[0] else "image_sequence"
        x = torch.stack([torch.as_tensor(d[key]) for d in data], 0).float()
        seq = None
        if x.dim() == 5:
            seq = x.shape[:2]
            x = x.reshape(-1, *x.shape[2:])
        return x, seq

    def _stage(self, x, train):
        w = self.engine.workspace(x.shape[0], train=train)
        w.x.copy_(x, non_blocking=True)
        return w

    def forward(self, data, mode="inference"):
        x, seq = self.preprocess_data(data)
        if mode in ("supervised", "generator"):
            w = self._stage(x, train=True)
            if self._graphed and self.use_codebook_ema:
                losses = _GraphedSupervisedLoss.apply(self._anchor, self, w, self._graphs_for(w))
            else:
                losses = _SupervisedLoss.apply(self._anchor, self, w)
            out = {"loss_reconstruction": losses[0], "loss_commitment": losses[1]}
            if not self.use_codebook_ema:  # vqvae.py:84-85 (the key really is 'loss_dict')
                out["loss_dict"] = losses[2]
            return out
        if mode == "inference":
            w = self._stage(x, train=False)
            recon, idx = self.engine.inference(w, precise=self.precise_latents)
            recon, idx = recon.clone(), idx.clone()
            if self._single:
                idx = idx[:, 0]
            if seq is not None:
                recon = recon.view(*seq, *recon.shape[1:])
                idx = idx.view(*seq, *idx.shape[1:])
            return [{"reconstruction": recon[i], "latent": idx[i]} for i in range(recon.shape[0])]
        raise ValueError("|mode| is invalid")

    @torch.no_grad()
    def encode(self, x01):
        """vqvae.py:93-101 on images in [0, 1]: (n, 3, 64, 64) or (b, t, 3, 64, 64) -> int64 codes."""
        seq = x01.shape[:2] if x01.dim() == 5 else None
        x = x01.reshape(-1, *x01.shape[-3:])
        w = self._stage(x, train=False)
        if self.precise_latents:
            self.engine.encode_precise(w)
        else:
            self.engine.encode(w)
        self.engine.quantize(w, train=False)
        idx = w.idx.clone()
        if self._single:
            idx = idx[:, 0]
        return idx.view(*seq, *idx.shape[1:]) if seq is not None else idx

    @torch.no_grad()
    def decode(self, latents):
        """vqvae.py:103-106: codes (n, num, 16, 16) -> x_tilde in [-1, 1] (n, 3, 64, 64)."""
        latents = latents.to(self.device)
        if self._single:
            latents = latents.unsqueeze(1)
        w = self.engine.workspace(latents.shape[0], train=False)
        return self.engine.decode_indices(w, latents.contiguous()).clone()

    def configure_optimizers_and_checkpointers(self):
        """ae.py:224-244 + vqvae.py:108-124: optimizers for netE / netG, checkpointers netE / netG / netC."""
        from ...engine.checkpoint import Checkpointer
        o, c = [], []
        for net, name, node in ((self.encoder, "netE", self.cfg.MODEL.ENCODER), (self.generator, "netG", self.cfg.MODEL.GENERATOR)):
            opt = build_optimizer(net, self.cfg, suffix="_G")
            o.append({"optimizer": opt, "scheduler": build_lr_scheduler(self.cfg, opt), "type": "generator"})
            os.makedirs(os.path.join(self.cfg.OUTPUT_DIR, name), exist_ok=True)
            c.append({"checkpointer": Checkpointer(net, os.path.join(self.cfg.OUTPUT_DIR, name)), "pretrained": node.WEIGHTS})
        os.makedirs(os.path.join(self.cfg.OUTPUT_DIR, "netC"), exist_ok=True)
        c.append({"checkpointer": Checkpointer(self.codebook, os.path.join(self.cfg.OUTPUT_DIR, "netC")),
                  "pretrained": self.cfg.MODEL.CODEBOOK.WEIGHTS})
        return o, c
