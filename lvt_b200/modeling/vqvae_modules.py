"""Encoder / generator / codebook modules with the reference's registry + from_config surface
(vidgen/modeling/encoder/{build,resencoder}.py, generator/{build,resdecoder}.py, vq/vq_embedding.py).
They are parameter trees over ONE shared VQVAEEngine (flat buffers); the arithmetic lives in the engine."""
import torch
from torch import nn

from .. import ops
from ..utils.registry import Registry
from .param_tree import ParamTree, attach_store

ENCODER_REGISTRY = Registry("ENCODER")
GENERATOR_REGISTRY = Registry("GENERATOR")


class Encoder(ParamTree):
    pass


class Generator(ParamTree):
    pass


def build_encoder(cfg, **kwargs):
    enc = ENCODER_REGISTRY.get(cfg.MODEL.ENCODER.NAME).from_config(cfg, **kwargs)
    assert isinstance(enc, Encoder)
    return enc


def build_generator(cfg, **kwargs):
    gen = GENERATOR_REGISTRY.get(cfg.MODEL.GENERATOR.NAME).from_config(cfg, **kwargs)
    assert isinstance(gen, Generator)
    return gen


def _check_res_cfg(node, what):
    if node.NORM != "" or node.SPECTRAL:
        raise NotImplementedError(f"{what}: NORM / SPECTRAL are empty/False in every shipped config")


@ENCODER_REGISTRY.register()
class ResEncoder(Encoder):
    """resencoder.py:24-76 (stride 4): parameters `layers.{0,2,4}.*`, `layers.{5+i}.block.{1,3}.*`."""

    @classmethod
    def from_config(cls, cfg, **kwargs):
        _check_res_cfg(cfg.MODEL.ENCODER, "ResEncoder")
        return cls(kwargs["engine"])

    def __init__(self, engine):
        super().__init__()
        object.__setattr__(self, "engine", engine)
        attach_store(self, engine.store, prefix="E.", strip="E.")


@GENERATOR_REGISTRY.register()
class ResDecoder(Generator):
    """resdecoder.py:24-75 (stride 4)."""

    @classmethod
    def from_config(cls, cfg, **kwargs):
        _check_res_cfg(cfg.MODEL.GENERATOR, "ResDecoder")
        return cls(kwargs["engine"])

    def __init__(self, engine):
        super().__init__()
        object.__setattr__(self, "engine", engine)
        attach_store(self, engine.store, prefix="G.", strip="G.")


class VQEmbedding(nn.Module):
    """vq_embedding.py:9-66: `embedding.weight` (K, D), buffers running_size (K), running_sum (K, D) — views of
    the engine's stacked codebook state."""

    def __init__(self, engine, index, ema=True, standalone=False):
        super().__init__()
        self.K, self.ema = engine.spec.K, ema
        self.embedding = nn.Embedding(engine.spec.K, engine.spec.D, _weight=engine.codebook[index])
        self.embedding.weight.requires_grad_(False)
        self.register_buffer("running_size", engine.running_size[index])
        self.register_buffer("running_sum", engine.running_sum[index])
        if standalone:  # CODEBOOK.NUM == 1 (vqvae.py:26-27): this module IS the model's codebook
            object.__setattr__(self, "engine", engine)
            with torch.no_grad():  # vq_embedding.py:12-21
                engine.codebook.uniform_(-1.0 / engine.spec.K, 1.0 / engine.spec.K)
                engine.running_sum.copy_(engine.codebook)
                engine.running_size.zero_()

    def forward(self, z_e_x, mode=""):
        """vq_embedding.py:23-34, modes "" (indices (n, h, w)) and "emb" (codes -> NHWC vectors)."""
        eng = self.engine
        if mode == "":
            return ops.vq_argmin(z_e_x.contiguous().float(), eng.codebook)[:, 0]
        if mode == "emb":
            return ops.vq_gather(z_e_x.unsqueeze(1).contiguous(), eng.codebook).permute(0, 2, 3, 1)
        raise ValueError("mode 'st' runs inside VQVAEEngine.forward_train (EMA + straight-through)")


class DVQEmbedding(nn.Module):
    """vq_embedding.py:69-99: modes "" (indices) and "emb" (codes -> vectors) run on the codebook kernels."""

    def __init__(self, engine, ema=True):
        super().__init__()
        object.__setattr__(self, "engine", engine)
        s = engine.spec
        self.num, self.D = s.num, s.num * s.D
        self.ve = nn.ModuleList([VQEmbedding(engine, i, ema) for i in range(s.num)])
        with torch.no_grad():  # VQEmbedding.__init__: uniform(-1/K, 1/K); running_sum starts as a copy
            engine.codebook.uniform_(-1.0 / s.K, 1.0 / s.K)
            engine.running_sum.copy_(engine.codebook)
            engine.running_size.zero_()

    def forward(self, z_e_x, mode=""):
        eng = self.engine
        if mode == "":
            assert z_e_x.dim() == 4
            return ops.vq_argmin(z_e_x.contiguous().float(), eng.codebook)            # (n, num, h, w) int64
        if mode == "emb":
            out = ops.vq_gather(z_e_x.contiguous(), eng.codebook)                      # (n, num*D, h, w)
            return out.permute(0, 2, 3, 1)                                            # reference returns NHWC
        raise ValueError("mode 'st' runs inside VQVAEEngine.forward_train (EMA + straight-through)")
