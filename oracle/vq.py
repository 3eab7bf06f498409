"""ORACLE (test infrastructure): ctypes front-end of vq_oracle.c plus the literal torch
restatement of vidgen/modeling/vq/vq_utils.py:7-24 used to cross-check it."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvq_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.vq_argmin_oracle.restype = None
        _lib.vq_argmin_oracle.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 5
    return _lib


def vq_argmin_c(z_e, codebook, want_dist=False):
    """z_e [n, num*D, h, w] fp32 (NCHW), codebook [num, K, D] -> idx [n, num, h, w] int64."""
    z = np.ascontiguousarray(z_e.detach().cpu().numpy(), dtype=np.float32)
    cb = np.ascontiguousarray(codebook.detach().cpu().numpy(), dtype=np.float32)
    n, c, h, w = z.shape
    num, K, D = cb.shape
    assert c == num * D and D % 8 == 0
    idx = np.empty((n, num, h, w), dtype=np.int64)
    dist = np.empty((n, num, h, w, K), dtype=np.float32) if want_dist else None
    _load().vq_argmin_oracle(z.ctypes.data, cb.ctypes.data, idx.ctypes.data,
                             dist.ctypes.data if want_dist else None, n, num, K, D, h * w)
    if want_dist:
        return torch.from_numpy(idx), torch.from_numpy(dist)
    return torch.from_numpy(idx)


def vq_torch(inputs, codebook):
    """Literal restatement of VectorQuantization.forward (vq_utils.py:7-24). inputs [..., D]."""
    embedding_size = codebook.size(1)
    inputs_size = inputs.size()
    inputs_flatten = inputs.reshape(-1, embedding_size)
    codebook_sqr = torch.sum(codebook ** 2, dim=1)
    inputs_sqr = torch.sum(inputs_flatten ** 2, dim=1, keepdim=True)
    distances = torch.addmm(codebook_sqr + inputs_sqr, inputs_flatten, codebook.t(),
                            alpha=-2.0, beta=1.0)
    _, indices_flatten = torch.min(distances, dim=1)
    return indices_flatten.view(*inputs_size[:-1]), distances


def dvq_argmin_torch(z_e, codebook):
    """DVQEmbedding.forward(mode="") (vq_embedding.py:77-82): z_e NCHW -> [n, num, h, w]."""
    num, K, D = codebook.shape
    out = []
    for i, part in enumerate(z_e.split(D, dim=1)):
        x = part.permute(0, 2, 3, 1).contiguous()
        out.append(vq_torch(x, codebook[i])[0])
    return torch.stack(out, dim=1)
