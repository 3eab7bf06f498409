"""Host-side construction of DSFVT training samples (the input contract of the VT path).

Mirrors the slice block of the reference DatasetMapper (vidgen/data/dataset_mapper.py:113-149)
and its helpers (vidgen/modeling/autoregressive/vt_utils.py:6-57,104-128) with strided slicing
instead of Python triple loops.  Pure integer bookkeeping on the host, as in the reference
(DataLoader workers); the device only ever sees the resulting int64 tensors.
"""
import random as _random

import numpy as np
import torch
import torch.nn.functional as F


def subscale_order(st, sh, sw):
    """Raster order of slice offsets (a, b, c) (vt_utils.py:6-14)."""
    idx2abc = [(a, b, c) for a in range(st) for b in range(sh) for c in range(sw)]
    return idx2abc, {abc: i for i, abc in enumerate(idx2abc)}


def slice_mask(a, b, c, st, sh, sw, T, H, W, device=torch.device("cpu"), dtype=torch.float):
    """1 at the positions of slice (a, b, c) (vt_utils.py:24-33)."""
    m = torch.zeros(1, 1, T, H, W, device=device, dtype=dtype)
    m[0, 0, a::st, b::sh, c::sw] = 1
    return m


def visible_abc_mask(a, b, c, st, sh, sw, T, H, W, device=torch.device("cpu"), dtype=torch.float):
    """1 at the positions of every slice generated before (a, b, c) (vt_utils.py:48-57)."""
    idx2abc, abc2idx = subscale_order(st, sh, sw)
    m = torch.zeros(1, 1, T, H, W, device=device, dtype=torch.int32)
    for (ai, bi, ci) in idx2abc[:abc2idx[(a, b, c)]]:
        m[0, 0, ai::st, bi::sh, ci::sw] += 1
    return m.to(dtype)


def ss_shift(x, a, b, c, st, sh, sw, T, H, W, kt, kh, kw, pad_value=0):
    """Crop / pad so that a VALID strided conv is centred on slice (a, b, c) (vt_utils.py:104-128)."""
    crops, pads = [], []
    for off, size, s, k in ((a, T, st, kt), (b, H, sh, kh), (c, W, sw, kw)):
        n = size // s
        lo, hi = off, off + (n - 1) * s
        front, back = k // 2 - lo, k // 2 - (size - hi - 1)
        crops.append((max(0, -front), size - max(0, -back)))
        pads.append((max(0, front), max(0, back)))
    x = x[:, :, crops[0][0]:crops[0][1], crops[1][0]:crops[1][1], crops[2][0]:crops[2][1]]
    return F.pad(x, [pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]],
                 mode="constant", value=pad_value)


def sample_abc(stride, n_frames, n_prime, rng=_random):
    """Slice choice of the mapper (dataset_mapper.py:122-127)."""
    st, sh, sw = stride
    single = (n_frames // st == 1 and sh == 1 and sw == 1)
    a = rng.randint(n_prime, st - 1) if single else rng.randint(0, st - 1)
    return a, rng.randint(0, sh - 1), rng.randint(0, sw - 1)


def prepare_slices(video, abc, kernel, stride, n_prime, pad_value=-1):
    """video (T, nc, H, W) integer codes -> dict(context, slice, slice_idx, ignore_mask) exactly
    as the reference mapper emits them (dataset_mapper.py:113-149)."""
    st, sh, sw = stride
    v = torch.as_tensor(video)[None].transpose(1, 2)  # 1, nc, T, H, W
    _, nc, T, H, W = v.shape
    assert T % st == 0 and H % sh == 0 and W % sw == 0
    t, h, w = T // st, H // sh, W // sw
    a, b, c = abc
    _, abc2idx = subscale_order(st, sh, sw)
    slc = v[:, :, a::st, b::sh, c::sw].clone()
    vm = visible_abc_mask(a, b, c, st, sh, sw, T, H, W, dtype=torch.bool)
    ctx = ss_shift(v.masked_fill(~vm, pad_value), a, b, c, st, sh, sw, T, H, W, *kernel, pad_value=pad_value)
    ig = torch.zeros(1, 1, T, H, W, dtype=torch.bool)
    if n_prime > 0:
        ig[:, :, :n_prime] = True
    ig = ig[:, :, a::st, b::sh, c::sw].clone()
    return {"context": ctx[0].long(), "slice": slc[0].long(),
            "slice_idx": torch.tensor(abc2idx[(a, b, c)]).long(), "ignore_mask": ig[0]}


def synthetic_latent_video(seed, shape=(16, 4, 16, 16), nv=512):
    """(T, nc, H, W) int64 codes."""
    return torch.from_numpy(np.random.RandomState(seed).randint(0, nv, shape).astype(np.int64))


def synthetic_vt_batch(batch, seed, kernel=(7, 1, 1), stride=(16, 1, 1), n_prime=1, video_shape=(16, 4, 16, 16),
                       nv=512, pad_value=-1):
    """`batch` mapper-format samples stacked like VideoTransformerModel.preprocess_data
    (meta_arch/vt.py:284-299)."""
    rng = _random.Random(seed)
    samples = [prepare_slices(synthetic_latent_video(seed * 1000 + i, video_shape, nv),
                              sample_abc(stride, video_shape[0], n_prime, rng), kernel, stride, n_prime, pad_value)
               for i in range(batch)]
    return tuple(torch.stack([s[k] for s in samples], 0) for k in ("context", "slice", "slice_idx", "ignore_mask"))
