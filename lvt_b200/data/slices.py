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


def prepare_slices_batched(videos, abc, kernel, stride, n_prime, pad_value=-1):
    """Batched, device-side form of `prepare_slices` (dataset_mapper.py:113-149): one gather per output tensor
    instead of per-sample Python loops, on whatever device `videos` lives (a latent dataset resident in HBM needs no
    DataLoader workers).  videos (B, T, nc, H, W) integer codes, abc (B, 3) integer slice offsets ->
    context (B, nc, Tc, Hc, Wc) int64, slice (B, nc, t, h, w) int64, slice_idx (B,) int64, ignore_mask (B, 1, t, h, w) bool,
    identical to stacking prepare_slices over the batch.

    Derivation of the context window (ss_shift, vt_utils.py:104-128): along an axis of size S with stride s, kernel k and
    offset o the window starts at source index o - k//2 and has (S//s - 1)*s + k entries; an entry is visible iff its
    source position lies inside the video and belongs to a slice generated before (a, b, c) in raster order
    (visible_abc_mask, vt_utils.py:48-57), otherwise it holds pad_value."""
    v = torch.as_tensor(videos)
    dev = v.device
    abc = torch.as_tensor(abc, device=dev).long()
    B, T, nc, H, W = v.shape
    (st, sh, sw), (kt, kh, kw) = stride, kernel
    assert T % st == 0 and H % sh == 0 and W % sw == 0
    t, h, w = T // st, H // sh, W // sw
    a, b, c = abc[:, 0], abc[:, 1], abc[:, 2]
    ar = lambda n: torch.arange(n, device=dev)  # noqa: E731
    # ---- slice: video[bi, a + ts*st, :, b + hs*sh, c + ws*sw]
    ti = (a[:, None] + ar(t)[None] * st)                      # (B, t)
    hi = (b[:, None] + ar(h)[None] * sh)
    wi = (c[:, None] + ar(w)[None] * sw)
    bi = ar(B)[:, None, None, None]
    slc = v[bi, ti[:, :, None, None], :, hi[:, None, :, None], wi[:, None, None, :]]      # (B, t, h, w, nc)
    slc = slc.permute(0, 4, 1, 2, 3).contiguous().long()
    # ---- context window
    Tc, Hc, Wc = (t - 1) * st + kt, (h - 1) * sh + kh, (w - 1) * sw + kw
    tau = a[:, None] - kt // 2 + ar(Tc)[None]                 # (B, Tc) source indices, may be out of range
    eta = b[:, None] - kh // 2 + ar(Hc)[None]
    omg = c[:, None] - kw // 2 + ar(Wc)[None]
    inside = ((tau >= 0) & (tau < T))[:, :, None, None] & ((eta >= 0) & (eta < H))[:, None, :, None] & \
             ((omg >= 0) & (omg < W))[:, None, None, :]
    src_idx = ((tau % st)[:, :, None, None] * sh + (eta % sh)[:, None, :, None]) * sw + (omg % sw)[:, None, None, :]
    slice_idx = (a * sh + b) * sw + c
    visible = inside & (src_idx < slice_idx[:, None, None, None])
    g = v[bi, tau.clamp(0, T - 1)[:, :, None, None], :, eta.clamp(0, H - 1)[:, None, :, None],
          omg.clamp(0, W - 1)[:, None, None, :]]                                           # (B, Tc, Hc, Wc, nc)
    ctx = torch.where(visible[..., None], g.long(), torch.full((), pad_value, dtype=torch.long, device=dev))
    ctx = ctx.permute(0, 4, 1, 2, 3).contiguous()
    # ---- ignore mask: primed frames of the slice
    ig = (ti < n_prime)[:, None, :, None, None].expand(B, 1, t, h, w).contiguous() if n_prime > 0 else \
        torch.zeros((B, 1, t, h, w), dtype=torch.bool, device=dev)
    return {"context": ctx, "slice": slc, "slice_idx": slice_idx, "ignore_mask": ig}
