"""ORACLE (test infrastructure only) — torch-CPU fp32 functional restatement of the reference's
hot path.  Every function cites the reference file:line it follows (paths relative to the
reference repo root, package `vidgen`).  Weights are passed as plain dicts keyed exactly like the
reference's state_dict(), so the same tensors drive the reference (tests/golden/make_golden.py),
this oracle and the CUDA path.

Pinned by: tests/golden/*.npz, generated in the authoring container by importing the UNMODIFIED
reference through oracle/ref_shim.py (tests/test_oracle_golden.py checks this file against them
on every CPU run; tests/test_oracle_vs_reference.py re-checks against the live reference when
/root/reference is present).
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ============================================================================================
# configuration records (values come from the reference's YAML configs; see lvt_b200/config)
# ============================================================================================
@dataclass
class VTConfig:
    """MODEL.AUTOREGRESSIVE.VT.* (config/defaults.py:36-53, configs/vt/DSFVT.yaml:11-24)."""
    nc: int = 4
    nv: int = 512
    kernel: Tuple[int, int, int] = (7, 1, 1)
    stride: Tuple[int, int, int] = (16, 1, 1)
    de: int = 128
    d: int = 512
    da: int = 128
    blocks_e: Sequence[Tuple[int, int, int]] = tuple([(1, 16, 16)] * 8)
    heads_e: Sequence[int] = tuple([8] * 8)
    blocks_d: Sequence[Tuple[int, int, int]] = tuple([(1, 16, 16)] * 8)
    heads_d: Sequence[int] = tuple([8] * 8)
    n_prime: int = 1
    pad_value: int = -1
    ignore_index: int = -100
    video_shape: Tuple[int, int, int] = (16, 16, 16)  # (T, H, W) of the latent video
    share_p: bool = False  # SHARE_P: one output Linear for all channels (videotransformer.py:121-123; every shipped config: False)
    share_embeddings: bool = False  # SHARE_EMBEDDINGS: P: d -> de, logits against ch_embedder[k] (videotransformer.py:124-125,152-154)
    class_num: int = 0  # CLASS_NUM: class embedding concatenated before the encoder's projector (videotransformer.py:29-33,54-57)

    @property
    def slice_shape(self):
        return tuple(v // s for v, s in zip(self.video_shape, self.stride))


@dataclass
class VQVAEConfig:
    """configs/vqvae/PR-DVQVAE2.yaml + Base-VQVAE.yaml (K-DVQVAE: n_layers=4)."""
    in_channels: int = 3
    nf: int = 256
    res_channels: int = 128
    n_layers: int = 2
    codebook_num: int = 4
    codebook_size: int = 512
    codebook_dim: int = 256
    beta: float = 1.0
    ema_decay: float = 0.99
    ema_eps: float = 1e-5
    pixel_lambda: float = 1.0
    ema: bool = True  # MODEL.CODEBOOK.EMA (every shipped config: True); False trains the codebook by gradient


# ============================================================================================
# VQ-VAE: encoder / decoder / codebook
# ============================================================================================
def res_block(x: Tensor, w1, b1, w2, b2) -> Tensor:
    """ResBlock.forward (encoder/resencoder.py:10-21, generator/resdecoder.py:10-21).
    The block starts with an in-place ReLU, so the skip connection carries relu(x):
    out = relu(x) + conv1x1(relu(conv3x3(relu(x))))."""
    r = torch.relu(x)
    h = F.conv2d(r, w1, b1, stride=1, padding=1)
    h = F.conv2d(torch.relu(h), w2, b2)
    return r + h


def res_encoder(x: Tensor, sd: Dict[str, Tensor], n_layers: int) -> Tensor:
    """ResEncoder.forward, stride 4 (encoder/resencoder.py:46-52,60-76); sd keys `layers.N.*`."""
    h = torch.relu(F.conv2d(x, sd["layers.0.weight"], sd["layers.0.bias"], stride=2, padding=1))
    h = torch.relu(F.conv2d(h, sd["layers.2.weight"], sd["layers.2.bias"], stride=2, padding=1))
    h = F.conv2d(h, sd["layers.4.weight"], sd["layers.4.bias"], stride=1, padding=1)
    for i in range(n_layers):
        p = f"layers.{5 + i}.block."
        h = res_block(h, sd[p + "1.weight"], sd[p + "1.bias"], sd[p + "3.weight"], sd[p + "3.bias"])
    return h


def res_decoder(z: Tensor, sd: Dict[str, Tensor], n_layers: int, out_activation="tanh") -> Tensor:
    """ResDecoder.forward, stride 4 (generator/resdecoder.py:48-57,66-75)."""
    h = F.conv2d(z, sd["layers.0.weight"], sd["layers.0.bias"], stride=1, padding=1)
    for i in range(n_layers):
        p = f"layers.{1 + i}.block."
        h = res_block(h, sd[p + "1.weight"], sd[p + "1.bias"], sd[p + "3.weight"], sd[p + "3.bias"])
    k = 1 + n_layers
    h = torch.relu(h)
    h = F.conv_transpose2d(h, sd[f"layers.{k + 1}.weight"], sd[f"layers.{k + 1}.bias"], stride=2, padding=1)
    h = torch.relu(h)
    h = F.conv_transpose2d(h, sd[f"layers.{k + 3}.weight"], sd[f"layers.{k + 3}.bias"], stride=2, padding=1)
    if out_activation == "tanh":
        h = torch.tanh(h)
    elif out_activation == "sigmoid":
        h = torch.sigmoid(h)
    return h


def vq_indices(x_nhwc: Tensor, codebook: Tensor) -> Tensor:
    """VectorQuantization.forward (vq/vq_utils.py:7-24): expanded-form fp32 distance, first min."""
    flat = x_nhwc.reshape(-1, codebook.size(1))
    c2 = torch.sum(codebook ** 2, dim=1)
    x2 = torch.sum(flat ** 2, dim=1, keepdim=True)
    dist = torch.addmm(c2 + x2, flat, codebook.t(), alpha=-2.0, beta=1.0)
    return torch.min(dist, dim=1)[1].view(*x_nhwc.shape[:-1])


def dvq_indices(z_e: Tensor, codebooks: Tensor) -> Tensor:
    """DVQEmbedding.forward(mode="") (vq/vq_embedding.py:77-82, 23-27): [n,num*D,h,w] -> [n,num,h,w]."""
    num, K, D = codebooks.shape
    parts = z_e.split(D, dim=1)
    return torch.stack([vq_indices(p.permute(0, 2, 3, 1).contiguous(), codebooks[g])
                        for g, p in enumerate(parts)], dim=1)


def dvq_embed(idx: Tensor, codebooks: Tensor) -> Tensor:
    """DVQEmbedding.forward(mode="emb") + the permute of VQVAEModel.decode
    (vq/vq_embedding.py:92-97, meta_arch/vqvae.py:104): [n,num,h,w] -> [n,num*D,h,w]."""
    outs = [codebooks[g][idx[:, g]] for g in range(codebooks.shape[0])]  # n,h,w,D each
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous()


def dvq_straight_through(z_e: Tensor, codebooks: Tensor, running_size: Tensor, running_sum: Tensor,
                         cfg: VQVAEConfig, world_counts=None, world_sums=None):
    """DVQEmbedding "st" mode with EMA (vq/vq_embedding.py:34-66,83-91), GPU semantics
    (running_sum is NOT aliased to the codebook; SURVEY parity trap 2).
    Returns (z_q_st value, z_q_bar, idx, new_codebooks, new_running_size, new_running_sum).
    z_q_st is gathered from the PRE-update codebook, z_q_bar from the POST-update one.
    world_counts/world_sums: statistics already summed over ranks (AllReduce, :46-47,53-54)."""
    num, K, D = codebooks.shape
    idx = dvq_indices(z_e, codebooks)
    zq_st = dvq_embed(idx, codebooks)
    new_cb, new_rs, new_rsum = [], [], []
    for g in range(num):
        ind = idx[:, g].reshape(-1)
        x = z_e[:, g * D:(g + 1) * D].permute(0, 2, 3, 1).reshape(-1, D)
        size = torch.zeros(K).index_add_(0, ind, torch.ones(ind.numel()))
        s = torch.zeros(K, D).index_add_(0, ind, x.float())
        if world_counts is not None:
            size, s = world_counts[g], world_sums[g]
        rs = running_size[g] * cfg.ema_decay + (1 - cfg.ema_decay) * size
        rsum = running_sum[g] * cfg.ema_decay + (1 - cfg.ema_decay) * s
        n = rs.sum()
        size_ = (rs + cfg.ema_eps) / (n + K * cfg.ema_eps) * n
        new_cb.append(rsum / size_.unsqueeze(1))
        new_rs.append(rs)
        new_rsum.append(rsum)
    new_cb = torch.stack(new_cb)
    zq_bar = dvq_embed(idx, new_cb)
    return zq_st, zq_bar, idx, new_cb, torch.stack(new_rs), torch.stack(new_rsum)


def vqvae_inference(x01: Tensor, sdE, sdG, codebooks, cfg: VQVAEConfig):
    """AutoEncoderModel.forward(mode='inference') for VQVAEModel (meta_arch/ae.py:120-147,
    151-168; vqvae.py:93-106): x01 in [0,1] -> (reconstruction in [0,1], latent int64)."""
    x = (x01 - 0.5) / 0.5
    idx = dvq_indices(res_encoder(x, sdE, cfg.n_layers), codebooks)
    out = res_decoder(dvq_embed(idx, codebooks), sdG, cfg.n_layers)
    return (out * 0.5 + 0.5).clamp_(0.0, 1.0), idx


def vqvae_supervised_loss(x01: Tensor, sdE, sdG, codebooks, running_size, running_sum,
                          cfg: VQVAEConfig):
    """VQVAEModel.compute_supervised_loss with EMA codebook (meta_arch/vqvae.py:66-91).
    sdE/sdG tensors may require grad; the straight-through estimator passes d/dz_q to z_e
    (vq_utils.py:50-53).  Returns (loss dict, aux dict)."""
    x = (x01 - 0.5) / 0.5
    z_e = res_encoder(x, sdE, cfg.n_layers)
    if not cfg.ema:
        # MODEL.CODEBOOK.EMA False (vq_embedding.py:36-38,61-66; vqvae.py:84-88): vq_st sees the DETACHED weight, so the
        # codebook (pass it with requires_grad) gets its gradient only through index_select under the extra loss
        # mse(z_q, sg[z_e]), which the reference stores under the key 'loss_dict'
        with torch.no_grad():
            idx = dvq_indices(z_e.detach(), codebooks)
        zq_bar = dvq_embed(idx, codebooks)
        z_st = z_e + (zq_bar.detach() - z_e).detach()
        x_tilde = res_decoder(z_st, sdG, cfg.n_layers)
        losses = {
            "loss_reconstruction": cfg.pixel_lambda * F.mse_loss(x_tilde, x),
            "loss_dict": F.mse_loss(zq_bar, z_e.detach()),
            "loss_commitment": cfg.beta * F.mse_loss(z_e, zq_bar.detach()),
        }
        return losses, dict(z_e=z_e, idx=idx, x_tilde=x_tilde, codebooks=codebooks, running_size=running_size,
                            running_sum=running_sum)
    with torch.no_grad():
        zq_st, zq_bar, idx, new_cb, new_rs, new_rsum = dvq_straight_through(
            z_e.detach(), codebooks, running_size, running_sum, cfg)
    z_st = z_e + (zq_st - z_e).detach()  # value = zq_st, gradient = identity to z_e
    x_tilde = res_decoder(z_st, sdG, cfg.n_layers)
    losses = {
        "loss_reconstruction": cfg.pixel_lambda * F.mse_loss(x_tilde, x),
        "loss_commitment": cfg.beta * F.mse_loss(z_e, zq_bar),
    }
    aux = dict(z_e=z_e, idx=idx, x_tilde=x_tilde, codebooks=new_cb, running_size=new_rs,
               running_sum=new_rsum)
    return losses, aux


# ============================================================================================
# DSFVT: subscale slicing helpers (host-side input contract)
# ============================================================================================
def subscale_order(st, sh, sw):
    """vt_utils.py:6-14: raster order of the (a,b,c) slice offsets."""
    idx2abc = [(a, b, c) for a in range(st) for b in range(sh) for c in range(sw)]
    return idx2abc, {abc: i for i, abc in enumerate(idx2abc)}


def slice_mask(a, b, c, st, sh, sw, T, H, W, dtype=torch.bool):
    """vt_utils.py:24-33 (vectorised): 1 at positions belonging to slice (a,b,c)."""
    m = torch.zeros(1, 1, T, H, W, dtype=dtype)
    m[0, 0, a::st, b::sh, c::sw] = 1
    return m


def visible_abc_mask(a, b, c, st, sh, sw, T, H, W, dtype=torch.bool):
    """vt_utils.py:48-57: union of all slices that precede (a,b,c) in subscale order."""
    idx2abc, abc2idx = subscale_order(st, sh, sw)
    m = torch.zeros(1, 1, T, H, W, dtype=torch.int64)
    for (ai, bi, ci) in idx2abc[:abc2idx[(a, b, c)]]:
        m += slice_mask(ai, bi, ci, st, sh, sw, T, H, W, dtype=torch.int64)
    return m.to(dtype)


def ss_shift(x, a, b, c, st, sh, sw, T, H, W, kt, kh, kw, pad_value=0):
    """vt_utils.py:104-128: crop/pad so that a VALID conv of kernel (kt,kh,kw), stride (st,sh,sw)
    is centred on the elements of slice (a,b,c)."""
    def axis(off, size, s, k):
        n = size // s
        lo, hi = off, off + (n - 1) * s
        front, back = k // 2 - lo, k // 2 - (size - hi - 1)
        return max(0, -front), max(0, -back), max(0, front), max(0, back)
    ct0, ct1, pt0, pt1 = axis(a, T, st, kt)
    ch0, ch1, ph0, ph1 = axis(b, H, sh, kh)
    cw0, cw1, pw0, pw1 = axis(c, W, sw, kw)
    x = x[:, :, ct0:T - ct1, ch0:H - ch1, cw0:W - cw1]
    return F.pad(x, [pw0, pw1, ph0, ph1, pt0, pt1], mode="constant", value=pad_value)


def prepare_slice(video: Tensor, abc, cfg: VTConfig):
    """DatasetMapper slice construction (data/dataset_mapper.py:113-149) for a chosen (a,b,c).
    video: (T, nc, H, W) int64 -> dict(context, slice, slice_idx, ignore_mask)."""
    st, sh, sw = cfg.stride
    v = video[None].transpose(1, 2)  # 1, nc, T, H, W
    _, nc, T, H, W = v.shape
    t, h, w = T // st, H // sh, W // sw
    a, b, c = abc
    _, abc2idx = subscale_order(st, sh, sw)
    sm = slice_mask(a, b, c, st, sh, sw, T, H, W)
    sl = v.masked_select(sm).clone().view(1, nc, t, h, w)
    vm = visible_abc_mask(a, b, c, st, sh, sw, T, H, W)
    ctx = ss_shift(v.masked_fill(~vm, cfg.pad_value), a, b, c, st, sh, sw, T, H, W, *cfg.kernel,
                   pad_value=cfg.pad_value)
    ig = torch.zeros(1, 1, T, H, W, dtype=torch.bool)
    if cfg.n_prime > 0:
        ig[:, :, :cfg.n_prime] = True
    ig = ig.masked_select(sm).clone().view(1, 1, t, h, w)
    return {"context": ctx[0].long(), "slice": sl[0].long(),
            "slice_idx": torch.tensor(abc2idx[(a, b, c)]).long(), "ignore_mask": ig[0]}


def sample_abc(rng, cfg: VTConfig):
    """data/dataset_mapper.py:123-127: `rng` is a random.Random (the mapper uses the global one)."""
    st, sh, sw = cfg.stride
    t = cfg.video_shape[0] // st
    single = (t == 1 and sh == 1 and sw == 1)
    a = rng.randint(cfg.n_prime, st - 1) if single else rng.randint(0, st - 1)
    return a, rng.randint(0, sh - 1), rng.randint(0, sw - 1)


# ============================================================================================
# DSFVT: network
# ============================================================================================
def positional_encoding_table(d_model: int, shape, min_ts=1.0, max_ts=1.0e4) -> Tensor:
    """PositionalEncoding (vt_attention.py:10-50) as an additive table [d_model, t, h, w].
    d//6 timescales; channel ranges [0,2n) t, [2n,4n) h, [4n,6n) w; remaining channels untouched."""
    n = d_model // 6
    inc = np.log(max_ts / min_ts) / n
    inv = min_ts * torch.exp(torch.arange(n).float() * -inc)
    tab = torch.zeros(d_model, *shape)
    for dim in range(3):
        pos = torch.arange(shape[dim], dtype=torch.float)
        st = pos.view(-1, 1) * inv.view(1, -1)
        sig = torch.cat([torch.sin(st), torch.cos(st)], 1).T  # [2n, L]
        view = [2 * n, 1, 1, 1]
        view[1 + dim] = shape[dim]
        tab[dim * 2 * n:(dim + 1) * 2 * n] += sig.reshape(view)
    return tab


def relpos_bias(dt_bank, dh_bank, dw_bank, block) -> Tensor:
    """BlockLocalAttention.get_B (vt_attention.py:146-174): B[h,i,j] = dt[h,ti-tj+t-1] +
    dh[h,hi-hj+h-1] + dw[h,wi-wj+w-1] -> [heads, 1, L, L]."""
    t, h, w = block
    L = t * h * w
    i = torch.arange(L)
    ti, hi, wi = i // (h * w), (i // w) % h, i % w
    B = (dt_bank[:, ti[:, None] - ti[None, :] + t - 1] + dh_bank[:, hi[:, None] - hi[None, :] + h - 1]
         + dw_bank[:, wi[:, None] - wi[None, :] + w - 1])
    return B.unsqueeze(1)


def multi_head_attention(x: Tensor, p: Dict[str, Tensor], B: Tensor, causal: bool) -> Tensor:
    """MultiHeadAttention.forward + ScaledDotProductAttention.forward
    (vt_attention.py:114-129, 61-81). x: (b, L, d). p keys: layer_norm.{weight,bias}, w_q, w_k,
    w_v (heads, d, da), proj.weight (d, heads*da)."""
    b, L, d = x.shape
    na, _, da = p["w_q"].shape
    xn = F.layer_norm(x, (d,), p["layer_norm.weight"], p["layer_norm.bias"])
    xe = xn.reshape(1, b * L, d).expand(na, b * L, d)
    q = torch.bmm(xe, p["w_q"]).view(na, b, L, da)
    k = torch.bmm(xe, p["w_k"]).view(na, b, L, da)
    v = torch.bmm(xe, p["w_v"]).view(na, b, L, da)
    attn = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(da) + B
    if causal:
        attn = attn.masked_fill(torch.triu(torch.ones(L, L), diagonal=1).bool(), -1e4)
    attn = torch.softmax(attn, dim=3)
    out = torch.matmul(attn, v)  # na, b, L, da
    out = out.permute(1, 2, 0, 3).reshape(b, L, na * da)  # head-major concat (:125-126)
    return F.linear(out, p["proj.weight"]) + x


def block_local_attention(x: Tensor, p: Dict[str, Tensor], block, causal: bool) -> Tensor:
    """BlockLocalAttention.forward (vt_attention.py:176-202), both the block==slice fast path and
    the general tiled path.  x: (B, C, T, H, W)."""
    Bn, C, T, H, W = x.shape
    t, h, w = block
    nt, nh, nw = T // t, H // h, W // w
    # (B, C, nt, t, nh, h, nw, w) -> (B*nt*nh*nw, t*h*w, C)
    xb = x.view(Bn, C, nt, t, nh, h, nw, w).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(-1, t * h * w, C)
    bias = relpos_bias(p["dt_bank"], p["dh_bank"], p["dw_bank"], block)
    mha = {k[4:]: v for k, v in p.items() if k.startswith("mha.")}
    y = multi_head_attention(xb, mha, bias, causal)
    f = F.layer_norm(y, (C,), p["ffn.0.weight"], p["ffn.0.bias"])
    f = F.linear(torch.relu(F.linear(f, p["ffn.1.weight"], p["ffn.1.bias"])), p["ffn.3.weight"], p["ffn.3.bias"])
    y = f + y
    y = y.view(Bn, nt, nh, nw, t, h, w, C).permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(Bn, C, T, H, W)
    return y


def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def vt_encoder(context: Tensor, slice_idx: Tensor, sd, cfg: VTConfig, class_idx: Tensor = None) -> Tensor:
    """VTEncoder.forward (videotransformer.py:35-59). context (b, nc, Tc, Hc, Wc) int64 with
    pad_value entries; the one-hot of a padded entry is all-zero (:41-48).  positional_encoder
    exists but is never applied (:18)."""
    pad = context == cfg.pad_value
    oh = F.one_hot(context.masked_fill(pad, 0), cfg.nv).masked_fill(pad.unsqueeze(-1), 0)
    b, nc, Tc, Hc, Wc, nv = oh.shape
    xin = oh.permute(0, 1, 5, 2, 3, 4).reshape(b, nc * nv, Tc, Hc, Wc).float()
    x = F.conv3d(xin, sd["encoder.conv.weight"], sd["encoder.conv.bias"], stride=tuple(cfg.stride))
    x = x + sd["encoder.slice_embedding.weight"][slice_idx][:, :, None, None, None]
    if cfg.class_num > 0 and class_idx is not None:  # videotransformer.py:54-57
        x = torch.cat([x, sd["encoder.class_embedding.weight"][class_idx][:, :, None, None, None].expand_as(x)], dim=1)
    x = F.conv3d(x, sd["encoder.linear_projector.weight"])
    for i, blk in enumerate(cfg.blocks_e):
        x = block_local_attention(x, _sub(sd, f"encoder.block_local_attention.{i}."), blk, causal=False)
    return x


def masked_conv3d(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """MaskedConv3d.forward (vt_utils.py:183-200): causal pad (w: k//2 both sides, h and t: k-1 in
    front), taps [:, :, -1, -1, kw//2:] zeroed, VALID conv."""
    kt, kh, kw = weight.shape[2:]
    if kw // 2 > 0:
        # like the reference, the masked taps are zeroed IN PLACE in the parameter's data (so
        # autograd still reports a non-zero gradient for them; they are re-zeroed before every
        # use, hence never influence an output).  The CUDA path keeps them at zero instead.
        with torch.no_grad():
            weight[:, :, -1, -1, kw // 2:] = 0
    return F.conv3d(F.pad(x, [kw // 2, kw // 2, kh - 1, 0, kt - 1, 0]), weight, bias)


def vt_decoder(slc: Tensor, zl: Tensor, sd, cfg: VTConfig) -> Tensor:
    """VTDecoder.forward (videotransformer.py:80-101)."""
    emb = sum(sd[f"decoder.ch_embedder.{k}.weight"][slc[:, k]] for k in range(cfg.nc))  # b,t,h,w,de
    x = masked_conv3d(emb.permute(0, 4, 1, 2, 3), sd["decoder.conv.conv.weight"], sd["decoder.conv.conv.bias"])
    x = x + positional_encoding_table(cfg.d, x.shape[2:])[None]
    x = x + F.conv3d(zl, sd["decoder.linear_projector.weight"])
    for i, blk in enumerate(cfg.blocks_d):
        x = block_local_attention(x, _sub(sd, f"decoder.block_local_attention.{i}."), blk, causal=True)
    return x


def channel_predictor_logits(slc: Tensor, yl: Tensor, sd, cfg: VTConfig) -> List[Tensor]:
    """ChannelPredictor.forward(mode="logits"), SHARE_P False (videotransformer.py:138-160):
    channel k sees LN(y) and the one-hot codes of channels < k."""
    b, d, t, h, w = yl.shape
    y = F.layer_norm(yl.view(b, d, -1).transpose(1, 2), (d,), sd["ch_predictor.layer_norm.weight"],
                     sd["ch_predictor.layer_norm.bias"])
    oh = F.one_hot(slc.view(b, cfg.nc, -1).transpose(1, 2), cfg.nv).view(b, t * h * w, -1).float()
    outs = []
    for k in range(cfg.nc):
        inp = y if k == 0 else torch.cat((y, oh[:, :, :k * cfg.nv]), dim=2)
        u = torch.relu(F.linear(inp, sd[f"ch_predictor.U.{k}.weight"], sd[f"ch_predictor.U.{k}.bias"]))
        pk = "ch_predictor.P" if (cfg.share_p or cfg.share_embeddings) else f"ch_predictor.P.{k}"  # videotransformer.py:150-155
        o = F.linear(u, sd[pk + ".weight"], sd[pk + ".bias"])
        if cfg.share_embeddings:
            o = F.linear(o, sd[f"decoder.ch_embedder.{k}.weight"])
        outs.append(o.transpose(1, 2).reshape(b, cfg.nv, t, h, w))
    return outs


def vt_logits(context, slc, slice_idx, sd, cfg: VTConfig, class_idx=None) -> List[Tensor]:
    """VideoTransformer.forward(mode="logits") (videotransformer.py:232-239)."""
    zl = vt_encoder(context, slice_idx, sd, cfg, class_idx)
    return channel_predictor_logits(slc, vt_decoder(slc, zl, sd, cfg), sd, cfg)


def vt_supervised_loss(context, slc, slice_idx, ignore_mask, sd, cfg: VTConfig, class_idx=None) -> Tensor:
    """VideoTransformerModel.compute_supervised_loss (meta_arch/vt.py:301-314)."""
    target = slc.masked_fill(ignore_mask, cfg.ignore_index)
    pred = vt_logits(context, slc, slice_idx, sd, cfg, class_idx)
    loss = sum(F.cross_entropy(pred[k], target[:, k], ignore_index=cfg.ignore_index) for k in range(cfg.nc))
    return loss / cfg.nc


def stack_batch(samples: List[Dict[str, Tensor]]):
    """VideoTransformerModel.preprocess_data (meta_arch/vt.py:284-299)."""
    return tuple(torch.stack([s[k] for s in samples], 0) for k in ("context", "slice", "slice_idx", "ignore_mask"))


# ============================================================================================
# optimizers (torch.optim semantics the reference configures; solver/build.py:62-72)
# ============================================================================================
def rmsprop_step(p, g, sq, buf, lr, alpha=0.95, momentum=0.9, eps=1e-8):
    """torch.optim.RMSprop (centered=False, weight_decay=0): DSFVT.yaml:28-32."""
    sq.mul_(alpha).addcmul_(g, g, value=1 - alpha)
    avg = sq.sqrt().add_(eps)
    buf.mul_(momentum).addcdiv_(g, avg)
    p.add_(buf, alpha=-lr)


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.9, eps=1e-8):
    """torch.optim.Adam (weight_decay=0, amsgrad=False): config/defaults.py:113-114 betas (.9,.9)."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# ============================================================================================
# deterministic synthetic weights / inputs shared by reference, oracle and CUDA path
# ============================================================================================
def dsfvt_param_shapes(cfg: VTConfig) -> Dict[str, Tuple[int, ...]]:
    """Parameter names/shapes of VideoTransformer (videotransformer.py:11-33,62-78,104-137,
    vt_attention.py:98-104,132-144); identical to the reference state_dict (buffers excluded)."""
    s = {}
    kt, kh, kw = cfg.kernel
    s["encoder.conv.weight"] = (cfg.de, cfg.nc * cfg.nv, kt, kh, kw)
    s["encoder.conv.bias"] = (cfg.de,)
    s["encoder.slice_embedding.weight"] = (cfg.stride[0] * cfg.stride[1] * cfg.stride[2], cfg.de)
    if cfg.class_num:
        s["encoder.class_embedding.weight"] = (cfg.class_num, cfg.de)
    s["encoder.linear_projector.weight"] = (cfg.d, cfg.de * (2 if cfg.class_num else 1), 1, 1, 1)

    def bla(prefix, block, heads):
        t, h, w = block
        s[prefix + "dt_bank"] = (heads, 2 * t - 1)
        s[prefix + "dh_bank"] = (heads, 2 * h - 1)
        s[prefix + "dw_bank"] = (heads, 2 * w - 1)
        for n in ("w_q", "w_k", "w_v"):
            s[prefix + "mha." + n] = (heads, cfg.d, cfg.da)
        s[prefix + "mha.layer_norm.weight"] = (cfg.d,)
        s[prefix + "mha.layer_norm.bias"] = (cfg.d,)
        s[prefix + "mha.proj.weight"] = (cfg.d, heads * cfg.da)
        s[prefix + "ffn.0.weight"] = (cfg.d,)
        s[prefix + "ffn.0.bias"] = (cfg.d,)
        s[prefix + "ffn.1.weight"] = (cfg.d, cfg.d)
        s[prefix + "ffn.1.bias"] = (cfg.d,)
        s[prefix + "ffn.3.weight"] = (cfg.d, cfg.d)
        s[prefix + "ffn.3.bias"] = (cfg.d,)

    for i, (blk, nh) in enumerate(zip(cfg.blocks_e, cfg.heads_e)):
        bla(f"encoder.block_local_attention.{i}.", blk, nh)
    for k in range(cfg.nc):
        s[f"decoder.ch_embedder.{k}.weight"] = (cfg.nv, cfg.de)
    s["decoder.conv.conv.weight"] = (cfg.d, cfg.de, 3, 3, 3)
    s["decoder.conv.conv.bias"] = (cfg.d,)
    s["decoder.linear_projector.weight"] = (cfg.d, cfg.d, 1, 1, 1)
    for i, (blk, nh) in enumerate(zip(cfg.blocks_d, cfg.heads_d)):
        bla(f"decoder.block_local_attention.{i}.", blk, nh)
    s["ch_predictor.layer_norm.weight"] = (cfg.d,)
    s["ch_predictor.layer_norm.bias"] = (cfg.d,)
    for k in range(cfg.nc):
        s[f"ch_predictor.U.{k}.weight"] = (cfg.d, cfg.d + k * cfg.nv)
        s[f"ch_predictor.U.{k}.bias"] = (cfg.d,)
    shared = cfg.share_p or cfg.share_embeddings
    for k in range(1 if shared else cfg.nc):
        pk = "ch_predictor.P" if shared else f"ch_predictor.P.{k}"
        s[pk + ".weight"] = (cfg.de if cfg.share_embeddings else cfg.nv, cfg.d)
        s[pk + ".bias"] = (cfg.de if cfg.share_embeddings else cfg.nv,)
    return s


def vqvae_param_shapes(cfg: VQVAEConfig):
    """netE / netG parameter shapes (resencoder.py:46-62, resdecoder.py:48-57)."""
    nf, rc = cfg.nf, cfg.res_channels
    e = {"layers.0.weight": (nf // 2, cfg.in_channels, 4, 4), "layers.0.bias": (nf // 2,),
         "layers.2.weight": (nf, nf // 2, 4, 4), "layers.2.bias": (nf,),
         "layers.4.weight": (nf, nf, 3, 3), "layers.4.bias": (nf,)}
    g = {"layers.0.weight": (nf, cfg.codebook_dim, 3, 3), "layers.0.bias": (nf,)}
    for i in range(cfg.n_layers):
        for d, base in ((e, 5), (g, 1)):
            p = f"layers.{base + i}.block."
            d[p + "1.weight"], d[p + "1.bias"] = (rc, nf, 3, 3), (rc,)
            d[p + "3.weight"], d[p + "3.bias"] = (nf, rc, 1, 1), (nf,)
    k = 1 + cfg.n_layers
    g[f"layers.{k + 1}.weight"], g[f"layers.{k + 1}.bias"] = (nf, nf // 2, 4, 4), (nf // 2,)
    g[f"layers.{k + 3}.weight"], g[f"layers.{k + 3}.bias"] = (nf // 2, cfg.in_channels, 4, 4), (cfg.in_channels,)
    return e, g


def synth_weights(shapes: Dict[str, Tuple[int, ...]], seed: int, bias_scale=0.02, bank_scale=0.5):
    """Seeded synthetic weights (numpy PCG64, independent of torch's RNG streams): fan-in scaled
    normals for matrices/filters, LayerNorm weights around 1, and NON-zero biases and
    relative-position banks so that every term of the path is exercised (the reference
    initialises banks to zero, vt_attention.py:142-144; SURVEY parity trap 6)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for name, shp in shapes.items():
        if name.endswith("_bank"):
            a = rng.standard_normal(shp) * bank_scale
        elif "layer_norm.weight" in name or name.endswith("ffn.0.weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shp)
        elif len(shp) == 1:
            a = rng.standard_normal(shp) * bias_scale
        elif "embed" in name:
            a = rng.standard_normal(shp) * 0.5
        elif ".w_" in name:  # (heads, d, da): fan-in is d
            a = rng.standard_normal(shp) / math.sqrt(shp[1])
        elif name == "encoder.conv.weight":  # one-hot input: <= nc*kt active rows per position
            a = rng.standard_normal(shp) * 0.2
        else:
            fan_in = int(np.prod(shp[1:]))
            a = rng.standard_normal(shp) / math.sqrt(fan_in)
        out[name] = torch.from_numpy(a.astype(np.float32))
    return out


def synth_latent_video(seed: int, cfg: VTConfig) -> Tensor:
    """(T, nc, H, W) int64 codes, np.random.RandomState(seed).randint."""
    T, H, W = cfg.video_shape
    return torch.from_numpy(np.random.RandomState(seed).randint(0, cfg.nv, (T, cfg.nc, H, W)).astype(np.int64))


def synth_vt_batch(batch: int, seed: int, cfg: VTConfig):
    """`batch` training samples exactly as the mapper would hand them to the model."""
    import random
    rng = random.Random(seed)
    samples = [prepare_slice(synth_latent_video(seed * 1000 + i, cfg), sample_abc(rng, cfg), cfg)
               for i in range(batch)]
    return stack_batch(samples)
