"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, imported
through oracle/ref_shim.py) on seeded synthetic weights and inputs.  Run in the authoring
container only:   python tests/golden/make_golden.py

The fixtures are small (sub-sampled outputs + checksums); weights and inputs are NOT stored —
they are regenerated from seeds by oracle.lvt_oracle.synth_* on every machine.
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import lvt_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def load_into(module, weights):
    sd = module.state_dict()
    missing = [k for k in sd if k not in weights and not k.endswith(("dt", "dh", "dw", "mask", "inv_timescales"))]
    assert not missing, missing
    module.load_state_dict({**sd, **weights}, strict=True)


def small_vt_cfg(layers, share_p=False, video_shape=(16, 16, 16), share_embeddings=False, class_num=0):
    return O.VTConfig(blocks_e=tuple([(1, 16, 16)] * layers), heads_e=tuple([8] * layers),
                      blocks_d=tuple([(1, 16, 16)] * layers), heads_d=tuple([8] * layers), share_p=share_p,
                      video_shape=video_shape, share_embeddings=share_embeddings, class_num=class_num)


def golden_dsfvt():
    """DSFVT (configs/vt/DSFVT.yaml) with 2+2 layers (same modules, fewer repeats) and the full
    8+8 network: loss, sub-sampled logits and gradient checksums."""
    ref_shim.install()
    from vidgen.modeling.meta_arch import build_model
    from vidgen.utils.events import EventStorage
    which = os.environ.get("LVT_GOLDEN_DSFVT", "dsfvt_l2,dsfvt_full,dsfvt_l2_sharep,dsfvt_l2_tiled,dsfvt_l2_shareemb,dsfvt_l2_class").split(",")
    # dsfvt_l2_sharep: MODEL.AUTOREGRESSIVE.VT.SHARE_P True, the reference's config default (config/defaults.py),
    # which every shipped YAML overrides with False
    # dsfvt_l2_tiled: a 32-frame latent video => slices of (2, 16, 16) over (1, 16, 16) attention blocks: the general
    # tiled path of BlockLocalAttention.forward (vt_attention.py:189-200), which no shipped config reaches
    for tag, layers, batch, share_p in (("dsfvt_l2", 2, 3, False), ("dsfvt_full", 8, 2, False), ("dsfvt_l2_sharep", 2, 3, True),
                                        ("dsfvt_l2_tiled", 2, 2, False), ("dsfvt_l2_shareemb", 2, 3, False),
                                        ("dsfvt_l2_class", 2, 4, False)):
        if tag not in which:
            continue
        vshape = (32, 16, 16) if tag.endswith("tiled") else (16, 16, 16)
        share_emb = tag.endswith("shareemb")  # SHARE_EMBEDDINGS: P: d -> de, logits against the channel's embedding table
        class_num = 5 if tag.endswith("class") else 0  # CLASS_NUM: class-conditioned encoder (two samples share a class)
        blocks = str(tuple([(1, 16, 16)] * layers))
        heads = str(tuple([8] * layers))
        cfg = ref_shim.reference_cfg("configs/vt/DSFVT.yaml", [
            "MODEL.AUTOREGRESSIVE.VT.BLOCKS_E", blocks, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_E", heads,
            "MODEL.AUTOREGRESSIVE.VT.BLOCKS_D", blocks, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_D", heads,
            "MODEL.AUTOREGRESSIVE.VT.SHARE_P", share_p, "MODEL.AUTOREGRESSIVE.VT.SHARE_EMBEDDINGS", share_emb,
            "MODEL.AUTOREGRESSIVE.VT.CLASS_NUM", class_num])
        torch.manual_seed(0)
        model = build_model(cfg)
        ocfg = small_vt_cfg(layers, share_p, vshape, share_emb, class_num)
        weights = O.synth_weights(O.dsfvt_param_shapes(ocfg), seed=1234)
        load_into(model.model, weights)
        model.train()
        context, slc, slice_idx, ignore = O.synth_vt_batch(batch, seed=77, cfg=ocfg)
        data = [{"context": context[i], "slice": slc[i], "slice_idx": slice_idx[i], "ignore_mask": ignore[i]}
                for i in range(batch)]
        class_idx = None
        if class_num:
            class_idx = torch.tensor([3, 0, 3, 4][:batch])
            for i in range(batch):
                data[i]["class"] = class_idx[i]
        with EventStorage(0):
            loss = model(data, mode="supervised")["loss_cross_entropy"]
        loss.backward()
        with torch.no_grad():
            logits = model.model(context, slc, slice_idx, class_idx=class_idx)  # list nc x (b, nv, t, h, w)
        grads = {k: p.grad for k, p in model.model.named_parameters()}
        # the reference's own data path must agree with the oracle's restatement of the mapper
        fix = {"loss": loss.detach().numpy(),
               "logits_sub": torch.stack(logits)[:, :, ::7, 0, ::3, ::5].numpy(),
               "logits_sum": torch.stack(logits).double().sum().numpy(),
               "slice_idx": slice_idx.numpy(), "context_sum": context.sum().numpy()}
        for k in (("encoder.class_embedding.weight", "encoder.linear_projector.weight") if class_num else ()) + \
                 ("encoder.conv.weight", "encoder.slice_embedding.weight", "decoder.conv.conv.weight",
                  "decoder.ch_embedder.1.weight", "ch_predictor.U.2.weight",
                  "ch_predictor.P.bias" if (share_p or share_emb) else "ch_predictor.P.3.bias",
                  "ch_predictor.P.weight" if (share_p or share_emb) else "ch_predictor.P.0.weight",
                  f"decoder.block_local_attention.{layers - 1}.dh_bank",
                  "encoder.block_local_attention.0.mha.w_k", "encoder.block_local_attention.0.ffn.1.weight",
                  "decoder.block_local_attention.0.mha.proj.weight", "decoder.linear_projector.weight"):
            g = grads[k]
            fix["gnorm:" + k] = g.double().norm().numpy()
            fix["gsub:" + k] = g.reshape(-1)[::max(1, g.numel() // 64)][:64].numpy()
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), **fix)
        print(tag, "loss", float(loss))


def golden_mapper():
    """The reference DatasetMapper's slice construction for all four VT configs."""
    ref_shim.install()
    from vidgen.data.dataset_mapper import DatasetMapper
    import vidgen.data.dataset_mapper as dm
    out = {}
    for name in ("DSFVT", "DSSVT", "DSTSVT"):
        cfg = ref_shim.reference_cfg(f"configs/vt/{name}.yaml")
        vt = cfg.MODEL.AUTOREGRESSIVE.VT
        T = cfg.INPUT.N_FRAMES_PER_VIDEO_TRAIN
        ocfg = O.VTConfig(kernel=tuple(vt.KERNEL), stride=tuple(vt.STRIDE), n_prime=vt.N_PRIME,
                          video_shape=(T, 16, 16))
        mapper = DatasetMapper(cfg, True)
        for i in range(3):
            video = O.synth_latent_video(500 + i, ocfg)
            random.seed(900 + i)
            # feed the mapper an in-memory video: bypass its np.load of latent files
            d = {"image_sequence": video.numpy().astype("float32")}
            d = _run_mapper_slices(mapper, dm, d)
            for k in ("context", "slice", "slice_idx", "ignore_mask"):
                out[f"{name}:{i}:{k}"] = d[k].numpy()
    np.savez_compressed(os.path.join(OUT, "mapper.npz"), **out)
    print("mapper ok")


def _run_mapper_slices(mapper, dm, dataset_dict):
    """Executes exactly the `if self.prepare_slices:` block of DatasetMapper.__call__
    (data/dataset_mapper.py:113-149) by calling the unmodified method on a dict that already
    carries "image_sequence" (the file-reading branches above it are skipped by key)."""
    import inspect
    src = inspect.getsource(type(mapper).__call__)
    assert "prepare_slices" in src
    # The method reads files unless none of its dataset keys is present; with only
    # "image_sequence" given every loader branch is skipped and the slice block runs.
    out = mapper(dataset_dict)
    assert out is not None
    return out


def golden_vqvae():
    """PR-DVQVAE2 (configs/vqvae/PR-DVQVAE2.yaml): inference latents (bit-exact target),
    reconstruction, and one supervised step's losses/grad checksums, GPU (un-aliased) semantics."""
    ref_shim.install()
    from vidgen.modeling.meta_arch import build_model
    cfg = ref_shim.reference_cfg("configs/vqvae/PR-DVQVAE2.yaml")
    torch.manual_seed(0)
    model = build_model(cfg)
    ocfg = O.VQVAEConfig()
    eshape, gshape = O.vqvae_param_shapes(ocfg)
    we, wg = O.synth_weights(eshape, seed=11), O.synth_weights(gshape, seed=12)
    load_into(model.encoder, we)
    load_into(model.generator, wg)
    gen = torch.Generator().manual_seed(1234)
    x = torch.rand((8, 3, 64, 64), generator=gen)
    fix = {}
    for init in ("default", "spread"):
        gcb = torch.Generator().manual_seed(5)
        if init == "default":
            cb = (torch.rand((4, 512, 64), generator=gcb) * 2 - 1) / 512
        else:
            with torch.no_grad():
                z_std = model.encoder((x - 0.5) / 0.5).std()
            cb = torch.randn((4, 512, 64), generator=gcb) * z_std
            fix["spread_std"] = z_std.numpy()
        for g in range(4):
            ve = model.codebook.ve[g]
            ve.embedding.weight.data.copy_(cb[g])
            # GPU semantics: running_sum is a separate buffer (SURVEY parity trap 2)
            ve.running_sum = cb[g].clone()
            ve.running_size.zero_()
        model.eval()
        with torch.no_grad():
            out = model([{"image": x[i]} for i in range(x.shape[0])])
        fix[f"{init}:latent"] = torch.stack([o["latent"] for o in out]).numpy()
        fix[f"{init}:recon_sub"] = torch.stack([o["reconstruction"] for o in out])[:, :, ::5, ::7].numpy()
        model.train()
        model.zero_grad()
        losses = model([{"image": x[i]} for i in range(x.shape[0])], mode="supervised")
        sum(losses.values()).backward()
        fix[f"{init}:loss_reconstruction"] = losses["loss_reconstruction"].detach().numpy()
        fix[f"{init}:loss_commitment"] = losses["loss_commitment"].detach().numpy()
        fix[f"{init}:codebook_after_sub"] = torch.stack(
            [model.codebook.ve[g].embedding.weight.data for g in range(4)])[:, ::16, ::8].numpy()
        fix[f"{init}:running_size"] = torch.stack([model.codebook.ve[g].running_size for g in range(4)]).numpy()
        for k in ("layers.0.weight", "layers.4.weight", "layers.6.block.3.weight"):
            fix[f"{init}:gE:{k}"] = dict(model.encoder.named_parameters())[k].grad.double().norm().numpy()
        for k in ("layers.0.weight", "layers.4.weight", "layers.6.weight"):
            fix[f"{init}:gG:{k}"] = dict(model.generator.named_parameters())[k].grad.double().norm().numpy()
    np.savez_compressed(os.path.join(OUT, "vqvae.npz"), **fix)
    print("vqvae ok")


def golden_vq_only():
    """vq() on raw vectors: indices + full fp32 distance rows for a few vectors (bit-level pin of
    oracle/vq_oracle.c)."""
    ref_shim.install()
    from vidgen.modeling.vq.vq_utils import vq
    g = torch.Generator().manual_seed(3)
    fix = {}
    for D in (64, 256):
        x = torch.randn((1024, D), generator=g) * 0.3
        cb = torch.randn((512, D), generator=g) * 0.3
        idx = vq(x, cb)
        c2 = torch.sum(cb ** 2, dim=1)
        x2 = torch.sum(x ** 2, dim=1, keepdim=True)
        dist = torch.addmm(c2 + x2, x, cb.t(), alpha=-2.0, beta=1.0)
        fix[f"D{D}:idx"] = idx.numpy()
        fix[f"D{D}:dist_rows"] = dist[:4].numpy()
    np.savez_compressed(os.path.join(OUT, "vq.npz"), **fix)
    print("vq ok")


SUBSCALE = {"DSSVT": ((1, 3, 3), (1, 2, 2), (4, 16, 16)), "DSTSVT": ((5, 3, 3), (4, 2, 2), (16, 16, 16))}


def golden_subscale():
    """configs/vt/DSSVT.yaml and DSTSVT.yaml with 2+2 layers (spatially strided one-hot conv, t > 1 masked
    conv, (4,8,8) blocks with a three-axis bias, partially true ignore mask): loss, sub-sampled logits and
    gradient checksums -- pins the oracle paths the DSFVT fixtures do not reach."""
    ref_shim.install()
    from vidgen.modeling.meta_arch import build_model
    from vidgen.utils.events import EventStorage
    layers, batch = 2, 2
    blocks = str(tuple([(4, 8, 8)] * layers))
    heads = str(tuple([8] * layers))
    for name, (kernel, stride, vshape) in SUBSCALE.items():
        cfg = ref_shim.reference_cfg(f"configs/vt/{name}.yaml", [
            "MODEL.AUTOREGRESSIVE.VT.BLOCKS_E", blocks, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_E", heads,
            "MODEL.AUTOREGRESSIVE.VT.BLOCKS_D", blocks, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_D", heads])
        torch.manual_seed(0)
        model = build_model(cfg)
        ocfg = O.VTConfig(kernel=kernel, stride=stride, video_shape=vshape, blocks_e=tuple([(4, 8, 8)] * layers),
                          heads_e=tuple([8] * layers), blocks_d=tuple([(4, 8, 8)] * layers), heads_d=tuple([8] * layers))
        weights = O.synth_weights(O.dsfvt_param_shapes(ocfg), seed=4321)
        load_into(model.model, weights)
        model.train()
        context, slc, slice_idx, ignore = O.synth_vt_batch(batch, seed=5, cfg=ocfg)
        data = [{"context": context[i], "slice": slc[i], "slice_idx": slice_idx[i], "ignore_mask": ignore[i]}
                for i in range(batch)]
        with EventStorage(0):
            loss = model(data, mode="supervised")["loss_cross_entropy"]
        loss.backward()
        with torch.no_grad():
            logits = torch.stack(model.model(context, slc, slice_idx))  # nc, b, nv, t, h, w
        grads = {k: p.grad for k, p in model.model.named_parameters()}
        fix = {"loss": loss.detach().numpy(), "logits_sub": logits[:, :, ::7, :, ::3, ::5].numpy(),
               "logits_sum": logits.double().sum().numpy(), "slice_idx": slice_idx.numpy(),
               "context_sum": context.sum().numpy(), "ignore_sum": ignore.sum().numpy()}
        for k in ("encoder.conv.weight", "encoder.slice_embedding.weight", "decoder.conv.conv.weight",
                  "decoder.ch_embedder.2.weight", "ch_predictor.U.3.weight",
                  "decoder.block_local_attention.1.dt_bank", "decoder.block_local_attention.0.dh_bank",
                  "encoder.block_local_attention.1.dw_bank", "encoder.block_local_attention.0.mha.w_q",
                  "decoder.block_local_attention.1.ffn.3.weight", "decoder.linear_projector.weight"):
            g = grads[k]
            fix["gnorm:" + k] = g.double().norm().numpy()
            fix["gsub:" + k] = g.reshape(-1)[::max(1, g.numel() // 64)][:64].numpy()
        np.savez_compressed(os.path.join(OUT, name.lower() + "_l2.npz"), **fix)
        print(name, "loss", float(loss))


def golden_vqvae_noema():
    """PR-DVQVAE2 with MODEL.CODEBOOK.EMA False (no shipped config; vq_embedding.py:36-38,61-66, vqvae.py:84-88):
    the three losses and the gradients of one supervised step, codebook included."""
    ref_shim.install()
    from vidgen.modeling.meta_arch import build_model
    cfg = ref_shim.reference_cfg("configs/vqvae/PR-DVQVAE2.yaml", ["MODEL.CODEBOOK.EMA", False])
    torch.manual_seed(0)
    model = build_model(cfg)
    ocfg = O.VQVAEConfig(ema=False)
    eshape, gshape = O.vqvae_param_shapes(ocfg)
    we, wg = O.synth_weights(eshape, seed=11), O.synth_weights(gshape, seed=12)
    load_into(model.encoder, we)
    load_into(model.generator, wg)
    x = torch.rand((8, 3, 64, 64), generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        z_std = model.encoder((x - 0.5) / 0.5).std()
    cb = torch.randn((4, 512, 64), generator=torch.Generator().manual_seed(5)) * z_std
    for g in range(4):
        model.codebook.ve[g].embedding.weight.data.copy_(cb[g])
    model.train()
    model.zero_grad()
    losses = model([{"image": x[i]} for i in range(x.shape[0])], mode="supervised")
    sum(losses.values()).backward()
    fix = {"spread_std": z_std.numpy()}
    for k, v in losses.items():
        fix["loss:" + k] = v.detach().numpy()
    gcb = torch.stack([model.codebook.ve[g].embedding.weight.grad for g in range(4)])
    fix["gcb_norm"] = gcb.double().norm().numpy()
    fix["gcb_sub"] = gcb[:, ::16, ::8].numpy()
    for k in ("layers.0.weight", "layers.4.weight", "layers.6.block.3.weight"):
        fix[f"gE:{k}"] = dict(model.encoder.named_parameters())[k].grad.double().norm().numpy()
    for k in ("layers.0.weight", "layers.4.weight", "layers.6.weight"):
        fix[f"gG:{k}"] = dict(model.generator.named_parameters())[k].grad.double().norm().numpy()
    np.savez_compressed(os.path.join(OUT, "vqvae_noema.npz"), **fix)
    print("vqvae_noema ok", {k: float(v) for k, v in losses.items()})


def golden_kdvqvae():
    """K-DVQVAE (configs/vqvae/K-DVQVAE.yaml, N_LAYERS 4): inference latents / reconstruction and one
    supervised step's losses (spread codebook)."""
    ref_shim.install()
    from vidgen.modeling.meta_arch import build_model
    cfg = ref_shim.reference_cfg("configs/vqvae/K-DVQVAE.yaml")
    torch.manual_seed(0)
    model = build_model(cfg)
    ocfg = O.VQVAEConfig(n_layers=4)
    eshape, gshape = O.vqvae_param_shapes(ocfg)
    we, wg = O.synth_weights(eshape, seed=21), O.synth_weights(gshape, seed=22)
    load_into(model.encoder, we)
    load_into(model.generator, wg)
    x = torch.rand((4, 3, 64, 64), generator=torch.Generator().manual_seed(4321))
    with torch.no_grad():
        z_std = model.encoder((x - 0.5) / 0.5).std()
    cb = torch.randn((4, 512, 64), generator=torch.Generator().manual_seed(6)) * z_std
    fix = {"spread_std": z_std.numpy()}
    for g in range(4):
        ve = model.codebook.ve[g]
        ve.embedding.weight.data.copy_(cb[g])
        ve.running_sum = cb[g].clone()
        ve.running_size.zero_()
    model.eval()
    with torch.no_grad():
        out = model([{"image": x[i]} for i in range(x.shape[0])])
    fix["latent"] = torch.stack([o["latent"] for o in out]).numpy()
    fix["recon_sub"] = torch.stack([o["reconstruction"] for o in out])[:, :, ::5, ::7].numpy()
    model.train()
    model.zero_grad()
    losses = model([{"image": x[i]} for i in range(x.shape[0])], mode="supervised")
    sum(losses.values()).backward()
    fix["loss_reconstruction"] = losses["loss_reconstruction"].detach().numpy()
    fix["loss_commitment"] = losses["loss_commitment"].detach().numpy()
    for k in ("layers.0.weight", "layers.8.block.1.weight"):
        fix[f"gE:{k}"] = dict(model.encoder.named_parameters())[k].grad.double().norm().numpy()
    for k in ("layers.0.weight", "layers.4.block.3.weight", "layers.8.weight"):
        fix[f"gG:{k}"] = dict(model.generator.named_parameters())[k].grad.double().norm().numpy()
    np.savez_compressed(os.path.join(OUT, "kdvqvae.npz"), **fix)
    print("kdvqvae ok")


if __name__ == "__main__":
    assert ref_shim.available(), "run in the authoring container (needs /root/reference)"
    which = sys.argv[1:] or ["vq", "vqvae", "vqvae_noema", "kdvqvae", "mapper", "dsfvt", "subscale"]
    for w in which:
        {"vq": golden_vq_only, "vqvae": golden_vqvae, "vqvae_noema": golden_vqvae_noema, "kdvqvae": golden_kdvqvae, "mapper": golden_mapper,
         "dsfvt": golden_dsfvt, "subscale": golden_subscale}[w]()
