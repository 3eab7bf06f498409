"""GPU: the reference-facing Python surface (registries, build_model, forward(data, mode), Trainer, sampling)
drives the same engines; state_dict keys equal the reference's."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
SMALL = ["MODEL.AUTOREGRESSIVE.VT.BLOCKS_E", ((1, 16, 16),) * 2, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_E", (8, 8),
         "MODEL.AUTOREGRESSIVE.VT.BLOCKS_D", ((1, 16, 16),) * 2, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_D", (8, 8)]


def test_vt_model_state_dict_and_logits_match_oracle(cuda_lib, tmp_path):
    from lvt_b200.config.presets import preset
    from lvt_b200.modeling import build_model
    from oracle import lvt_oracle as O
    cfg = preset("DSFVT", SMALL + ["OUTPUT_DIR", str(tmp_path)])
    cfg.freeze()
    model = build_model(cfg)
    ocfg = O.VTConfig(blocks_e=((1, 16, 16),) * 2, heads_e=(8, 8), blocks_d=((1, 16, 16),) * 2, heads_d=(8, 8))
    shapes = O.dsfvt_param_shapes(ocfg)
    sd = model.model.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items() if k in shapes} == {k: tuple(v) for k, v in shapes.items()}
    extra = set(sd) - set(shapes)   # the reference's buffers
    assert all(k.endswith((".dt", ".dh", ".dw", ".mask", "inv_timescales")) for k in extra), extra
    weights = O.synth_weights(shapes, seed=1234)
    model.model.load_state_dict(weights, strict=False)
    # BitsEvaluator path: teacher-forced logits of a whole video (meta_arch/vt.py:230-282)
    video = O.synth_latent_video(3, ocfg)
    cfg2 = preset("DSFVT", SMALL + ["TEST.EVALUATORS", "BitsEvaluator", "OUTPUT_DIR", str(tmp_path)])
    model.cfg = cfg2
    model.train(False)
    out = model([{"image_sequence": video}])[0]
    assert out["logits"].shape == (4, 512, 16, 16, 16) and out["ignore_mask"][0, 0].all() and not out["ignore_mask"][0, 1].any()
    # oracle logits for slice a = 5
    sample = O.prepare_slice(video, (5, 0, 0), ocfg)
    with torch.no_grad():
        want = torch.stack(O.vt_logits(sample["context"][None], sample["slice"][None], sample["slice_idx"][None], weights, ocfg))
    got = out["logits"][:, :, 5].cpu()            # (nc, nv, H, W)
    want = want[:, 0, :, 0]                       # (nc, nv, h, w)
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()


def test_trainer_runs_supervised_steps(cuda_lib, tmp_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import train_net
    from lvt_b200.config.presets import preset
    from lvt_b200.engine import Trainer
    for name, over in (("DSFVT", SMALL + ["SOLVER.IMS_PER_BATCH", 4]), ("PR-DVQVAE2", ["SOLVER.IMS_PER_BATCH", 8])):
        cfg = preset(name, over + ["SOLVER.MAX_ITER", 3, "SOLVER.CHECKPOINT_PERIOD", 0, "OUTPUT_DIR", str(tmp_path / name), "SEED", 1])
        cfg.freeze()
        tr = Trainer(cfg, data_loader=train_net.synthetic_loader(cfg))
        tr.train()
        hist = tr.storage.history("total_loss")
        assert len(hist) >= 1 and all(np.isfinite(v) for v, _ in hist)
        assert os.path.exists(tmp_path / name / ("netG" if name == "DSFVT" else "netE") / "model_final.pth")


@pytest.mark.parametrize("name", ["DSFVT", "PR-DVQVAE2"])
def test_trainer_graph_replay_matches_eager_launches(cuda_lib, tmp_path, monkeypatch, name):
    """Trainer fast path: model(data, 'supervised') replaying forward + backward as CUDA graphs gives the loss
    trajectory and the parameters of the eager launch sequence (same seeds, same batches; fp32 sums accumulated with
    atomics in both, hence 1e-5 rather than bit equality)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import train_net
    from lvt_b200.config.presets import preset
    from lvt_b200.engine import Trainer
    over = (SMALL + ["SOLVER.IMS_PER_BATCH", 4]) if name == "DSFVT" else ["SOLVER.IMS_PER_BATCH", 8]
    runs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("LVT_TRAINER_GRAPH", mode)
        cfg = preset(name, over + ["SOLVER.MAX_ITER", 4, "SOLVER.CHECKPOINT_PERIOD", 0,
                                   "OUTPUT_DIR", str(tmp_path / (name + mode)), "SEED", 3])
        cfg.freeze()
        torch.manual_seed(11)
        tr = Trainer(cfg, data_loader=train_net.synthetic_loader(cfg))
        assert getattr(tr.model, "_graphed") == (mode == "1")
        losses = []
        tr.model.train()
        from lvt_b200.utils.events import EventStorage
        eng = tr.model.model.engine if name == "DSFVT" else tr.model.engine
        grad1 = None
        with EventStorage(0) as tr.storage:
            for tr.iter in range(4):
                data = next(tr._iter)
                ld = tr.model(data, mode="supervised")
                sum(ld.values()).backward()
                losses.append(float(sum(v.detach() for v in ld.values())))
                if grad1 is None:
                    grad1 = eng.store.grad.clone()   # same parameters in both modes: the gradients must agree
                for o in tr.optimizers:
                    o["optimizer"].step()
                for o in tr.optimizers:
                    o["optimizer"].zero_grad()
        runs[mode] = (losses, grad1)
    la, lb = runs["0"][0], runs["1"][0]
    assert all(abs(a - b) <= 1e-4 * abs(a) for a, b in zip(la, lb)), (la, lb)
    # (parameters after several steps are not compared: RMSprop / Adam turn every noise-level gradient entry into a
    # full-size step of random sign, so they differ between ANY two runs with atomically accumulated sums)
    ga, gb = runs["0"][1], runs["1"][1]
    assert (ga - gb).abs().max().item() <= 1e-5 * ga.abs().max().item()


def test_vqvae_model_inference_and_sampling_roundtrip(cuda_lib, tmp_path):
    from lvt_b200.config.presets import preset
    from lvt_b200.modeling import build_model
    cfg = preset("PR-DVQVAE2", ["OUTPUT_DIR", str(tmp_path)])
    cfg.freeze()
    vq = build_model(cfg)
    vq.train(False)
    frames = torch.rand(5, 3, 64, 64)
    out = vq([{"image_sequence": frames}])[0]
    assert out["latent"].shape == (5, 4, 16, 16) and out["latent"].dtype == torch.int64
    assert out["reconstruction"].shape == (5, 3, 64, 64)
    assert torch.equal(vq.encode(frames.cuda()), out["latent"])
    dec = vq.decode(out["latent"])
    assert torch.allclose(vq.back_normalizer(dec).clamp_(0, 1), out["reconstruction"], atol=1e-6)
    # VT sampling of one frame after 15 primed ones (vt.py:81-136), 2+2 layers
    cfgv = preset("DSFVT", SMALL + ["TEST.EVALUATORS", "VTSampler", "TEST.VT_SAMPLER.N_PRIME", 15,
                                    "TEST.VT_SAMPLER.NUM_SAMPLES", 1, "OUTPUT_DIR", str(tmp_path)])
    cfgv.freeze()
    vt = build_model(cfgv)
    vt.train(False)
    seq = torch.randint(0, 512, (16, 4, 16, 16))
    torch.manual_seed(0)
    sample = vt([{"image_sequence": seq}])[0]["samples"][0]
    assert sample.shape == (4, 16, 16, 16)
    assert torch.equal(sample[:, :15].cpu(), seq.transpose(0, 1)[:, :15])
    assert int(sample.min()) >= 0 and int(sample.max()) < 512


def test_base_vqvae_single_codebook(cuda_lib, tmp_path):
    """configs/vqvae/Base-VQVAE.yaml (CODEBOOK.NUM 1: one 512 x 256 VQEmbedding, vqvae.py:26-27): latents are (n, h, w),
    bit-exact against the C oracle on the engine's own z_e (generic D = 256 search path); encode -> decode round trip;
    a supervised step runs through the Trainer surface; netC checkpoint keys are the reference's."""
    from oracle import vq as ovq
    from lvt_b200.config.presets import preset
    from lvt_b200.modeling import build_model
    cfg = preset("Base-VQVAE", ["OUTPUT_DIR", str(tmp_path)])
    cfg.freeze()
    torch.manual_seed(3)
    model = build_model(cfg)
    assert set(model.codebook.state_dict()) == {"embedding.weight", "running_size", "running_sum"}
    assert tuple(model.codebook.embedding.weight.shape) == (512, 256)
    with torch.no_grad():
        model.engine.codebook.normal_(0.0, 0.05)
    model.train(False)
    x = torch.rand((6, 3, 64, 64), generator=torch.Generator().manual_seed(1))
    out = model([{"image": x[i]} for i in range(6)])
    lat = torch.stack([o["latent"] for o in out])
    assert lat.shape == (6, 16, 16) and out[0]["reconstruction"].shape == (3, 64, 64)
    w = model.engine.workspace(6, train=False)
    z_e = w.z_e.cpu().view(6, 16, 16, 256).permute(0, 3, 1, 2).contiguous()   # the z_e the latents were taken from
    want = ovq.vq_argmin_c(z_e, model.engine.codebook.cpu())[:, 0]
    assert torch.equal(lat.cpu(), want)
    assert torch.equal(model.encode(x.cuda()).cpu(), lat.cpu())
    assert model.decode(lat).shape == (6, 3, 64, 64)
    model.train(True)
    from lvt_b200.utils.events import EventStorage
    with EventStorage(0):
        losses = model([{"image": x[i]} for i in range(6)], mode="supervised")
        sum(losses.values()).backward()
    assert all(torch.isfinite(v) for v in losses.values())
    assert model.engine.store.grad.abs().sum().item() > 0


def test_vqvae_model_without_codebook_ema(cuda_lib, tmp_path):
    """MODEL.CODEBOOK.EMA False through the reference surface (vqvae.py:84-88,108-116): a third loss under the key
    'loss_dict' (its value equals loss_commitment / beta), the EMA buffers stay put, and the optimizers the model
    configures move the codebook by at most lr on the first Adam step."""
    from lvt_b200.config.presets import preset
    from lvt_b200.modeling import build_model
    from lvt_b200.utils.events import EventStorage
    cfg = preset("PR-DVQVAE2", ["OUTPUT_DIR", str(tmp_path), "MODEL.CODEBOOK.EMA", False])
    cfg.freeze()
    torch.manual_seed(3)
    model = build_model(cfg)
    with torch.no_grad():
        model.engine.codebook.normal_(0.0, 0.05)
    optimizers, _ = model.configure_optimizers_and_checkpointers()
    model.train(True)
    x = torch.rand((4, 3, 64, 64), generator=torch.Generator().manual_seed(1))
    cb0, rs0 = model.engine.codebook.clone(), model.engine.running_size.clone()
    with EventStorage(0):
        losses = model([{"image": x[i]} for i in range(4)], mode="supervised")
        sum(losses.values()).backward()
    assert set(losses) == {"loss_reconstruction", "loss_commitment", "loss_dict"}
    assert abs(losses["loss_dict"].item() - losses["loss_commitment"].item() / cfg.MODEL.CODEBOOK.BETA) <= 1e-6
    assert torch.equal(model.engine.codebook, cb0) and model.engine.cb_grad.abs().sum().item() > 0
    for o in optimizers:
        o["optimizer"].step()
    moved = (model.engine.codebook - cb0).abs()
    assert 0 < moved.max().item() <= cfg.SOLVER.LR_G * 1.001
    assert torch.equal(model.engine.running_size, rs0)


def test_graph_sampler_matches_per_pixel_loop(cuda_lib, tmp_path):
    """VideoTransformer.sample_slice (one CUDA-graph replay per position) against the same fused per-position step
    launched eagerly and against the reference-shaped per-pixel loop (vt.py:107-134).  At temperature 1e-10 the
    multinomial draw IS the argmax of identical logits (every other probability underflows to exactly 0), so the three
    sampled frames must be identical code for code whatever random stream each path consumes.  (At 1e-4, as in round 1,
    a randomly initialised network that has collapsed to two alternating codes with logits ~1e-7 apart still leaves
    the draw to the noise, and torch's graph-captured generator does not replay the eager offsets: flaky.)"""
    from lvt_b200.config.presets import preset
    from lvt_b200.modeling import build_model
    cfgv = preset("DSFVT", SMALL + ["TEST.EVALUATORS", "VTSampler", "OUTPUT_DIR", str(tmp_path)])
    cfgv.freeze()
    vt = build_model(cfgv)
    vt.train(False)
    video = torch.randint(0, 512, (1, 4, 16, 16, 16)).cuda()
    video[:, :, 15:] = 0
    vt.model.sample_incremental = False   # the full-pass step: bitwise the logits of the per-pixel loop
    outs = {}
    for graph in (True, "eager", False):
        vt.sampler_graph = graph
        torch.manual_seed(0)
        outs[graph] = vt.sample_video(video.clone(), temp=1e-10, n_prime=15).cpu()
    assert torch.equal(outs[True], outs["eager"])
    assert torch.equal(outs[True], outs[False])
    assert torch.equal(outs[True][:, :, :15], video[:, :, :15].cpu())


def test_tiled_attention_sampling_and_loss_surface(cuda_lib, tmp_path):
    """A 32-frame latent video with the DSFVT network: slices of (2, 16, 16) over (1, 16, 16) attention blocks, i.e.
    the general tiled path of BlockLocalAttention.forward (vt_attention.py:189-200; parity of the train step is in
    test_dsfvt_gpu.py).  Through the model surface: the supervised loss is finite, and sample_video falls back to
    the full decoder pass per position (the K/V-cached row decoder assumes one block per slice), samples valid codes
    for the last frame and leaves the 31 priming frames untouched; graph replay == eager launches at temperature 1e-10."""
    from lvt_b200.config.presets import preset
    from lvt_b200.modeling import build_model
    cfgv = preset("DSFVT", SMALL + ["TEST.EVALUATORS", "VTSampler", "OUTPUT_DIR", str(tmp_path)])
    cfgv.freeze()
    vt = build_model(cfgv)
    vt.train(False)
    video = torch.randint(0, 512, (1, 4, 32, 16, 16)).cuda()
    video[:, :, 31:] = 0
    outs = {}
    for graph in (True, "eager"):
        vt.sampler_graph = graph
        torch.manual_seed(0)
        outs[graph] = vt.sample_video(video.clone(), temp=1e-10, n_prime=31).cpu()
    assert torch.equal(outs[True], outs["eager"])
    assert torch.equal(outs[True][:, :, :31], video[:, :, :31].cpu())
    assert int(outs[True].min()) >= 0 and int(outs[True].max()) < 512


def test_class_conditioned_model_surface(cuda_lib, tmp_path):
    """MODEL.AUTOREGRESSIVE.VT.CLASS_NUM > 0 through the model surface (videotransformer.py:29-33,54-57, meta_arch/vt.py:
    284-299): the "class" entry of the data dicts reaches the encoder, its embedding gets a gradient, the logits
    depend on the class, and a batch without classes is refused like the reference's shape error."""
    from lvt_b200 import _lib
    from lvt_b200.config.presets import preset
    from lvt_b200.modeling import build_model
    from lvt_b200.utils.events import EventStorage
    from oracle import lvt_oracle as O
    cfgv = preset("DSFVT", SMALL + ["MODEL.AUTOREGRESSIVE.VT.CLASS_NUM", 3, "OUTPUT_DIR", str(tmp_path)])
    cfgv.freeze()
    vt = build_model(cfgv)
    sd = vt.model.state_dict()
    assert tuple(sd["encoder.class_embedding.weight"].shape) == (3, 128)
    assert tuple(sd["encoder.linear_projector.weight"].shape) == (512, 256, 1, 1, 1)
    ocfg = O.VTConfig(blocks_e=((1, 16, 16),) * 2, heads_e=(8, 8), blocks_d=((1, 16, 16),) * 2, heads_d=(8, 8))
    context, slc, slice_idx, ignore = O.synth_vt_batch(2, seed=5, cfg=ocfg)
    data = [{"context": context[i], "slice": slc[i], "slice_idx": slice_idx[i], "ignore_mask": ignore[i],
             "class": torch.tensor(i + 1)} for i in range(2)]
    vt.train(True)
    with EventStorage(0):
        loss = vt(data, mode="supervised")["loss_cross_entropy"]
        loss.backward()
    g = vt.model.engine.store.g["encoder.class_embedding.weight"]
    assert torch.isfinite(loss) and g[0].abs().sum().item() == 0 and g[1].abs().sum().item() > 0 and g[2].abs().sum().item() > 0
    vt.train(False)
    with torch.no_grad():
        a = torch.stack(vt.model(context.cuda(), slc.cuda(), slice_idx.cuda(), class_idx=torch.tensor([0, 0]).cuda()))
        b = torch.stack(vt.model(context.cuda(), slc.cuda(), slice_idx.cuda(), class_idx=torch.tensor([0, 2]).cuda()))
    assert torch.equal(a[:, 0], b[:, 0]) and not torch.equal(a[:, 1], b[:, 1])
    with pytest.raises(_lib.LvtError):
        vt.model(context.cuda(), slc.cuda(), slice_idx.cuda())


def test_codes_extractor_round_trip(cuda_lib, tmp_path):
    """VQ-VAE -> latent tree -> loader: batched extraction writes exactly the codes VQVAEModel.encode
    returns, in the reference's on-disk format, and the transformer's loader reads them back."""
    from lvt_b200.config.presets import preset
    from lvt_b200.data import extract_codes, get_latent_video_paths, load_latent_video
    from lvt_b200.modeling import build_model
    cfg = preset("PR-DVQVAE2", ["OUTPUT_DIR", str(tmp_path)])
    cfg.freeze()
    vq = build_model(cfg)
    vq.train(False)
    g = torch.Generator().manual_seed(3)
    videos = [{"image_sequence": torch.rand((6, 3, 64, 64), generator=g), "video_idx": i} for i in range(5)]
    assert extract_codes(vq, videos, "bair_test", str(tmp_path / "inference"), videos_per_batch=2) == {"latents": {}}
    entries = get_latent_video_paths(str(tmp_path / "inference" / "bair_test"))
    assert len(entries) == 5
    for e in entries:
        i = int(os.path.basename(e["video_path"]).split("_")[1])
        got = load_latent_video(e, -1)
        want = vq.encode(videos[i]["image_sequence"].cuda()).cpu()
        assert got.shape == (6, 4, 16, 16) and torch.equal(got, want)


def test_fused_sampling_step_consumes_torch_multinomial_stream(cuda_lib, tmp_path):
    """At temperature 1 the per-position step of sample_slice (q ~ Exp(1) from torch's generator + one kernel for
    softmax / divide / argmax) draws exactly the codes torch.multinomial draws in the reference-shaped per-pixel loop
    (videotransformer.py:161-185) from the same seed: same random stream, same arithmetic."""
    from lvt_b200.config.presets import preset
    from lvt_b200.modeling import build_model
    cfgv = preset("DSFVT", SMALL + ["TEST.EVALUATORS", "VTSampler", "OUTPUT_DIR", str(tmp_path)])
    cfgv.freeze()
    vt = build_model(cfgv)
    vt.train(False)
    video = torch.randint(0, 512, (2, 4, 16, 16, 16)).cuda()
    video[:, :, 15:] = 0
    outs = []
    vt.model.sample_incremental = False
    for mode in ("eager", False):
        vt.sampler_graph = mode
        torch.manual_seed(123)
        outs.append(vt.sample_video(video.clone(), temp=1.0, n_prime=15).cpu())
    same = (outs[0] == outs[1]).float().mean().item()
    assert same >= 0.999, same   # (a differing code would change everything after it)
    assert len(torch.unique(outs[0][:, :, 15])) > 100   # really sampling, not an argmax


@pytest.mark.parametrize("block,kernel,stride,vshape", [((1, 16, 16), (7, 1, 1), (16, 1, 1), (16, 16, 16)),
                                                        ((4, 8, 8), (5, 3, 3), (4, 2, 2), (16, 16, 16))])
def test_incremental_decoder_matches_full_pass(cuda_lib, block, kernel, stride, vshape):
    """IncrementalDecoder (one row per sequence against cached keys / values, csrc/sampler.cu) reproduces the
    teacher-forced logits of the full 256-token pass at every position and channel: max |diff| <= 1e-2 of the logit
    scale (bf16 roundings at the same places; only the summation order differs)."""
    from oracle import lvt_oracle as O
    from lvt_b200.modeling.autoregressive import VTEngine, VTSpec
    from lvt_b200.modeling.autoregressive.incremental import IncrementalDecoder
    layers, batch = 2, 3
    blocks = (block,) * layers
    cfg = O.VTConfig(kernel=kernel, stride=stride, video_shape=vshape, blocks_e=blocks, heads_e=(8,) * layers,
                     blocks_d=blocks, heads_d=(8,) * layers)
    weights = O.synth_weights(O.dsfvt_param_shapes(cfg), seed=77)
    context, slc, slice_idx, ignore = O.synth_vt_batch(batch, seed=9, cfg=cfg)
    eng = VTEngine(VTSpec(kernel=kernel, stride=stride, blocks_e=blocks, heads_e=(8,) * layers, blocks_d=blocks,
                          heads_d=(8,) * layers))
    eng.load_state_dict(weights)
    ws = eng.workspace(batch, cfg.slice_shape, tuple(context.shape[2:]), train=False)
    eng.set_inputs(ws, context, slc, slice_idx, None)
    eng.forward(ws, train=False, want_loss=False)
    full = ws.logits.clone().view(cfg.nc, batch, ws.thw, cfg.nv)
    dec = IncrementalDecoder(eng, ws)
    dec.begin_slice()
    scale = full.abs().max().item()
    worst = 0.0
    for p in range(ws.thw):
        dec.pos.fill_(p)
        dec.decode_row()
        for k in range(cfg.nc):
            dec.channel_logits(k)
            worst = max(worst, (dec.logits - full[k][:, p]).abs().max().item())
    assert worst <= 1e-2 * scale, (worst, scale)


def test_incremental_sampler_runs_and_respects_priming(cuda_lib, tmp_path):
    from lvt_b200.config.presets import preset
    from lvt_b200.modeling import build_model
    cfgv = preset("DSFVT", SMALL + ["TEST.EVALUATORS", "VTSampler", "OUTPUT_DIR", str(tmp_path)])
    cfgv.freeze()
    vt = build_model(cfgv)
    vt.train(False)
    video = torch.randint(0, 512, (2, 4, 16, 16, 16)).cuda()
    video[:, :, 14:] = 0
    torch.manual_seed(5)
    out = vt.sample_video(video.clone(), temp=1.0, n_prime=14).cpu()      # two sampled frames, graph replay
    assert torch.equal(out[:, :, :14], video[:, :, :14].cpu())
    assert int(out.min()) >= 0 and int(out.max()) < 512
    assert len(torch.unique(out[:, :, 14:])) > 100


@pytest.mark.parametrize("B,cfg_name,n_prime", [(1, "DSFVT", 15), (3, "DSFVT", 14), (2, "DSTSVT", 12)])
def test_fused_decode_step_samples_the_codes_of_the_per_stage_path(cuda_lib, tmp_path, monkeypatch, B, cfg_name, n_prime):
    """lvt_vt_decode_step (one persistent kernel per position, csrc/decode_step.cu) against the launch-per-stage
    incremental decoder (lvt_rows_qkv / lvt_attn_row / lvt_rows_linear / lvt_vt_sample_pixel): same arithmetic, same
    random stream => the sampled videos are identical code for code (temperature 1, primed and sampled positions,
    blocks (1,16,16) and (4,8,8))."""
    from lvt_b200.config.presets import preset
    from lvt_b200.modeling import build_model
    small = list(SMALL)
    if cfg_name == "DSTSVT":
        small = ["MODEL.AUTOREGRESSIVE.VT.BLOCKS_E", ((4, 8, 8),) * 2, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_E", (8, 8),
                 "MODEL.AUTOREGRESSIVE.VT.BLOCKS_D", ((4, 8, 8),) * 2, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_D", (8, 8)]
    outs = []
    video = torch.randint(0, 512, (B, 4, 16, 16, 16), generator=torch.Generator().manual_seed(4)).cuda()
    video[:, :, n_prime:] = 0
    for fused in ("0", "1"):
        monkeypatch.setenv("LVT_SAMPLER_FUSED", fused)
        cfgv = preset(cfg_name, small + ["TEST.EVALUATORS", "VTSampler", "OUTPUT_DIR", str(tmp_path)])
        cfgv.freeze()
        torch.manual_seed(21)
        vt = build_model(cfgv)
        vt.train(False)
        vt.model.sample_incremental = True
        torch.manual_seed(77)
        outs.append(vt.sample_video(video.clone(), temp=1.0, n_prime=n_prime).cpu())
    assert torch.equal(outs[0], outs[1]), (outs[0] != outs[1]).float().mean().item()
    assert torch.equal(outs[0][:, :, :n_prime], video[:, :, :n_prime].cpu())
    assert len(torch.unique(outs[0][:, :, n_prime:])) > 100


def test_device_side_slice_construction_feeds_the_engine(cuda_lib):
    """VTEngine.set_inputs_from_videos (latent videos resident in HBM + slice offsets -> context / slice / ignore on the
    device, lvt_b200.data.prepare_slices_batched) gives the same loss as staging the host-side mapper output."""
    import random
    from oracle import lvt_oracle as O
    from lvt_b200.data import prepare_slices, sample_abc, synthetic_latent_video
    from lvt_b200.modeling.autoregressive import VTEngine, VTSpec
    layers, batch = 2, 3
    blocks = ((1, 16, 16),) * layers
    spec = VTSpec(blocks_e=blocks, heads_e=(8,) * layers, blocks_d=blocks, heads_d=(8,) * layers)
    cfg = O.VTConfig(blocks_e=blocks, heads_e=(8,) * layers, blocks_d=blocks, heads_d=(8,) * layers)
    eng = VTEngine(spec)
    eng.load_state_dict(O.synth_weights(O.dsfvt_param_shapes(cfg), seed=5))
    rng = random.Random(2)
    vids = torch.stack([synthetic_latent_video(300 + i) for i in range(batch)])
    abcs = [sample_abc(spec.stride, 16, 1, rng) for _ in range(batch)]
    host = [prepare_slices(vids[i], abcs[i], spec.kernel, spec.stride, 1, -1) for i in range(batch)]
    ws = eng.workspace(batch, (1, 16, 16), (7, 16, 16), train=True)
    eng.set_inputs(ws, *[torch.stack([h[k] for h in host]) for k in ("context", "slice", "slice_idx", "ignore_mask")])
    want = eng.forward(ws, train=True).item()
    staged = [t.clone() for t in (ws.context, ws.slice, ws.slice_idx, ws.ignore)]
    eng.set_inputs_from_videos(ws, vids.cuda(), torch.tensor(abcs).cuda(), n_prime=1)
    for a, b in zip(staged, (ws.context, ws.slice, ws.slice_idx, ws.ignore)):
        assert torch.equal(a, b)
    got = eng.forward(ws, train=True).item()
    assert abs(got - want) <= 1e-6 * abs(want)   # (the loss is accumulated with atomics across blocks)
