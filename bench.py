#!/usr/bin/env python
"""bench.py — DSFVT train step (forward + backward + RMSprop) on synthetic BAIR-shaped latents.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Prints ONE JSON line (rank 0).  metric = latent tokens/s (BASELINE.json): one latent token = one
predicted code index = B * nc * t*h*w per step (1024 per sample).
  value     : device-resident inputs, CUDA-graph replay, CUDA-event timed, max over ranks
  e2e       : same step driven from pinned HOST tensors (H2D of context/slice/slice_idx/ignore
              every step) with the loss read back to the host every step
  roofline  : whole-step useful FLOPs (81.7 GFLOP/sample) / step time vs the measured
              bf16 peak, plus per-kernel figures (QKV GEMM alone; VQ argmin vs HBM)
  cpu_baseline / --impl reference : the UNMODIFIED reference (pip-installed under baseline/_ref by build(),
              kind "reference": its own build_model / forward(data, 'supervised') / optimizer) on the host cores,
              else the oracle port (oracle/lvt_oracle.py, kind "port"); bounded sample (batch 8 slices / 32 frames
              per step); the line declares the steps / warm-up / batch it actually ran
  incumbent : the oracle port and (when baseline/_ref is there) the unmodified reference modules run as PyTorch
              eager ON THE B200 (fp32, TF32, autocast-bf16) -- what a user of the reference would otherwise run
              on this box (SURVEY 8d last row); measurement only
  --workload vqvae : PR-DVQVAE2 train step, frames/s (second half of BASELINE.json's metric), also at N > 1
  N > 1     : adds "strong" (global batch 64 split over the ranks, the reference's semantics,
              data/build.py:62-74) and "dp_parity" (loss trajectory of N ranks vs 1 GPU on a fixed global batch)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOKENS_PER_SAMPLE = 4 * 256            # nc * t*h*w (configs/vt/DSFVT.yaml:12-19)
USEFUL_FLOP_PER_SAMPLE = 81.7e9        # fwd+bwd, one-hot multiplies by zero excluded
PER_GPU_BATCH = 64                     # SOLVER.IMS_PER_BATCH (DSFVT.yaml:26), weak scaling


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples and self.samples[0][1].isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def cpu_reference_arm(steps, warmup, batch=8):
    """The reference's CPU path (oracle port, all host threads): DSFVT fwd + bwd + RMSprop."""
    if live_reference_root():
        return live_reference_arm(steps, warmup, batch=batch)
    import torch
    from oracle import lvt_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.VTConfig()
    sd = {k: v.requires_grad_(True) for k, v in O.synth_weights(O.dsfvt_param_shapes(cfg), seed=1234).items()}
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in sd.items()}
    ctx, slc, sidx, ign = O.synth_vt_batch(batch, seed=5, cfg=cfg)

    def step():
        for p in sd.values():
            p.grad = None
        loss = O.vt_supervised_loss(ctx, slc, sidx, ign, sd, cfg)
        loss.backward()
        with torch.no_grad():
            for k, p in sd.items():
                O.rmsprop_step(p, p.grad, state[k][0], state[k][1], lr=2e-5, alpha=0.95, momentum=0.9)
        return loss.item()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": batch * TOKENS_PER_SAMPLE / dt, "unit": "latent tokens/s", "cores": cores, "kind": "port",
            "sample": f"{steps} DSFVT train steps (fwd+bwd+RMSprop), batch {batch} slices, fp32, torch CPU "
                      f"{torch.get_num_threads()} threads", "ms_per_step": dt * 1e3}


def live_reference_root():
    """baseline/_ref holds the UNMODIFIED reference, pip-installed by __graft_entry__.build() in the authoring
    container (`pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>`); git-ignored,
    not gpurun-ignored, so it is on the GPU box.  None when it is absent (then the arms fall back to the oracle port)."""
    p = os.path.join(ROOT, "baseline", "_ref")
    return p if os.path.isdir(os.path.join(p, "vidgen")) and os.environ.get("LVT_REF_PORT", "0") != "1" else None


def _live_reference_model(preset_name, device="cpu"):
    """build_model(cfg) of the unmodified reference for one of its shipped configurations, with the optimizer its
    own Trainer would build (engine/trainer.py:39-40)."""
    import tempfile
    from oracle import ref_shim
    ref_shim.use_root(live_reference_root())
    ref_shim.install()
    from vidgen.config import get_cfg
    from vidgen.modeling.meta_arch import build_model
    from lvt_b200.config.presets import PRESETS     # the YAML's values as an override list (pure Python, no kernels)
    cfg = get_cfg()
    cfg.merge_from_list(list(PRESETS[preset_name]) + ["MODEL.DEVICE", device, "OUTPUT_DIR", tempfile.mkdtemp(prefix="lvt_ref_")])
    cfg.freeze()
    model = build_model(cfg)
    optimizers, _ = model.configure_optimizers_and_checkpointers()
    model.train()
    return model, [o["optimizer"] for o in optimizers]


def live_reference_arm(steps, warmup, batch=8, workload="dsfvt"):
    """The UNMODIFIED reference through its own public API on the host cores: build_model(cfg), model(data,
    mode='supervised'), backward, optimizer.step(), zero_grad() = the body of Trainer.run_step
    (engine/trainer.py:78-87) on the mapper's per-sample dicts (data/dataset_mapper.py:144-149)."""
    import torch
    from oracle import lvt_oracle as O       # synthetic batch generator only
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    if workload == "vqvae":
        model, opts = _live_reference_model("PR-DVQVAE2")
        x = torch.rand((batch, 3, 64, 64), generator=torch.Generator().manual_seed(3))
        data = [{"image": x[i]} for i in range(batch)]
        units, unit, what = batch, "frames/s", f"PR-DVQVAE2 train steps (fwd+bwd+Adam+EMA), {batch} frames 64x64"
    else:
        model, opts = _live_reference_model("DSFVT")
        ctx, slc, sidx, ign = O.synth_vt_batch(batch, seed=5, cfg=O.VTConfig())
        data = [{"context": ctx[i], "slice": slc[i], "slice_idx": sidx[i], "ignore_mask": ign[i]} for i in range(batch)]
        units, unit, what = batch * TOKENS_PER_SAMPLE, "latent tokens/s", f"DSFVT train steps (fwd+bwd+RMSprop), batch {batch} slices"
    from vidgen.utils.events import EventStorage

    def step():
        losses = sum(model(data, mode="supervised").values())
        losses.backward()
        for o in opts:
            o.step()
        for o in opts:
            o.zero_grad()
        return losses.item()

    with EventStorage(0):
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
    return {"value": units / dt, "unit": unit, "cores": cores, "kind": "reference",
            "sample": f"{steps} {what}, fp32, the unmodified reference modules (baseline/_ref) on torch CPU "
                      f"{torch.get_num_threads()} threads", "ms_per_step": dt * 1e3}


def cpu_vqvae_arm(steps, warmup, frames=32):
    """The reference's CPU path for the VQ-VAE half (oracle port): PR-DVQVAE2 fwd + bwd + Adam(0.9, 0.9) + EMA."""
    if live_reference_root():
        return live_reference_arm(steps, warmup, batch=frames, workload="vqvae")
    import torch
    from oracle import lvt_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = _oracle_vqvae_stepper(O, torch, frames, "cpu")
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": frames / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{steps} PR-DVQVAE2 train steps (fwd+bwd+Adam+EMA), {frames} frames 64x64, fp32, torch CPU "
                      f"{torch.get_num_threads()} threads", "ms_per_step": dt * 1e3}


def _oracle_vqvae_stepper(O, torch, frames, device):
    cfg = O.VQVAEConfig()
    shE, shG = O.vqvae_param_shapes(cfg)
    sdE = {k: v.to(device).requires_grad_(True) for k, v in O.synth_weights(shE, seed=77).items()}
    sdG = {k: v.to(device).requires_grad_(True) for k, v in O.synth_weights(shG, seed=78).items()}
    g = torch.Generator().manual_seed(3)
    cb = (torch.randn(cfg.codebook_num, cfg.codebook_size, cfg.codebook_dim // cfg.codebook_num, generator=g) * 0.3).to(device)
    state = {"cb": cb, "rs": torch.full(cb.shape[:2], 5.0, device=device), "rsum": cb.clone(), "t": 0}
    params = list(sdE.values()) + list(sdG.values())
    opt = torch.optim.Adam(params, lr=3e-4, betas=(0.9, 0.9))
    x = torch.rand((frames, 3, 64, 64), generator=g).to(device)

    def step():
        opt.zero_grad(set_to_none=True)
        losses, aux = O.vqvae_supervised_loss(x, sdE, sdG, state["cb"], state["rs"], state["rsum"], cfg)
        (losses["loss_reconstruction"] + losses["loss_commitment"]).backward()
        opt.step()
        state["cb"], state["rs"], state["rsum"] = aux["codebooks"], aux["running_size"], aux["running_sum"]
    return step


def incumbent_arms(batch, frames, steps=5, warmup=2):
    """PyTorch eager on this B200 (the oracle port of the reference's modules on `cuda`, torch.optim like the
    reference's solver/build.py:62-72) in fp32, TF32 and autocast-bf16: the incumbent a user of the reference
    would run on the same box.  Measurement only -- nothing of it is on the product path."""
    import torch
    from oracle import lvt_oracle as O
    out = {}
    modes = (("fp32", False, None), ("tf32", True, None), ("bf16_autocast", True, torch.bfloat16))
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)

    def timed(step):
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    try:
        # synthetic weights / inputs are built on the host (the oracle's generators mix numpy and torch), then moved
        cfg = O.VTConfig()
        sd_host = O.synth_weights(O.dsfvt_param_shapes(cfg), seed=1234)
        batch_host = O.synth_vt_batch(batch, seed=5, cfg=cfg)
        with torch.device("cuda"):
            # ---- DSFVT train step, same batch as the headline line
            sd = {k: v.cuda().requires_grad_(True) for k, v in sd_host.items()}
            ctx, slc, sidx, ign = (t.cuda() for t in batch_host)
            opt = torch.optim.RMSprop(list(sd.values()), lr=2e-5, alpha=0.95, momentum=0.9, eps=1e-8)
            res = {}
            for name, tf32, ac in modes:
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32

                def step():
                    opt.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=ac, enabled=ac is not None):
                        loss = O.vt_supervised_loss(ctx, slc, sidx, ign, sd, cfg)
                    loss.backward()
                    opt.step()
                try:
                    ms = timed(step)
                    res[name] = {"ms_per_step": ms, "value": batch * TOKENS_PER_SAMPLE / (ms * 1e-3), "unit": "latent tokens/s"}
                except Exception as ex:
                    res[name] = {"error": repr(ex)[:200]}
            out["dsfvt"] = dict(res, batch=batch, note="oracle port of the reference modules (oracle/lvt_oracle.py) as "
                                "PyTorch eager on cuda:0 + torch.optim.RMSprop; incl. the reference's dense one-hot convs")
            del sd, opt, ctx, slc, sidx, ign
            torch.cuda.empty_cache()
            # ---- PR-DVQVAE2 train step
            res = {}
            for name, tf32, ac in modes:
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                try:
                    with torch.device("cpu"):
                        inner = _oracle_vqvae_stepper(O, torch, frames, "cuda")

                    def step():
                        with torch.autocast("cuda", dtype=ac, enabled=ac is not None):
                            inner()
                    ms = timed(step)
                    res[name] = {"ms_per_step": ms, "value": frames / (ms * 1e-3), "unit": "frames/s"}
                except Exception as ex:
                    res[name] = {"error": repr(ex)[:200]}
            out["vqvae"] = dict(res, frames=frames, note="oracle port of PR-DVQVAE2 (cuDNN convs, ATen codebook search) "
                                "as PyTorch eager on cuda:0 + torch.optim.Adam(0.9, 0.9) + EMA")
            torch.cuda.empty_cache()
        if live_reference_root():
            out["reference_modules"] = _incumbent_live_reference(batch, frames, timed, modes[:2])
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return out


def _incumbent_live_reference(batch, frames, timed, modes):
    """The UNMODIFIED reference (baseline/_ref) on cuda:0 through its own API: build_model(cfg) with MODEL.DEVICE
    cuda, model(data, mode='supervised'), backward, its own optimizer -- Trainer.run_step's body (engine/trainer.py:78-87)."""
    import torch
    from oracle import lvt_oracle as O
    res = {}
    for wl, preset_name in (("dsfvt", "DSFVT"), ("vqvae", "PR-DVQVAE2")):
        try:
            torch.manual_seed(0)
            model, opts = _live_reference_model(preset_name, device="cuda")
            from vidgen.utils.events import EventStorage
            if wl == "dsfvt":
                ctx, slc, sidx, ign = (t.cuda() for t in O.synth_vt_batch(batch, seed=5, cfg=O.VTConfig()))
                data = [{"context": ctx[i], "slice": slc[i], "slice_idx": sidx[i], "ignore_mask": ign[i]} for i in range(batch)]
                units, unit = batch * TOKENS_PER_SAMPLE, "latent tokens/s"
            else:
                x = torch.rand((frames, 3, 64, 64), generator=torch.Generator().manual_seed(3)).cuda()
                data = [{"image": x[i]} for i in range(frames)]
                units, unit = frames, "frames/s"

            def step():
                sum(model(data, mode="supervised").values()).backward()
                for o in opts:
                    o.step()
                for o in opts:
                    o.zero_grad()
            r = {}
            with EventStorage(0):
                for name, tf32, _ in modes:
                    torch.backends.cuda.matmul.allow_tf32 = tf32
                    torch.backends.cudnn.allow_tf32 = tf32
                    ms = timed(step)
                    r[name] = {"ms_per_step": ms, "value": units / (ms * 1e-3), "unit": unit}
            res[wl] = dict(r, units_per_step=units)
            del model, opts, data
            torch.cuda.empty_cache()
        except Exception as ex:
            res[wl] = {"error": repr(ex)[:300]}
    res["note"] = ("the unmodified reference package (pip-installed under baseline/_ref) as PyTorch eager on cuda:0, "
                   "its own build_model / forward(data, 'supervised') / optimizer")
    return res


def step_traffic():
    """DRAM bytes of one whole DSFVT step from the tracked ncu launch list (profiles/*_step_traffic.json, written by
    tools/launch_table.py --json from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over the step)."""
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_step_traffic.json")))
    if not cands:
        return None, "no tracked step launch list"
    d = json.load(open(cands[-1]))
    return d.get("dram_bytes"), (f"DRAM bytes of one whole step ({d.get('launches')} launches), ncu dram__bytes_read+write, "
                                f"profiles/{os.path.basename(cands[-1])} <- {d.get('source')}")


def vqvae_main(args, rank, world, local_rank):
    import torch
    from lvt_b200 import _lib
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _lib.require_device()
    line = vqvae_measure(args, rank, world, local_rank, dist)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def vqvae_measure(args, rank, world, local_rank, dist, with_cpu=True):
    """BASELINE.json config 3: PR-DVQVAE2 training on synthetic 16-frame 64x64 clips, data-parallel: 32 clips = 512
    frames per GPU and step (weak scaling), EMA counts / sums and the flat gradient summed over ranks with NCCL
    (vq_embedding.py:44-59, ae.py:69-73); CUDA-graph replay, max over ranks.  Returns the JSON line (rank 0)."""
    import torch
    from lvt_b200 import _lib
    from lvt_b200.modeling.vqvae_engine import GraphedVQVAEStep, VQVAEEngine, VQVAESpec
    _, _, tf_sust, peak_src = measured_peaks()
    nfr = 512
    if args.strong:  # the reference's semantics: IMS_PER_BATCH (32 clips = 512 frames) is the GLOBAL batch (data/build.py:62-74)
        nfr = 512 // max(1, world)
    spec = VQVAESpec(n_layers=2)
    ve = VQVAEEngine(spec)
    gq = torch.Generator().manual_seed(7)
    init_w = {}
    for name, shp in spec.param_shapes().items():
        fan = 1
        for s_ in shp[1:]:
            fan *= s_
        init_w[name] = torch.randn(shp, generator=gq) / (fan ** 0.5) if len(shp) > 1 else torch.zeros(shp)
    ve.store.load(init_w)
    ve.load_state_dict(codebook=torch.randn(4, 512, 64, generator=gq) * 0.3, running_size=torch.full((4, 512), 5.0))
    ve.init_optimizer(lr=3e-4, betas=(0.9, 0.9))
    vw = ve.workspace(nfr, train=True)
    host = torch.rand((nfr, 3, 64, 64), generator=torch.Generator().manual_seed(100 + rank)).pin_memory()
    vw.x.copy_(host)
    allreduce = (lambda t: dist.all_reduce(t)) if world > 1 else None
    step = GraphedVQVAEStep(ve, vw, world_size=world, allreduce=allreduce)
    n_before = _lib.launch_count()
    step.capture(warmup=2)
    # kernels inside the captured graphs, per replayed step (capture = 2 eager warm-up steps + 1 captured step,
    # each eager step also launching Adam once outside the graphs)
    step.graph_launches = (_lib.launch_count() - n_before - 2) // 3

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step.step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if sampler else None
    launches = (_lib.launch_count() - n0) + step.graph_launches * args.steps
    # end to end: frames from pinned host memory in, the two losses out, every step.  The H2D copy of batch i+1 (25 MB)
    # runs on a copy stream into a staging buffer while step i runs -- what a pin_memory DataLoader does -- and is
    # moved into the graph's static input buffer device-to-device before step i+1.
    stage, copy_stream = torch.empty_like(vw.x), torch.cuda.Stream()
    copied, committed = torch.cuda.Event(), torch.cuda.Event()

    def prefetch():
        copy_stream.wait_event(committed)
        with torch.cuda.stream(copy_stream):
            stage.copy_(host, non_blocking=True)
            copied.record()

    committed.record()
    t0 = time.perf_counter()
    e0.record()
    prefetch()
    for i in range(args.steps):
        torch.cuda.current_stream().wait_event(copied)
        vw.x.copy_(stage, non_blocking=True)
        committed.record()
        step.step()
        if i + 1 < args.steps:
            prefetch()
        _ = vw.loss.tolist()
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps
    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    line = None
    if rank == 0:
        n_gpus = max(1, world)
        tf = 5.57e9 * nfr / (ms * 1e-3) / 1e12
        cpu = cpu_vqvae_arm(3, 1, frames=32) if (with_cpu and n_gpus == 1 and not args.quick) else None
        line = ({
            "metric": "VQ-VAE frames/sec (PR-DVQVAE2 train step)", "value": nfr * n_gpus / (ms * 1e-3), "unit": "frames/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"PR-DVQVAE2 training on synthetic 16-frame 64x64 clips, {nfr // 16} clips ({nfr} frames) per GPU "
                                   "and step, Adam + EMA codebook, data-parallel (EMA statistics + flat gradient all-reduced)",
                       "frames_per_gpu": nfr, "global_frames": nfr * n_gpus, "parallelism": f"dp{n_gpus}",
                       "l2": f"{nfr} frames of activations ({nfr * 2.3e-3:.2f} GB) exceed the 126 MB L2; no explicit flush"},
            "losses": vw.loss.tolist(),
            "e2e": {"value": nfr * n_gpus / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(host.numel() * 4),
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches), "gpu_launches_per_step": int(launches // max(1, args.steps)),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": tf_sust, "unit": "TFLOP/s", "frac": tf / tf_sust,
                         "traffic": None, "peak_source": peak_src,
                         "kernel": "gemm_bf16_kernel implicit-GEMM convolutions: 5.57 GFLOP per frame fwd+bwd / step time, per GPU"},
            "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None)})
    del step, ve, vw
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="slices per GPU")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="one gradient all-reduce after the backward instead of two overlapped buckets")
    ap.add_argument("--workload", default="dsfvt", choices=["dsfvt", "vqvae"],
                    help="dsfvt (default, the headline line) | vqvae: PR-DVQVAE2 data-parallel training, frames/s "
                         "(BASELINE.json config 3)")
    ap.add_argument("--quick", action="store_true", help="DSFVT step only: skip the per-kernel figures and the CPU baseline")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the per-GPU batch is the config's global batch "
                    "(64 slices / 512 frames) divided by the number of ranks, as the reference does (data/build.py:62-74)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "DSFVT training on synthetic BAIR-shaped latents (16x16x16 grid, nc=4, 512-way "
                          "codebook), configs/vt/DSFVT.yaml network (8+8 layers, d=512, 8 heads), RMSprop",
              "per_gpu_batch": args.batch, "global_batch": args.batch * max(1, args.gpus),
              "tokens_per_sample": TOKENS_PER_SAMPLE, "parallelism": f"dp{max(1, args.gpus)}",
              "l2": "working set per step (>5 GB activations + 0.6 GB weights/optimizer state) exceeds the "
                    "126 MB L2; no explicit flush"}

    if args.strong:
        assert PER_GPU_BATCH % max(1, world) == 0
        args.batch = PER_GPU_BATCH // max(1, world)
        config.update(per_gpu_batch=args.batch, global_batch=PER_GPU_BATCH)
    if args.workload == "vqvae" and args.impl == "ours":
        return vqvae_main(args, rank, world, local_rank)
    if args.impl == "reference":
        # The reference's own CPU implementation of the path on the box's host cores (oracle port, all threads).
        # A step is a BOUNDED sample of the workload (8 slices / 32 frames instead of 64 / 512: a CPU step of the
        # full batch takes ~8 s); the line declares exactly what ran.  tokens/s and frames/s are batch-normalised.
        if rank != 0:
            return
        steps_run, warm_run = max(1, min(args.steps, 20)), max(1, min(args.warmup, 5))
        if args.workload == "vqvae":
            r = cpu_vqvae_arm(steps_run, warm_run, frames=32)
            metric, unit, sample_cfg = "VQ-VAE frames/sec (PR-DVQVAE2 train step)", "frames/s", {"frames_per_step": 32}
            wl = "PR-DVQVAE2 training on synthetic 64x64 frames, Adam + EMA codebook"
        else:
            r = cpu_reference_arm(steps_run, warm_run, batch=8)
            metric, unit, sample_cfg = "latent tokens/sec DSFVT train step", "latent tokens/s", {"per_gpu_batch": 8, "global_batch": 8}
            wl = config["workload"]
        base_cfg = config if args.workload != "vqvae" else {"l2": "n/a (CPU)"}
        ref_config = dict(base_cfg, workload=wl, parallelism="cpu", **sample_cfg)
        ref_config["sample_of"] = ("bounded sample of the GPU arm's workload: same network, same synthetic data generator, "
                                   "smaller batch per step (the metric is per token / per frame)")
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus,
                "steps": steps_run, "warmup": warm_run, "steps_requested": args.steps, "warmup_requested": args.warmup,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": ref_config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    from lvt_b200 import _lib
    from lvt_b200.data import synthetic_vt_batch
    from lvt_b200.modeling.autoregressive import VTEngine, VTSpec
    from lvt_b200.modeling.autoregressive.vt_engine import GraphedTrainStep

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # the gradient all-reduce runs NEXT TO the backward: it gets LVT_COMM_SMS SMs (NCCL_MAX_CTAS), the
        # overlapped backward segments are captured with that many SMs fewer (GraphedTrainStep)
        os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("LVT_COMM_SMS", "8") if int(os.environ.get("LVT_COMM_SMS", "8")) > 0 else "32")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _lib.require_device()
    hbm_peak, tf_burst, tf_sust, peak_src = measured_peaks()

    spec = VTSpec()
    eng = VTEngine(spec)
    # random-init weights of the DSFVT architecture (same generator family as the parity tests)
    g = torch.Generator().manual_seed(1234)
    init = {}
    for name, shp in spec.param_shapes().items():
        if name.endswith("_bank"):
            t = torch.randn(shp, generator=g) * 0.1
        elif "layer_norm.weight" in name or name.endswith("ffn.0.weight"):
            t = torch.ones(shp)
        elif len(shp) == 1:
            t = torch.zeros(shp)
        elif "embed" in name:
            t = torch.randn(shp, generator=g) * 0.5
        elif ".w_" in name:
            t = torch.randn(shp, generator=g) / (shp[1] ** 0.5)
        elif name == "encoder.conv.weight":
            t = torch.randn(shp, generator=g) * 0.2
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            t = torch.randn(shp, generator=g) / (fan_in ** 0.5)
        init[name] = t
    eng.load_state_dict(init)
    eng.init_optimizer("rmsprop", lr=2e-5, alpha=0.95, momentum=0.9, eps=1e-8)

    B = args.batch
    host = [t.pin_memory() for t in synthetic_vt_batch(B, seed=1000 + rank)]
    ctx_shape = tuple(host[0].shape[2:])
    ws = eng.workspace(B, (1, 16, 16), ctx_shape, train=True)
    eng.set_inputs(ws, *host)

    allreduce = None
    if world > 1:
        def allreduce(flat):
            dist.all_reduce(flat)
    stepper = GraphedTrainStep(eng, ws, world_size=world, allreduce=allreduce, overlap=not args.no_overlap)
    if args.no_graph:
        def one_step():
            eng.zero_grad(); eng.forward(ws, train=True); eng.backward(ws)
            if allreduce:
                allreduce(eng.store.grad)
            eng.optimizer_step(1.0 / world)
        n0 = _lib.launch_count(); one_step(); launches_per_step = _lib.launch_count() - n0
    else:
        stepper.capture(warmup=2)
        one_step = stepper.step
        launches_per_step = stepper.launches_per_step

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: device-resident inputs
    for _ in range(max(3, args.warmup)):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if sampler else None
    loss_value = ws.loss.item()

    # ---------------- e2e: pinned host inputs in, loss out, every step
    h2d = sum(t.numel() * t.element_size() for t in host[:3]) + host[3].numel()  # ignore mask as uint8
    ign8 = host[3].to(torch.uint8).pin_memory()
    # (every step copies ITS batch from pinned host memory and reads ITS loss back; the copy of batch i+1 is started
    # on a copy stream while step i runs -- what a pin_memory DataLoader does -- and committed device-to-device)
    for _ in range(2):
        eng.set_inputs(ws, host[0], host[1], host[2], ign8); one_step(); ws.loss.item()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    eng.prefetch_inputs(ws, host[0], host[1], host[2], ign8)
    for i in range(args.steps):
        eng.commit_inputs(ws)
        one_step()
        if i + 1 < args.steps:   # queued behind the commit only (an event), so it overlaps the step just launched
            eng.prefetch_inputs(ws, host[0], host[1], host[2], ign8)
        _ = ws.loss.item()
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps

    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()

    n_gpus = max(1, world)
    tokens_per_step = B * TOKENS_PER_SAMPLE * n_gpus
    value = tokens_per_step / (ms * 1e-3)
    e2e = tokens_per_step / (ms_e2e * 1e-3)

    # ---------------- per-kernel roofline figures (rank 0, timed alone)
    extra = {}
    incumbent = None
    # (single-GPU runs only: at N > 1 the other ranks would sit in a collective while rank 0 measures, and the
    # Trainer figure would issue rank-0-only collectives)
    if rank == 0 and world == 1 and not args.quick:
        from lvt_b200 import ops
        from lvt_b200.ops import Operand
        # BlockLocalAttention layer alone: forward + backward of ONE layer (LN, QKV, attention, proj, FFN and
        # all their gradients) as a CUDA-graph replay on the step's own buffers; 4.832 GFLOP per 256-token sequence.
        try:
            ly, prefix = ws.layers[0], "encoder.block_local_attention.0."
            gb = torch.cuda.CUDAGraph()
            sidestream = torch.cuda.Stream()
            sidestream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(sidestream):
                def bla():
                    eng._layer_fwd(prefix, ws, ly, ws.x0, ly.y, causal=False)
                    eng._layer_bwd(prefix, ws, ly, ws.x0, ws.dy, ws.dy_bf16, ws.dy, ws.dy_bf16)
                    eng._side_join()
                bla()
                torch.cuda.synchronize()
                with torch.cuda.graph(gb):
                    bla()
            torch.cuda.current_stream().wait_stream(sidestream)
            torch.cuda.synchronize()
            for _ in range(3):
                gb.replay()
            e0.record()
            for _ in range(20):
                gb.replay()
            e1.record(); torch.cuda.synchronize()
            t_bla = e0.elapsed_time(e1) / 20 * 1e-3
            fl_bla = ws.nseq * 4.832e9
            extra["bla_block"] = {"bound": "tensor", "achieved": fl_bla / t_bla / 1e12, "peak": tf_sust, "unit": "TFLOP/s",
                                  "frac": fl_bla / t_bla / 1e12 / tf_sust, "frac_of_burst": fl_bla / t_bla / 1e12 / tf_burst,
                                  "us": t_bla * 1e6, "sequences": ws.nseq,
                                  "note": "one BlockLocalAttention layer, forward + backward incl. weight gradients, "
                                          "4.832 GFLOP per 256-token sequence (SURVEY 8d), CUDA-graph replay; target 0.60"}
            eng.zero_grad()
        except Exception as ex:
            extra["bla_block"] = {"error": repr(ex)[:300]}
        M, d, N = B * 256, 512, 3072
        a = torch.randn(M, d, device="cuda").to(torch.bfloat16)
        w = torch.randn(24, d, 128, device="cuda").to(torch.bfloat16)
        o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

        def qkv():
            ops.gemm(M, N, d, Operand(a.data_ptr(), d), Operand(w.data_ptr(), 128, mn_major=True, cin=128, s_blk=d * 128),
                     Operand(o.data_ptr(), N), out_bf16=o)
        for _ in range(5):
            qkv()
        e0.record()
        for _ in range(20):
            qkv()
        e1.record(); torch.cuda.synchronize()
        t_qkv = e0.elapsed_time(e1) / 20 * 1e-3
        extra["qkv_gemm"] = {"bound": "tensor", "achieved": 2.0 * M * N * d / t_qkv / 1e12, "peak": tf_burst,
                             "unit": "TFLOP/s", "frac": 2.0 * M * N * d / t_qkv / 1e12 / tf_burst,
                             "shape": [M, N, d], "us": t_qkv * 1e6}
        # VQ codebook argmin: 1056 algorithmic bytes per latent position
        nfr = 4096  # 2^20 positions
        z = torch.randn(nfr, 256, 16, 16, device="cuda") * 0.3
        cb = torch.randn(4, 512, 64, device="cuda") * 0.3
        for _ in range(2):
            ops.vq_argmin(z, cb)
        e0.record()
        for _ in range(5):
            ops.vq_argmin(z, cb)
        e1.record(); torch.cuda.synchronize()
        t_vq = e0.elapsed_time(e1) / 5 * 1e-3
        pos = nfr * 256
        extra["vq_argmin"] = {"bound": "hbm", "achieved": pos * 1056 / t_vq / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": pos * 1056 / t_vq / 1e9 / hbm_peak, "positions": pos, "ms": t_vq * 1e3,
                              "positions_per_s": pos / t_vq, "frames_per_s_equiv": nfr / t_vq,
                              "note": "tf32 tcgen05 scan of all 512 codes + exact fp32 re-rank of the candidates (bit-exact "
                                      "indices); 262144 FLOP/position: the tf32 tensor pipe (~0.27 ms per 2^20 positions), "
                                      "the alu pipe of the threshold scan and the per-tile dependency chain, not HBM, bound "
                                      "it (DESIGN.md 5; round-1 kernel: LVT_VQ_TC1=1)"}
        del z, a, w, o
        # VQ-VAE (PR-DVQVAE2) on synthetic 16-frame 64x64 clips: 32 clips = 512 frames per step
        # (configs/vqvae/Base-VQVAE.yaml IMS_PER_BATCH 32); second half of BASELINE.json's metric.
        try:
            from lvt_b200.modeling.vqvae_engine import VQVAEEngine, VQVAESpec
            nfr = 512
            ve = VQVAEEngine(VQVAESpec(n_layers=2))
            gq = torch.Generator().manual_seed(7)
            init_w = {}
            for name, shp in VQVAESpec(n_layers=2).param_shapes().items():
                fan = 1
                for s_ in shp[1:]:
                    fan *= s_
                init_w[name] = torch.randn(shp, generator=gq) / (fan ** 0.5) if len(shp) > 1 else torch.zeros(shp)
            ve.store.load(init_w)
            ve.load_state_dict(codebook=torch.randn(4, 512, 64, generator=gq) * 0.3,
                               running_size=torch.full((4, 512), 5.0))
            ve.init_optimizer(lr=3e-4, betas=(0.9, 0.9))
            vw = ve.workspace(nfr, train=True)
            vw.x.copy_(torch.rand((nfr, 3, 64, 64), generator=gq))
            from lvt_b200.modeling.vqvae_engine import GraphedVQVAEStep
            vstep = GraphedVQVAEStep(ve, vw)
            vstep.capture(warmup=2)
            for _ in range(3):
                vstep.step()
            e0.record()
            for _ in range(10):
                vstep.step()
            e1.record(); torch.cuda.synchronize()
            t_tr = e0.elapsed_time(e1) / 10 * 1e-3
            for _ in range(2):
                ve.inference(vw)
            e0.record()
            for _ in range(10):
                ve.inference(vw)
            e1.record(); torch.cuda.synchronize()
            t_inf = e0.elapsed_time(e1) / 10 * 1e-3
            for _ in range(2):
                ve.inference(vw, precise=True)
            e0.record()
            for _ in range(10):
                ve.inference(vw, precise=True)
            e1.record(); torch.cuda.synchronize()
            t_infp = e0.elapsed_time(e1) / 10 * 1e-3
            extra["vqvae"] = {"train_frames_per_s": nfr / t_tr, "train_ms_per_step": t_tr * 1e3,
                              "inference_frames_per_s": nfr / t_inf, "frames_per_step": nfr,
                              "inference_precise_encoder_frames_per_s": nfr / t_infp,
                              "train_tflops": 5.57e9 * nfr / t_tr / 1e12, "train_frac_of_bf16_peak": 5.57e9 * nfr / t_tr / 1e12 / tf_sust,
                              "losses": vw.loss.tolist(),
                              "note": "PR-DVQVAE2 fwd+bwd+Adam+EMA, CUDA-graph replay (Adam launch eager), 5.57 GFLOP/frame"}
        except Exception as ex:  # the DSFVT line must survive a VQ-VAE problem
            extra["vqvae"] = {"error": repr(ex)[:300]}
        # BASELINE.json config 5: autoregressive sampling of one 16x16 latent frame (256 positions x 4 channels),
        # one video, full 8+8-layer DSFVT: CUDA-graph replay per position vs the per-pixel Python loop.
        try:
            from lvt_b200.config.presets import preset
            from lvt_b200.modeling import build_model
            cfgv = preset("DSFVT", ["TEST.EVALUATORS", "VTSampler", "OUTPUT_DIR", "/tmp/lvt_bench_out"])
            cfgv.freeze()
            vt = build_model(cfgv)
            vt.train(False)
            video = torch.randint(0, 512, (1, 4, 16, 16, 16), device="cuda")
            res = {}
            for mode, flag in (("graph", True), ("per_pixel_loop", False)):
                vt.sampler_graph = flag
                vt.sample_video(video.clone(), n_prime=15)  # warm-up
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                vt.sample_video(video.clone(), n_prime=15)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                res[mode] = {"latent_frames_per_s": 1.0 / dt, "ms_per_position": dt * 1e3 / 256}
            extra["sampler"] = dict(res, note="one video, one sampled frame (slice) after 15 primed frames; "
                                              "wall clock incl. the encoder pass of the slice")
            del vt
        except Exception as ex:
            extra["sampler"] = {"error": repr(ex)[:300]}

        # BASELINE.json config 5 at script level: scripts/generate_videos.py sample_videos(args) -- 5 priming PNGs ->
        # VQ-VAE codes -> 11 sampled latent frames (DSFVT, 8+8 layers) -> VQ-VAE decode -> 16 PNGs; frames/s of the
        # whole call (model construction included, as in the reference's 338 s figure), second call = warm.
        try:
            import argparse as _ap
            import tempfile
            import numpy as np
            from PIL import Image
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import generate_videos as gv
            tmp = tempfile.mkdtemp(prefix="lvt_gen_")
            rs = np.random.RandomState(0)
            for i in range(5):
                Image.fromarray(rs.randint(0, 255, (64, 64, 3), dtype=np.uint8)).save(os.path.join(tmp, f"{i}.png"))
            ga = _ap.Namespace(video_dir=tmp, config_file="DSFVT", vqvae_config="PR-DVQVAE2",
                               out_dir=os.path.join(tmp, "out"), n_frames=16)
            times = []
            for _ in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                vid = gv.sample_videos(ga)
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
            extra["generate_videos"] = {
                "frames": 16, "sampled_frames": 11, "seconds_first_call": times[0], "seconds": times[1],
                "frames_per_s": 16 / times[1], "sampled_frames_per_s": 11 / times[1], "output_shape": list(vid.shape),
                "note": "scripts/generate_videos.py sample_videos(): PNG priming frames in, PNG video out, random-init "
                        "weights (no checkpoints offline); the reference's CPU run of the same script: 338 s (BASELINE.md)"}
        except Exception as ex:
            extra["generate_videos"] = {"error": repr(ex)[:300]}
        # The reference-facing training surface: Trainer.run_step -> model(list[dict], 'supervised') -> loss.backward()
        # -> optimizer.step() (engine/trainer.py), batches in the DatasetMapper's per-sample dict format on the host.
        try:
            import itertools
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import train_net
            from lvt_b200.config.presets import preset
            from lvt_b200.engine import Trainer
            from lvt_b200.utils.events import EventStorage
            cfgt = preset("DSFVT", ["OUTPUT_DIR", "/tmp/lvt_bench_trainer", "SOLVER.IMS_PER_BATCH", B,
                                    "SOLVER.CHECKPOINT_PERIOD", 0, "SEED", 1])
            cfgt.freeze()
            gen = train_net.synthetic_loader(cfgt)
            batches = [next(gen) for _ in range(3)]   # per-sample dicts of host tensors, as the mapper yields them
            tr = Trainer(cfgt, data_loader=itertools.cycle(batches))
            tr.model.train()
            with EventStorage(0) as tr.storage:
                tr.iter = 0
                for _ in range(3):
                    tr.run_step(); tr.iter += 1
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                nit = 10
                for _ in range(nit):
                    tr.run_step(); tr.iter += 1
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / nit
            extra["trainer_path"] = {"ms_per_step": dt * 1e3, "value": B * TOKENS_PER_SAMPLE / dt, "unit": "latent tokens/s",
                                     "graph_replay": bool(getattr(tr.model, "_graphed", False)),
                                     "note": "tools/train_net.py's Trainer.run_step on the DSFVT preset, batch 64: per-sample "
                                             "dicts stacked on the host, H2D, forward + backward as one CUDA-graph replay, "
                                             "RMSprop, LR scheduler; wall clock"}
            del tr
        except Exception as ex:
            extra["trainer_path"] = {"error": repr(ex)[:300]}

    # ---------------- N > 1: strong-scaling line (reference semantics) and data-parallel loss parity
    strong, dp_par = None, None
    if world > 1 and not args.quick and not args.strong:
        Bs = PER_GPU_BATCH // world
        if Bs >= 1 and PER_GPU_BATCH % world == 0:
            host_s = [t[:Bs].contiguous() for t in host]
            ws_s = eng.workspace(Bs, (1, 16, 16), ctx_shape, train=True)
            eng.set_inputs(ws_s, *host_s)
            st_s = GraphedTrainStep(eng, ws_s, world_size=world, allreduce=allreduce, overlap=not args.no_overlap)
            st_s.capture(warmup=2)
            for _ in range(3):
                st_s.step()
            barrier()
            e0.record()
            for _ in range(args.steps):
                st_s.step()
            e1.record()
            barrier()
            ts = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda")
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            ms_s = ts.item()
            strong = {"scaling": "strong", "global_batch": PER_GPU_BATCH, "per_gpu_batch": Bs, "ms_per_step": ms_s,
                      "value": PER_GPU_BATCH * TOKENS_PER_SAMPLE / (ms_s * 1e-3), "unit": "latent tokens/s",
                      "note": "the reference divides SOLVER.IMS_PER_BATCH (64) by the world size (data/build.py:62-74)"}
        try:
            from lvt_b200.utils.dp_check import run_dp_parity
            dp_par = run_dp_parity(rank, world, dist)
        except Exception as ex:
            dp_par = {"error": repr(ex)[:300]}

    # second half of BASELINE.json's metric at N > 1: PR-DVQVAE2 data-parallel frames/s (every rank takes part; at
    # N = 1 the same step is timed under roofline.kernels.vqvae)
    vq_dp = None
    if world > 1 and not args.quick and not args.strong:
        try:
            vq_dp = vqvae_measure(args, rank, world, local_rank, dist, with_cpu=False)
        except Exception as ex:
            vq_dp = {"error": repr(ex)[:300]}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    if not args.quick and n_gpus == 1:
        try:
            incumbent = incumbent_arms(B, 512)
        except Exception as ex:
            incumbent = {"error": repr(ex)[:300]}
    cpu = cpu_reference_arm(2, 1, batch=8) if not args.quick else {k: None for k in ("value", "unit", "cores", "kind", "sample")}
    traffic, traffic_note = step_traffic()
    step_flops = USEFUL_FLOP_PER_SAMPLE * B  # per GPU
    achieved = step_flops / (ms * 1e-3) / 1e12
    line = {
        "metric": "latent tokens/sec DSFVT train step", "value": value, "unit": "latent tokens/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
        "loss": loss_value,
        "e2e": {"value": e2e, "unit": "latent tokens/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": int(launches_per_step) * args.steps,
        "gpu_launches_per_step": int(launches_per_step),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_sust, "unit": "TFLOP/s",
                     "frac": achieved / tf_sust, "traffic": traffic, "peak_source": peak_src,
                     "traffic_note": traffic_note + "; achieved / peak are per step too",
                     "kernel": "gemm_bf16_kernel (tcgen05) — whole-step useful FLOPs / step time, per GPU",
                     "kernels": extra},
        "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
    }
    if world > 1 and not args.no_graph:
        line["comm"] = {"collective": "NCCL all-reduce of the flat fp32 gradient", "buckets": len(stepper.graphs),
                        "overlap": bool(stepper.overlap), "comm_sms": stepper.comm_sms,
                        "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS"), "grad_bytes": int(eng.store.numel * 4)}
    if args.strong:
        line["scaling"] = "strong"
    if incumbent is not None:
        line["incumbent"] = incumbent
    if strong is not None:
        line["strong"] = strong
    if dp_par is not None:
        line["dp_parity"] = dp_par
    if vq_dp is not None:
        line["vqvae_dp"] = {k: vq_dp.get(k) for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "scaling", "e2e",
                                                      "config", "losses", "error") if k in vq_dp}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
