"""Data-parallel loss parity: with a fixed GLOBAL batch and identical seeds, W ranks x (B/W) slices with
the flat-gradient all-reduce must follow the 1-GPU loss trajectory (reference: DDP + per-rank batch = IMS_PER_BATCH /
world, data/build.py:62-74, meta_arch/vt.py:61-63).  Used by tools/dp_parity.py, tests/test_dp_gpu.py and by bench.py
when it runs on more than one GPU (so that the driver's 2/4/8-GPU runs print the evidence)."""
import torch


def run_dp_parity(rank, world, dist=None, layers=2, global_b=16, steps=4, tol=1e-3):
    """Returns {"world", "global_batch", "losses", "losses_1gpu", "max_rel_diff", "ok"} on rank 0, None elsewhere.
    The 1-GPU trajectory is computed on rank 0 with the whole global batch after the W-rank run."""
    from ..data import synthetic_vt_batch
    from ..modeling.autoregressive import VTEngine, VTSpec
    spec = VTSpec(blocks_e=((1, 16, 16),) * layers, heads_e=(8,) * layers, blocks_d=((1, 16, 16),) * layers,
                  heads_d=(8,) * layers)

    def make_engine():
        eng = VTEngine(spec)
        g = torch.Generator().manual_seed(1234)
        init = {}
        for name, shp in spec.param_shapes().items():
            if "layer_norm.weight" in name or name.endswith("ffn.0.weight"):
                init[name] = torch.ones(shp)
            elif len(shp) == 1:
                init[name] = torch.zeros(shp)
            else:
                fan = 1
                for s_ in shp[1:]:
                    fan *= s_
                init[name] = torch.randn(shp, generator=g) * (0.1 if name.endswith("_bank") else 1.0 / fan ** 0.5)
        eng.load_state_dict(init)
        eng.init_optimizer("rmsprop", lr=2e-5, alpha=0.95, momentum=0.9, eps=1e-8)
        return eng

    assert global_b % world == 0, "the global batch must divide over the ranks"
    batch = synthetic_vt_batch(global_b, seed=99)  # context, slice, slice_idx, ignore_mask of the GLOBAL batch
    per = global_b // world
    mine = [t[rank * per:(rank + 1) * per].contiguous() for t in batch]
    eng = make_engine()
    ws = eng.workspace(per, (1, 16, 16), tuple(mine[0].shape[2:]), train=True)
    eng.set_inputs(ws, *mine)
    hook = (lambda flat: dist.all_reduce(flat)) if world > 1 else None
    stepper = None
    if world > 1:
        # the production multi-GPU path: CUDA-graph segments with the bucketed all-reduce overlapped (GraphedTrainStep)
        from ..modeling.autoregressive.vt_engine import GraphedTrainStep
        stepper = GraphedTrainStep(eng, ws, world_size=world, allreduce=hook, overlap=True)
        stepper.capture(warmup=1)
    dp = []
    for _ in range(steps):
        loss = (stepper.step() if stepper is not None else eng.train_step(ws, grad_hook=hook, grad_scale=1.0 / world)).clone()
        if world > 1:
            dist.all_reduce(loss)
            loss /= world
        dp.append(loss.item())
    out = None
    if rank == 0:
        ref_eng = make_engine()
        ws1 = ref_eng.workspace(global_b, (1, 16, 16), tuple(batch[0].shape[2:]), train=True)
        ref_eng.set_inputs(ws1, *batch)
        ref = [ref_eng.train_step(ws1).item() for _ in range(steps)]
        err = max(abs(a - b) / abs(b) for a, b in zip(dp, ref))
        out = {"world": world, "global_batch": global_b, "path": "GraphedTrainStep, bucketed all-reduce overlapped", "layers": f"{layers}+{layers}", "steps": steps,
               "losses": dp, "losses_1gpu": ref, "max_rel_diff": err, "tol": tol, "ok": bool(err <= tol)}
    if world > 1:
        dist.barrier()
    return out
