"""Data-parallel parity on the GPU box: with a fixed GLOBAL batch and identical seeds, W ranks x (B/W)
slices with the flat-gradient all-reduce must follow the 1-GPU loss trajectory (reference: DDP + batch split,
data/build.py:62-74, meta_arch/vt.py:61-63).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_parity.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from lvt_b200.data import synthetic_vt_batch
from lvt_b200.modeling.autoregressive import VTEngine, VTSpec

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
LAYERS, GLOBAL_B, STEPS = 2, 8, 4
spec = VTSpec(blocks_e=((1, 16, 16),) * LAYERS, heads_e=(8,) * LAYERS, blocks_d=((1, 16, 16),) * LAYERS, heads_d=(8,) * LAYERS)


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


batch = synthetic_vt_batch(GLOBAL_B, seed=99)          # context, slice, slice_idx, ignore_mask of the GLOBAL batch
per = GLOBAL_B // world
mine = [t[rank * per:(rank + 1) * per].contiguous() for t in batch]
eng = make_engine()
ws = eng.workspace(per, (1, 16, 16), tuple(mine[0].shape[2:]), train=True)
eng.set_inputs(ws, *mine)
hook = (lambda flat: dist.all_reduce(flat)) if world > 1 else None
dp = []
for _ in range(STEPS):
    loss = eng.train_step(ws, grad_hook=hook, grad_scale=1.0 / world).clone()
    if world > 1:
        dist.all_reduce(loss)
        loss /= world
    dp.append(loss.item())
if rank == 0:
    ref_eng = make_engine()
    ws1 = ref_eng.workspace(GLOBAL_B, (1, 16, 16), tuple(batch[0].shape[2:]), train=True)
    ref_eng.set_inputs(ws1, *batch)
    ref = [ref_eng.train_step(ws1).item() for _ in range(STEPS)]
    err = max(abs(a - b) / abs(b) for a, b in zip(dp, ref))
    print(f"dp_parity world={world} global_batch={GLOBAL_B}: losses {['%.5f' % v for v in dp]} vs 1-GPU {['%.5f' % v for v in ref]} "
          f"max rel diff {err:.2e} -> {'OK' if err <= 1e-3 else 'MISMATCH'}")
    assert err <= 1e-3
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
