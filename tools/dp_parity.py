"""Data-parallel parity on the GPU box: with a fixed GLOBAL batch and identical seeds, W ranks x (B/W)
slices with the flat-gradient all-reduce must follow the 1-GPU loss trajectory (reference: DDP + batch split,
data/build.py:62-74, meta_arch/vt.py:61-63).  Works for 1, 2, 4 and 8 ranks (global batch 16).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_parity.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from lvt_b200.utils.dp_check import run_dp_parity

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
r = run_dp_parity(rank, world, dist if world > 1 else None)
if rank == 0:
    print(f"dp_parity world={world} global_batch={r['global_batch']}: losses {['%.5f' % v for v in r['losses']]} vs 1-GPU "
          f"{['%.5f' % v for v in r['losses_1gpu']]} max rel diff {r['max_rel_diff']:.2e} -> {'OK' if r['ok'] else 'MISMATCH'}")
    assert r["ok"]
if world > 1:
    dist.destroy_process_group()
