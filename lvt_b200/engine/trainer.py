"""Training loop with the reference's surface (vidgen/engine/{launch,train_loop,trainer,defaults}.py):
`launch`, `default_argument_parser`, `default_setup`, `Trainer(cfg).train()`.  Host orchestration only — the
arithmetic of a step is `model(data, mode='supervised')`, `.backward()`, `optimizer.step()` exactly as in
Trainer.run_step (trainer.py:56-128)."""
import argparse
import logging
import os
import time

import torch
import torch.distributed as dist

from ..modeling.meta_arch import build_model
from ..utils import comm
from ..utils.events import EventStorage


def default_argument_parser():
    """engine/defaults.py:37-69."""
    p = argparse.ArgumentParser(description="lvt_b200 training")
    p.add_argument("--config-file", default="", metavar="FILE", help="path to config file")
    p.add_argument("--resume", action="store_true")
    p.add_argument("--eval-only", action="store_true")
    p.add_argument("--num-gpus", type=int, default=1)
    p.add_argument("--num-machines", type=int, default=1)
    p.add_argument("--machine-rank", type=int, default=0)
    p.add_argument("--dist-url", default="tcp://127.0.0.1:29531")
    p.add_argument("opts", default=None, nargs=argparse.REMAINDER, help="KEY VALUE config overrides")
    return p


def default_setup(cfg, args):
    """engine/defaults.py:72-121: output dir, logger, config dump, per-rank seed."""
    if comm.is_main_process() and cfg.OUTPUT_DIR:
        os.makedirs(cfg.OUTPUT_DIR, exist_ok=True)
        with open(os.path.join(cfg.OUTPUT_DIR, "config.yaml"), "w") as f:
            f.write(cfg.dump())
    logging.basicConfig(level=logging.INFO if comm.is_main_process() else logging.WARNING)
    seed = cfg.SEED
    if seed >= 0:
        import random
        import numpy as np
        random.seed(seed + comm.get_rank())
        np.random.seed(seed + comm.get_rank())
        torch.manual_seed(seed + comm.get_rank())


def _worker(local_rank, main_func, world_size, dist_url, args):
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", init_method=dist_url, world_size=world_size, rank=local_rank,
                            device_id=torch.device("cuda", local_rank))
    comm.synchronize()
    main_func(*args)


def launch(main_func, num_gpus_per_machine, num_machines=1, machine_rank=0, dist_url=None, args=()):
    """engine/launch.py:25-96: one process per GPU (single machine), NCCL process group."""
    if num_gpus_per_machine * num_machines > 1:
        assert num_machines == 1, "single-node data parallelism"
        import torch.multiprocessing as mp
        mp.spawn(_worker, nprocs=num_gpus_per_machine,
                 args=(main_func, num_gpus_per_machine, dist_url or "tcp://127.0.0.1:29531", args))
    else:
        main_func(*args)


class Trainer:
    """trainer.py:9-128 + train_loop.py:112-133 + defaults.py:124-363 (hooks reduced to what they do here:
    LR scheduling, periodic checkpointing, metric printing)."""

    def __init__(self, cfg, data_loader=None):
        self.cfg = cfg
        self.model = self.build_model(cfg)
        self.optimizers, self.checkpointers = self.model.configure_optimizers_and_checkpointers()
        # run_step always follows model(data, 'supervised') with loss.backward(): the models may therefore replay
        # forward + backward as one CUDA graph (LVT_TRAINER_GRAPH=0 keeps the eager launch sequence)
        if hasattr(self.model, "enable_graphed_step") and os.environ.get("LVT_TRAINER_GRAPH", "1") != "0":
            self.model.enable_graphed_step(True)
        self.data_loader = data_loader
        self._iter = iter(data_loader) if data_loader is not None else None
        if comm.get_world_size() > 1:
            self.model.wrap_parallel(device_ids=[torch.cuda.current_device()], broadcast_buffers=False)
        self.start_iter, self.max_iter = 0, cfg.SOLVER.MAX_ITER
        self.storage = None

    @classmethod
    def build_model(cls, cfg):
        model = build_model(cfg)
        logging.getLogger(__name__).info("Model:\n{}".format(type(model).__name__))
        return model

    def resume_or_load(self, resume=True):
        for c in self.checkpointers:
            c["checkpointer"].resume_or_load(c["pretrained"], resume=resume)

    def run_step(self):
        assert self.model.training, "[Trainer] model was changed to eval mode!"
        t0 = time.perf_counter()
        data = next(self._iter)
        data_time = time.perf_counter() - t0
        loss_dict = self.model(data, mode="supervised")
        losses = sum(loss_dict.values())
        losses.backward()
        if (self.iter + 1) % self.cfg.SOLVER.ACCUMULATION_STEPS == 0:
            for o in self.optimizers:
                o["optimizer"].step()
            for o in self.optimizers:
                o["optimizer"].zero_grad()
        for o in self.optimizers:
            o["scheduler"].step()
        if (self.iter + 1) % 20 == 0 or self.iter == self.start_iter:   # PeriodicWriter period (host sync)
            vals = {k: float(v.detach()) for k, v in loss_dict.items()}
            if not all(map(lambda x: x == x and abs(x) != float("inf"), vals.values())):
                raise FloatingPointError(f"Loss became infinite or NaN at iteration={self.iter}!\nloss_dict = {vals}")
            self.storage.put_scalars(data_time=data_time, total_loss=sum(vals.values()), **vals)
            if comm.is_main_process():
                logging.getLogger(__name__).info("iter %d  %s", self.iter,
                                                 "  ".join(f"{k}: {v:.4f}" for k, v in vals.items()))

    def train(self):
        self.model.train()
        period = self.cfg.SOLVER.CHECKPOINT_PERIOD
        with EventStorage(self.start_iter) as self.storage:
            for self.iter in range(self.start_iter, self.max_iter):
                self.run_step()
                if comm.is_main_process() and period > 0 and (self.iter + 1) % period == 0:
                    for c in self.checkpointers:
                        c["checkpointer"].save(f"model_{self.iter:07d}")
                self.storage.step()
            if comm.is_main_process():
                for c in self.checkpointers:
                    c["checkpointer"].save("model_final")
