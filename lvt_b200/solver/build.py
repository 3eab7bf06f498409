"""Optimizers / LR schedulers with the reference's builder surface (vidgen/solver/build.py:46-105,
lr_scheduler.py:17-117).  The update itself is ONE fused kernel over the engine's flat buffers
(lvt_rmsprop_step / lvt_adam_step, torch.optim semantics, which also refreshes the bf16 shadow weights);
the reference's one-param-group-per-tensor layout (build.py:12-43) has identical hyper-parameters in
every group for the shipped configs (all weight decays 0), so a flat update is equivalent."""
import math
from bisect import bisect_right


class EngineOptimizer:
    """torch.optim.Optimizer look-alike bound to a VTEngine / VQVAEEngine."""

    def __init__(self, engine, name, lr, **hyper):
        self.engine, self.name = engine, name
        self.param_groups = [dict(lr=lr, initial_lr=lr, **hyper)]
        if name == "rmsprop":
            engine.init_optimizer("rmsprop", lr=lr, alpha=hyper["alpha"], momentum=hyper["momentum"], eps=1e-8)
        elif hasattr(engine, "spec") and engine.__class__.__name__ == "VQVAEEngine":
            engine.init_optimizer(lr=lr, betas=hyper["betas"], eps=1e-8)
        else:
            engine.init_optimizer("adam", lr=lr, betas=hyper["betas"], eps=1e-8)

    def step(self, closure=None):
        self.engine.opt["lr"] = self.param_groups[0]["lr"]
        self.engine.optimizer_step()

    def zero_grad(self, set_to_none=False):
        self.engine.store.grad.zero_()   # Parameter.grad views stay attached

    def state_dict(self):
        return {"param_groups": self.param_groups, "step": self.engine.opt["step"]}


class SharedStepOptimizer:
    """Second handle on an engine whose parameters are already updated by another EngineOptimizer
    (the reference steps netE and netG separately, trainer.py:83-87; here they share one launch)."""

    def __init__(self, primary):
        self.param_groups = [dict(primary.param_groups[0])]

    def step(self, closure=None):
        pass

    def zero_grad(self, set_to_none=False):
        pass

    def state_dict(self):
        return {"param_groups": self.param_groups}


def build_optimizer(model, cfg, suffix=""):
    """cfg.SOLVER.{OPTIMIZER_NAME, LR*, ADAM.*, RMSPROP.*} -> optimizer (build.py:46-74)."""
    lr = cfg.SOLVER["LR" + suffix]
    for k in ("BASE", "NORM", "BIAS"):
        if cfg.SOLVER.WEIGHT_DECAY[k + suffix] != 0.0:
            raise NotImplementedError("weight decay is 0 in every shipped config; the fused step omits it")
    engine = model.engine
    name = cfg.SOLVER.OPTIMIZER_NAME
    if getattr(engine, "_optimizer", None) is not None:
        return SharedStepOptimizer(engine._optimizer)
    if name == "adam":
        opt = EngineOptimizer(engine, "adam", lr, betas=(cfg.SOLVER.ADAM["BETA1" + suffix], cfg.SOLVER.ADAM["BETA2" + suffix]))
    elif name == "rmsprop":
        opt = EngineOptimizer(engine, "rmsprop", lr, alpha=cfg.SOLVER.RMSPROP["ALPHA" + suffix],
                              momentum=cfg.SOLVER.RMSPROP["MOMENTUM" + suffix])
    else:
        raise ValueError("Unknown optimizer: {}".format(name))
    engine._optimizer = opt
    return opt


class _Scheduler:
    def __init__(self, optimizer):
        self.optimizer, self.last_epoch = optimizer, 0
        self.base_lrs = [g["initial_lr"] for g in optimizer.param_groups]

    def factor(self, it):
        return 1.0

    def get_last_lr(self):
        return [g["lr"] for g in self.optimizer.param_groups]

    def _apply(self, it):
        for g, base in zip(self.optimizer.param_groups, self.base_lrs):
            g["lr"] = base * self.factor(it)

    def step(self):
        self.last_epoch += 1
        self._apply(self.last_epoch)


def _warmup(method, it, warmup_iters, warmup_factor):
    if it >= warmup_iters:
        return 1.0
    if method == "constant":
        return warmup_factor
    alpha = it / warmup_iters
    return warmup_factor * (1 - alpha) + alpha


class WarmupMultiStepLR(_Scheduler):
    def __init__(self, optimizer, milestones, gamma, warmup_factor, warmup_iters, warmup_method):
        super().__init__(optimizer)
        self.m, self.gamma, self.wf, self.wi, self.wm = list(milestones), gamma, warmup_factor, warmup_iters, warmup_method
        self._apply(0)  # torch's _LRScheduler does an initial step: iteration 0 runs at base_lr * factor(0)

    def factor(self, it):
        return _warmup(self.wm, it, self.wi, self.wf) * self.gamma ** bisect_right(self.m, it)


class WarmupCosineLR(_Scheduler):
    def __init__(self, optimizer, max_iters, warmup_factor, warmup_iters, warmup_method):
        super().__init__(optimizer)
        self.max_iters, self.wf, self.wi, self.wm = max_iters, warmup_factor, warmup_iters, warmup_method
        self._apply(0)

    def factor(self, it):
        return _warmup(self.wm, it, self.wi, self.wf) * 0.5 * (1.0 + math.cos(math.pi * it / self.max_iters))


def build_lr_scheduler(cfg, optimizer):
    """cfg.SOLVER.LR_SCHEDULER_NAME -> scheduler (build.py:77-105)."""
    name = cfg.SOLVER.LR_SCHEDULER_NAME
    if name == "WarmupMultiStepLR":
        return WarmupMultiStepLR(optimizer, cfg.SOLVER.STEPS, cfg.SOLVER.GAMMA, cfg.SOLVER.WARMUP_FACTOR,
                                 cfg.SOLVER.WARMUP_ITERS, cfg.SOLVER.WARMUP_METHOD)
    if name == "WarmupCosineLR":
        return WarmupCosineLR(optimizer, cfg.SOLVER.MAX_ITER, cfg.SOLVER.WARMUP_FACTOR, cfg.SOLVER.WARMUP_ITERS,
                              cfg.SOLVER.WARMUP_METHOD)
    if name == "Identity":
        return _Scheduler(optimizer)
    raise ValueError("Unknown LR scheduler: {}".format(name))
