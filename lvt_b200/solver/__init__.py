from .build import build_lr_scheduler, build_optimizer  # noqa: F401
