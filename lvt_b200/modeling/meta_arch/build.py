"""META_ARCH registry (reference vidgen/modeling/meta_arch/build.py:4-19)."""
from ...utils.registry import Registry

META_ARCH_REGISTRY = Registry("META_ARCH")


def build_model(cfg):
    """Build the whole model named by cfg.MODEL.META_ARCHITECTURE (no weights are loaded)."""
    return META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
