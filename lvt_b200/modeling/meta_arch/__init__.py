from .build import META_ARCH_REGISTRY, build_model  # noqa: F401
from .vqvae import VQVAEModel  # noqa: F401
from .vt import VideoTransformerModel  # noqa: F401
