from .autoregressive import AUTOREGRESSIVE_REGISTRY, build_autoregressive  # noqa: F401
from .vqvae_modules import ENCODER_REGISTRY, GENERATOR_REGISTRY, build_encoder, build_generator  # noqa: F401
from .meta_arch import META_ARCH_REGISTRY, build_model  # noqa: F401
