from .vt_engine import GraphedTrainStep, VTEngine, VTSpec, VTWorkspace  # noqa: F401
from .videotransformer import (AUTOREGRESSIVE_REGISTRY, Autoregressive, VideoTransformer,  # noqa: F401
                               build_autoregressive)
