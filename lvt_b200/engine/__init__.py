from .checkpoint import Checkpointer  # noqa: F401
from .trainer import Trainer, default_argument_parser, default_setup, launch  # noqa: F401
