from .vt_engine import VTEngine, VTSpec, VTWorkspace  # noqa: F401
