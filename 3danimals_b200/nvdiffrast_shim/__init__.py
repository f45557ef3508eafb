"""`nvdiffrast`-compatible package backed by libb2a.so (installed as `nvdiffrast` by 3danimals_b200.overlay)."""
from . import torch  # noqa: F401
