"""B200-native (sm_100a) reconstruction hot path of 3DAnimals: DMTet extraction, LBS articulation, differentiable
rasterizer - hand-written CUDA behind a C-ABI library (include/b2a.h), with drop-in replacements for the reference's
model/geometry + model/render Python API.

The directory name starts with a digit, so import it by string:

    import importlib
    b2a = importlib.import_module("3danimals_b200")
    importlib.import_module("3danimals_b200.overlay").install()   # reference tree now uses the B200 kernels

Sub-modules: `ops` (autograd wrappers over libb2a.so), `geometry.dmtet`, `geometry.skinning`, `render.mesh`,
`render.render`, `nvdiffrast_shim.torch`, `overlay`, `pipeline` (synthetic end-to-end step), `synthetic`.
"""
__version__ = "0.1.0"
