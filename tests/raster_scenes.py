"""Analytic rasterizer scenes shared by the GPU parity tests (tests/test_gpu_kernels.py), the oracle regression pins
(tests/golden/raster_scenes.npz, made by tests/golden/make_goldens.py raster_case) and tests/test_oracle_raster.py:
clip-space positions [1,V,4] + triangles [F,3] (SURVEY.md §8c "goldens to create" iv: single triangle, overlapping quads,
shared-edge fan, sub-pixel slivers, behind-camera clip, varying w)."""
import numpy as np

ANALYTIC = {
    "single_triangle": (np.array([[[-0.6, -0.5, 0.1, 1], [0.7, -0.4, 0.2, 1], [0.0, 0.8, 0.3, 1]]], np.float32), np.array([[0, 1, 2]], np.int32)),
    "overlapping_quads": (np.array([[[-0.8, -0.8, 0.5, 1], [0.4, -0.8, 0.5, 1], [0.4, 0.4, 0.5, 1], [-0.8, 0.4, 0.5, 1],
                                     [-0.3, -0.3, 0.2, 1], [0.9, -0.3, 0.2, 1], [0.9, 0.9, 0.2, 1], [-0.3, 0.9, 0.2, 1]]], np.float32),
                          np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.int32)),
    "shared_edge_fan": (np.array([[[0, 0, 0.3, 1], [0.9, 0, 0.3, 1], [0.6, 0.7, 0.3, 1], [-0.2, 0.9, 0.3, 1], [-0.8, 0.3, 0.3, 1],
                                   [-0.7, -0.6, 0.3, 1], [0.2, -0.9, 0.3, 1]]], np.float32),
                        np.array([[0, 1, 2], [0, 2, 3], [0, 3, 4], [0, 4, 5], [0, 5, 6], [0, 6, 1]], np.int32)),
    "slivers": (np.array([[[-0.9, -0.9, 0.1, 1], [0.9, -0.89, 0.1, 1], [0.9, -0.88, 0.1, 1], [-0.5, 0.1, 0.4, 1], [-0.49, 0.9, 0.4, 1],
                           [-0.48, 0.1, 0.4, 1]]], np.float32), np.array([[0, 1, 2], [3, 4, 5]], np.int32)),
    "behind_camera": (np.array([[[-0.5, -0.5, 0.2, 1.0], [0.5, -0.5, 0.2, 1.0], [0.0, 0.5, -0.8, -0.5], [0.3, 0.3, 1.5, 1.0],
                                 [0.8, 0.3, 0.5, 1.0], [0.5, 0.9, 0.5, 1.0]]], np.float32), np.array([[0, 1, 2], [3, 4, 5]], np.int32)),
    "perspective_w": (np.array([[[-1.2, -1.0, 0.4, 2.0], [1.5, -0.8, 1.0, 3.0], [0.1, 0.9, 0.2, 1.0]]], np.float32), np.array([[0, 1, 2]], np.int32)),
}

RESOLUTIONS = [(16, 16), (37, 53)]
