"""CPU oracle for the 3DAnimals reconstruction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / CPU baseline.
The product path (``3danimals_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Contents
--------
* ``reference_loader``  loads the reference's own geometry files by path from
  ``/root/reference`` (build container only; never at GPU-box run time) to pin
  the restatements below and to generate ``tests/golden/*.npz``.
* ``geometry_np``       numpy restatement of marching tets (R2), vertex normals
  (R3), bone estimation (R4), linear blend skinning (R5).
* ``torch_ref``         torch-CPU restatement with autograd of R3/R5/R8 and render_mesh (R6-R9) over the C ops.
* ``pipeline_ref``      the whole hot path on host cores (parity checker and CPU baseline of bench.py).
* ``obj_text``          line-by-line restatement of the OBJ writer (``render/obj.py:128-177``), pinned by the
  reference's own function (``tests/golden/obj_export.npz``).
* ``raster_ref.c``      C (OpenMP) restatement of the un-vendored nvdiffrast ops
  used by ``model/render/render.py`` (rasterize, interpolate, antialias) with
  forward and backward passes; built into ``oracle/_build/liboracle.so``.

Parity pinning status:
* R2, R4, R5, R8 are pinned against the reference's own Python (imported by
  path in the build container; goldens committed under tests/golden/).
* The rasterizer ops (R6, R7 interpolate, R9 antialias) restate nvdiffrast
  (NVlabs, un-pinned git dependency, absent from /root/reference) from its
  published algorithm: PARITY UNPINNED at that boundary - no reference test or
  golden vector exists for it.
"""
