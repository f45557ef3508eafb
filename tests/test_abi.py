"""The C-ABI boundary (include/b2a.h <-> libb2a.so) without a GPU: every declared symbol is exported, the header parser
that derives the ctypes prototypes sees them all, and the host-only entry points (version, workspace sizes, error
reporting) behave.  No compute call is made here."""
import ctypes
import os
import subprocess

from conftest import ROOT, pkg

EXPECTED = {
    "b2a_articulation_constraints_fwd", "b2a_articulation_constraints_bwd", "b2a_composite_up_pool_fwd", "b2a_composite_up_pool_bwd", "b2a_mlp_pack_weights_many", "b2a_rows_gather", "b2a_rows_scatter", "b2a_p2p_alloc", "b2a_p2p_open", "b2a_p2p_close", "b2a_p2p_free", "b2a_allreduce_p2p",
    "b2a_version", "b2a_last_error_string", "b2a_mt_workspace_bytes", "b2a_mt_tile_shape", "b2a_mt_count", "b2a_mt_emit", "b2a_mt_bwd",
    "b2a_estimate_bones_workspace_bytes", "b2a_estimate_bones", "b2a_lbs_bone_transforms", "b2a_lbs_fwd", "b2a_lbs_bwd", "b2a_lbs_bone_transforms_bwd", "b2a_vertex_normals_fwd",
    "b2a_vertex_normals_bwd", "b2a_xfm_points_fwd", "b2a_xfm_points_bwd", "b2a_rasterize_workspace_bytes", "b2a_rasterize_fwd",
    "b2a_rasterize_bwd", "b2a_interpolate_fwd", "b2a_interpolate_bwd", "b2a_edge_adjacency_workspace_bytes", "b2a_edge_adjacency",
    "b2a_antialias_workspace_bytes", "b2a_antialias_prepare", "b2a_antialias_fwd", "b2a_antialias_bwd", "b2a_antialias_pair_fwd", "b2a_antialias_pair_bwd", "b2a_composite_up_fwd", "b2a_composite_up_bwd", "b2a_gbuffer_pack_bytes", "b2a_gbuffer_fwd", "b2a_gbuffer_bwd_workspace_bytes", "b2a_gbuffer_bwd",
    "b2a_render_geometry_fwd", "b2a_render_geometry_bwd",
    "b2a_mt_tables_workspace_bytes", "b2a_mt_build_edges", "b2a_mt_emit_edges", "b2a_mt_build_tile_words",
    "b2a_mlp_packed_bytes", "b2a_mlp_pack_weights", "b2a_mlp_rows_gemm", "b2a_mlp_wgrad", "b2a_mlp_embed_fwd", "b2a_mlp_embed_bwd", "b2a_mlp_colsum_segments",
    "b2a_shade_directional_fwd", "b2a_shade_directional_bwd", "b2a_analytic_field_fwd", "b2a_analytic_field_bwd", "b2a_obj_text_bound", "b2a_obj_format",
}


def test_header_declares_expected_surface():
    protos = pkg("_lib").parse_header()
    assert set(protos) == EXPECTED
    # no torch / C++ types cross the boundary: plain pointers, sizes, scalars, an opaque stream handle
    import re
    text = open(os.path.join(ROOT, "include", "b2a.h")).read()
    code = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)          # comments cite nvdiffrast.torch; the declarations must not
    assert "torch" not in code and "at::" not in code and "std::" not in code and 'extern "C"' in code
    rt, params = protos["b2a_antialias_bwd"]
    assert rt is ctypes.c_int and [p[1] for p in params][-1] == "stream"


def test_library_exports_every_declared_symbol():
    lib_mod = pkg("_lib")
    handle = lib_mod.lib()          # raises if the .so is missing or lacks any symbol declared in the header
    exported = subprocess.run(["nm", "-D", "--defined-only", lib_mod.LIB_PATH], capture_output=True, text=True).stdout
    names = {line.split()[-1] for line in exported.splitlines() if " T " in line}
    assert EXPECTED <= names
    assert {n for n in names if n.startswith("b2a_")} == EXPECTED       # and nothing undeclared leaks out
    assert handle.b2a_version() == 100


def test_workspace_sizes_and_error_convention():
    lib_mod = pkg("_lib")
    h = lib_mod.lib()
    out = ctypes.c_size_t(0)
    assert h.b2a_rasterize_workspace_bytes(16, 50000, 256, 256, ctypes.byref(out)) == 0
    assert out.value >= 16 * 256 * 256 * 8
    assert h.b2a_mt_workspace_bytes(1000, 7000, 6000, ctypes.byref(out)) == 0 and out.value > 7000 * 4
    assert h.b2a_edge_adjacency_workspace_bytes(100, ctypes.byref(out)) == 0 and out.value >= 1024 * 16
    # errors: non-zero return + thread-local message, Python side raises
    assert h.b2a_rasterize_workspace_bytes(0, 1, 1, 1, ctypes.byref(out)) != 0
    assert b"invalid argument" in h.b2a_last_error_string()
    try:
        lib_mod.check(2)
        raise AssertionError("check() must raise")
    except lib_mod.B2AError as e:
        assert "invalid argument" in str(e)


def test_sass_is_sm100a():
    """The shipped library carries sm_100a SASS only (no multi-arch fallback)."""
    lib_mod = pkg("_lib")
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "-lelf", lib_mod.LIB_PATH], capture_output=True, text=True).stdout
    archs = {tok for line in out.splitlines() for tok in line.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


def test_fastcall_binding_covers_the_compute_entry_points():
    """The generated METH_FASTCALL module (build.py build_fastcall) wraps every status-returning entry point whose
    arguments are plain ints / floats / pointers, calls the same library instance (shared error string) and type-checks."""
    lib_mod = pkg("_lib")
    fast = lib_mod.fast()
    assert fast is not None, "fastcall extension missing (needs Python.h + gcc at build time)"
    protos = lib_mod.parse_header()
    cold = {n for n, (rt, ps) in protos.items() if rt is not ctypes.c_int or any(p[0] in (ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int)) for p in ps)}
    assert {n for n in dir(fast) if n.startswith("b2a_")} == set(protos) - cold
    n = len(protos["b2a_xfm_points_fwd"][1])
    assert fast.b2a_xfm_points_fwd(*([None] * 2 + [1, 1, 8] + [None] * (n - 5))) != 0      # null pointers -> status, no launch
    assert b"invalid argument" in lib_mod.lib().b2a_last_error_string()
    try:
        fast.b2a_xfm_points_fwd(1, 2)
        raise AssertionError("wrong arity must raise")
    except TypeError:
        pass
