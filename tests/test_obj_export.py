"""Mesh export: `write_obj` through libb2a.so (`b2a_obj_format`, host threads) produces the reference's
file BYTE FOR BYTE.  Pinned three ways: the committed output of the reference's own `write_obj` (tests/golden/obj_export.npz,
made by tests/golden/make_goldens.py obj_case), the line-by-line restatement in oracle/obj_text.py on larger seeded meshes, and
Python's own `repr` on a few million float32 bit patterns (the number format is the part worth breaking).  Host-only: no GPU."""
import contextlib
import ctypes
import io
import os
import types

import numpy as np
import pytest

from conftest import golden, pkg
from oracle import obj_text as oracle_obj
from oracle import reference_loader


def _mesh_arrays(rng, V, Vt, F, special=True):
    v_pos = (rng.standard_normal((V, 3)) * 2).astype(np.float32)
    if special:
        sp = oracle_obj.special_float32()
        n = min(len(sp) // 3, V)
        v_pos[:n] = sp[:3 * n].reshape(n, 3)
    v_nrm = rng.standard_normal((V, 3)).astype(np.float32)
    v_tex = rng.random((Vt, 2)).astype(np.float32)
    t_pos = rng.integers(0, V, (F, 3))
    t_tex = rng.integers(0, Vt, (F, 3))
    return v_pos, v_nrm, v_tex, t_pos, t_tex


def test_golden_bytes_of_the_reference_writer():
    g = golden("obj_export.npz")
    obj = pkg("render.obj")
    full = obj.obj_text(g["v_pos"], g["t_pos"], g["v_nrm"], g["t_pos"], g["v_tex"], g["t_tex"], mtl_name="golden_full")
    assert full.tobytes() == g["text_full"].tobytes()
    nomat = obj.obj_text(g["v_pos"], g["t_pos"], g["v_nrm"], g["t_pos"], g["v_tex"], g["t_tex"], mtl_name="golden_nomat", write_texcoords=False)
    assert nomat.tobytes() == g["text_nomat"].tobytes()
    bare = obj.obj_text(g["v_pos"], g["t_pos"], mtl_name="golden_bare")
    assert bare.tobytes() == g["text_bare"].tobytes()
    # the restatement is pinned by the same vectors
    assert oracle_obj.obj_text(g["v_pos"], g["t_pos"], g["v_nrm"], g["t_pos"], g["v_tex"], g["t_tex"], mtl_name="golden_full") == g["text_full"].tobytes()
    assert oracle_obj.obj_text(g["v_pos"], g["t_pos"], g["v_nrm"], g["t_pos"], g["v_tex"], g["t_tex"], mtl_name="golden_nomat",
                               save_material=False) == g["text_nomat"].tobytes()
    assert oracle_obj.obj_text(g["v_pos"], g["t_pos"], mtl_name="golden_bare") == g["text_bare"].tobytes()


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_matches_restatement_on_seeded_meshes(threads):
    """Several chunk boundaries (2048 lines per chunk), every layout, any thread count: same bytes."""
    obj = pkg("render.obj")
    rng = np.random.default_rng(7 + threads)
    v_pos, v_nrm, v_tex, t_pos, t_tex = _mesh_arrays(rng, 5000, 4100, 9000)
    for kw in (dict(v_nrm=v_nrm, t_nrm_idx=t_pos, v_tex=v_tex, t_tex_idx=t_tex), dict(v_nrm=v_nrm, t_nrm_idx=t_pos), dict(v_tex=v_tex, t_tex_idx=t_tex), dict()):
        ours = obj.obj_text(v_pos, t_pos, mtl_name="m_%d" % threads, threads=threads, **kw).tobytes()
        assert ours == oracle_obj.obj_text(v_pos, t_pos, mtl_name="m_%d" % threads, **kw)
    ours = obj.obj_text(v_pos, t_pos, v_nrm, t_pos, v_tex, t_tex, write_texcoords=False, threads=threads).tobytes()
    assert ours == oracle_obj.obj_text(v_pos, t_pos, v_nrm, t_pos, v_tex, t_tex, save_material=False)


def test_empty_and_tiny_meshes():
    obj = pkg("render.obj")
    e3, e2, ei = np.zeros((0, 3), np.float32), np.zeros((0, 2), np.float32), np.zeros((0, 3), np.int64)
    assert obj.obj_text(e3, ei, mtl_name="").tobytes() == oracle_obj.obj_text(e3, ei, mtl_name="") == b"mtllib .mtl\ng default\ns 1 \ng pMesh1\nusemtl defaultMat\n"
    assert obj.obj_text(e3, ei, e3, ei, e2, ei).tobytes() == oracle_obj.obj_text(e3, ei, e3, ei, e2, ei)
    v = np.asarray([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    f = np.asarray([[0, 1, 2]], np.int32)            # int32 indices are accepted like any integer array
    assert obj.obj_text(v, f).tobytes() == b"mtllib mesh.mtl\ng default\nv 0.0 0.0 0.0 \nv 1.0 0.0 0.0 \nv 0.0 1.0 0.0 \ns 1 \ng pMesh1\nusemtl defaultMat\nf  1// 2// 3//\n"


def test_number_format_is_python_repr_of_the_widened_double():
    """2 M random float32 bit patterns + every power of two and its neighbours + the texcoord flip, against repr()."""
    obj = pkg("render.obj")
    rng = np.random.default_rng(11)
    vals = rng.integers(0, 2 ** 32, size=3 * 700_000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    sp = oracle_obj.special_float32()
    vals[:len(sp)] = sp
    text = obj.obj_text(vals.reshape(-1, 3), np.zeros((0, 3), np.int64)).tobytes().decode()
    lines = text.split("\n")[2:-4]
    assert len(lines) == len(vals) // 3
    with np.errstate(all="ignore"):
        want = [repr(x) for x in vals.astype(np.float64).tolist()]
    got = [tok for ln in lines for tok in ln.split(" ")[1:4]]
    assert got == want
    # texcoords: u as is, v flipped in float32 (obj.py:148)
    uv = vals[: 2 * 200_000].reshape(-1, 2).copy()
    text = obj.obj_text(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int64), v_tex=uv, t_tex_idx=np.zeros((0, 3), np.int64)).tobytes().decode()
    got = [tok for ln in text.split("\n")[2:-4] for tok in ln.split(" ")[1:3]]
    with np.errstate(all="ignore"):
        flipped = np.stack([uv[:, 0], np.float32(1.0) - uv[:, 1]], -1).astype(np.float64)
    assert got == [repr(x) for x in flipped.ravel().tolist()]


def test_errors_and_shape_checks():
    obj = pkg("render.obj")
    lib_mod = pkg("_lib")
    h = lib_mod.lib()
    v = np.zeros((4, 3), np.float32)
    f = np.zeros((2, 3), np.int64)
    with pytest.raises(ValueError):
        obj.obj_text(np.zeros((4, 2), np.float32), f)
    with pytest.raises(AssertionError):                     # the reference's own asserts (obj.py:146,151)
        obj.obj_text(v, f, v_nrm=v, t_nrm_idx=np.zeros((3, 3), np.int64))
    with pytest.raises(AssertionError):
        obj.obj_text(v, f, v_tex=np.zeros((4, 2), np.float32), t_tex_idx=np.zeros((1, 3), np.int64))
    # C-ABI: a buffer below the bound is refused with a message, nothing is written
    bound, written = ctypes.c_size_t(), ctypes.c_size_t(123)
    assert h.b2a_obj_text_bound(4, 0, 0, 2, 4, ctypes.byref(bound)) == 0 and bound.value >= 4 * 40 + 2 * 60
    out = np.full(16, 7, np.uint8)
    name = ctypes.create_string_buffer(b"mesh")
    rc = h.b2a_obj_format(v.ctypes.data, 4, None, 0, None, 0, f.ctypes.data, None, None, 2, ctypes.cast(name, ctypes.c_void_p), 4,
                          out.ctypes.data, out.size, ctypes.byref(written), 1)
    assert rc != 0 and b"smaller than the bound" in h.b2a_last_error_string() and (out == 7).all() and written.value == 123
    assert h.b2a_obj_text_bound(-1, 0, 0, 0, 0, ctypes.byref(bound)) != 0


def _namespace_mesh(v_pos, v_nrm, v_tex, t_pos, t_tex, B=2):
    import torch
    T = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a))
    rep = lambda a, s: None if a is None else torch.stack([T(a) * (1 + i * s) for i in range(B)])
    return types.SimpleNamespace(v_pos=rep(v_pos, 0.5), v_nrm=rep(v_nrm, 0.0), v_tex=rep(v_tex, 0.0), t_pos_idx=T(t_pos)[None],
                                 t_nrm_idx=T(t_pos)[None] if v_nrm is not None else None, t_tex_idx=T(t_tex)[None] if v_tex is not None else None,
                                 material=None)


@pytest.mark.parametrize("save_material", [True, False])
def test_write_obj_file_and_console_lines(tmp_path, save_material):
    """The drop-in `write_obj`: instance `idx` of a batched mesh, same file name, same bytes, same console lines as the reference
    (run side by side when the reference tree is present; against the restatement otherwise)."""
    obj = pkg("render.obj")
    rng = np.random.default_rng(3)
    v_pos, v_nrm, v_tex, t_pos, t_tex = _mesh_arrays(rng, 700, 650, 1300)
    mesh = _namespace_mesh(v_pos, v_nrm, v_tex, t_pos, t_tex)
    ours_dir = tmp_path / "ours"
    ours_dir.mkdir()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        obj.write_obj(str(ours_dir), "animal_1", mesh, 1, save_material=save_material)
    data = (ours_dir / "animal_1.obj").read_bytes()
    assert sorted(os.listdir(ours_dir)) == ["animal_1.obj"]             # material is None: no .mtl (obj.py:169)
    want = oracle_obj.obj_text(mesh.v_pos[1].numpy(), t_pos, v_nrm, t_pos, v_tex, t_tex, mtl_name="animal_1", save_material=save_material)
    assert data == want
    if reference_loader.available():
        ref_dir = tmp_path / "ref"
        ref_dir.mkdir()
        rbuf = io.StringIO()
        with contextlib.redirect_stdout(rbuf), contextlib.redirect_stderr(io.StringIO()):
            reference_loader.reference_write_obj()(str(ref_dir), "animal_1", mesh, 1, save_material=save_material)
        assert (ref_dir / "animal_1.obj").read_bytes() == data
        assert buf.getvalue().replace(str(ours_dir), "D") == rbuf.getvalue().replace(str(ref_dir), "D")


def test_overlay_serves_write_obj():
    """`from ..render.obj import write_obj` (model/utils/misc.py:12) binds to this package once the overlay is installed."""
    overlay = pkg("overlay")
    overlay.install()
    try:
        spec = overlay._finder.find_spec("model.render.obj")
        mod = spec.loader.create_module(spec)
        assert mod is pkg("render.obj") and callable(mod.write_obj)
        with pytest.raises(AttributeError):         # standalone (no reference tree imported): only write_obj exists
            mod.no_such_name
    finally:
        overlay.uninstall()
