"""CPU restatement of the reference's OBJ writer - TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): only tests/ may
import it; the product (3danimals_b200/render/obj.py -> b2a_obj_format) never does.

Follows write_obj, model/render/obj.py:128-177, line for line in behaviour: one formatted line per vertex / texcoord /
normal / face, numbers through '{}'.format(np.float32) (numpy hands that to float.__format__: the repr of the value widened
to double), texcoord v flipped as np.float32 `1.0 - v[1]` (:148), faces with 1-based ' a/b/c' triples whose b / c columns
are empty without texcoords / normals (:166).  Pinned against the reference's own function by tests/golden/obj_export.npz
(tests/golden/make_goldens.py obj_case).
"""
import io

import numpy as np


def obj_text(v_pos, t_pos_idx, v_nrm=None, t_nrm_idx=None, v_tex=None, t_tex_idx=None, mtl_name="mesh", save_material=True):
    f = io.StringIO()
    f.write(f"mtllib {mtl_name}.mtl\n")                                     # :132
    f.write("g default\n")                                                  # :133
    for v in v_pos:                                                         # :142-143
        f.write('v {} {} {} \n'.format(v[0], v[1], v[2]))
    if v_tex is not None and save_material:                                 # :145-148
        assert len(t_pos_idx) == len(t_tex_idx)
        for v in v_tex:
            f.write('vt {} {} \n'.format(v[0], 1.0 - v[1]))
    if v_nrm is not None:                                                   # :150-154
        assert len(t_pos_idx) == len(t_nrm_idx)
        for v in v_nrm:
            f.write('vn {} {} {}\n'.format(v[0], v[1], v[2]))
    f.write("s 1 \n")                                                       # :157-159
    f.write("g pMesh1\n")
    f.write("usemtl defaultMat\n")
    for i in range(len(t_pos_idx)):                                         # :163-167
        f.write("f ")
        for j in range(3):
            f.write(' %s/%s/%s' % (str(t_pos_idx[i][j] + 1), '' if v_tex is None else str(t_tex_idx[i][j] + 1),
                                   '' if v_nrm is None else str(t_nrm_idx[i][j] + 1)))
        f.write("\n")
    return f.getvalue().encode()


def special_float32():
    """float32 values that exercise every branch of the number format: zeros, infinities, NaN, integers, powers of two and
    their neighbours (asymmetric rounding intervals), denormals, the positional / scientific switch-over on both sides."""
    vals = [0.0, -0.0, 1.0, -1.0, 0.1, 0.5, 1e16, 9.999999e15, 1e15, 1e-4, 9.9999e-5, 1e-5, 123456792.0, 16777216.0, 16777218.0,
            3.4028235e38, 1.17549435e-38, 1e-45, 5.9e-39, np.inf, -np.inf, np.nan, 50000008.0, 0.3, 2.5, 1e22, 1e23, 8388608.5]
    out = [np.float32(v) for v in vals]
    for e in range(-149, 128):
        p = np.float32(2.0) ** np.float32(e) if e > -127 else np.float32(np.ldexp(1.0, e))
        out += [p, np.nextafter(p, np.float32(0)), np.nextafter(p, np.float32(np.inf))]
    return np.asarray(out, np.float32)
