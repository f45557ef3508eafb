import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import pkg
from oracle import pipeline_ref as P, torch_ref as T
cuda = torch.device("cuda:0")
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda)
pipe = pkg("pipeline"); mesh_mod, render_mod = pkg("render.mesh"), pkg("render.render")
sc = pipe.SyntheticScene(grid_res=24, batch=2, image_res=64, sdf_noise=0.0)
ref = P.forward(sc)
B, r = sc.batch, sc.image_res
posed = ref["posed"][:, 0].detach().clone()
faces = ref["faces"]
mvp_t, w2c_t, cam_t = (torch.from_numpy(x) for x in (sc.mvp, sc.w2c, sc.campos))
ones = lambda gb_tex, cam_normal, gbuf: {"shaded": torch.ones_like(gb_tex)}
out_ref = T.render_mesh(posed, T.auto_normals(posed, faces), faces, mvp_t, w2c_t, cam_t, ones, (r, r), render_modes=("shaded",))
inst = mesh_mod.make_mesh(dev(posed.numpy()), dev(faces.numpy())[None], None, None, None)
out, = render_mod.render_mesh(None, inst, dev(sc.mvp), dev(sc.w2c), dev(sc.campos), None, None, (r, r), render_modes=["shaded"], bsdf="diffuse")
a, b = out.cpu().numpy(), out_ref["shaded"].numpy()
d = np.abs(a - b)
print("max diff", d.max(), "n diff", (d > 0).sum(), "of", d.size)
idx = np.argwhere(d > 0)[:10]
for i in idx: print(i, a[tuple(i)], b[tuple(i)])
