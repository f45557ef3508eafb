import importlib, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_ops_geometry as G
pipe = importlib.import_module("3danimals_b200.pipeline")
mesh_mod = importlib.import_module("3danimals_b200.render.mesh")
sk = importlib.import_module("3danimals_b200.geometry.skinning")
dev = torch.device("cuda:0")
scene = pipe.SyntheticScene(grid_res=128, batch=16, image_res=256)
hp = pipe.HotPath(scene, dev)
hp.forward()
tets64 = hp.tets.long()
bones, chain = hp.last["bones"].detach(), hp.kinematic_chain
V = int(hp.last["inst"].v_pos.shape[1])
gen = torch.Generator(device=dev).manual_seed(11)
g_pos = torch.randn(16, V, 3, device=dev, generator=gen)
g_nrm = torch.randn(16, V, 3, device=dev, generator=gen)

def theirs(gp, gn, dt=torch.float32):
    sdf = hp.sdf.detach().to(dt).clone().requires_grad_(True)
    ang = hp.angles.detach().to(dt).clone().requires_grad_(True)
    verts, faces = G.marching_tets(hp.grid_verts.to(dt), sdf, tets64)
    posed = G.skinning(verts[None, None], bones.to(dt), chain, ang, temperature=0.05)[:, 0]
    nrm = G.auto_normals(posed, faces)
    torch.autograd.backward([posed, nrm], [gp.to(dt), gn.to(dt)])
    return sdf.grad, ang.grad

def ours(gp, gn):
    sdf = hp.sdf.detach().clone().requires_grad_(True)
    ang = hp.angles.detach().clone().requires_grad_(True)
    verts, faces, uv_idx, faces32 = hp.dmtet.extract(hp.grid_verts, sdf, hp.grid)
    prior = mesh_mod.make_mesh(verts[None], faces[None], None, uv_idx[None], None, faces_i32=faces32)
    posed, _ = sk.skinning(prior.v_pos[:, None], bones, chain, ang, output_posed_bones=True, temperature=0.05)
    inst = mesh_mod.make_mesh(posed[:, 0], prior.t_pos_idx, None, prior.t_tex_idx, None, faces_i32=prior.tri_i32())
    torch.autograd.backward([inst.v_pos, inst.v_nrm], [gp, gn])
    return sdf.grad, ang.grad

rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
z = torch.zeros_like(g_pos)
for name, gp, gn in (("pos only", g_pos, z), ("nrm only", z, g_nrm), ("both", g_pos, g_nrm)):
    t32, o, t64 = theirs(gp, gn), ours(gp, gn), theirs(gp, gn, torch.float64)
    print("%-9s d_sdf: ours vs fp64 %.2e | torch32 vs fp64 %.2e | ours vs torch32 %.2e ;  d_ang: ours vs fp64 %.2e | torch32 vs fp64 %.2e" %
          (name, rel(o[0], t64[0]), rel(t32[0], t64[0]), rel(o[0], t32[0]), rel(o[1], t64[1]), rel(t32[1], t64[1])))
