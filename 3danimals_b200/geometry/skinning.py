"""Drop-in for the reference's model/geometry/skinning.py: estimate_bones, skinning, euler_angles_to_matrix.

`skinning` runs in the sm_100a kernels of libb2a.so (csrc/lbs.cu): bone transforms for the whole batch in two tiny
launches, then ONE fused kernel over all (image, vertex) pairs that recomputes the soft bone weights in registers
(no [K,B,F,V] weight tensor unless the caller reads aux['vertices_to_bones']).  The reference launches
K x chain-depth x ~15 micro-kernels (skinning.py:389-429).

`estimate_bones` (a @no_grad heuristic, skinning.py:49-248) keeps the reference's semantics with device-side masked
arg-min instead of per-(b,f) Python loops; the only host read left is the leg attachment index, needed to build the
Python kinematic-chain lists when `compute_kinematic_chain=True` (once per epoch in MagicPony).
"""
import math
import os

import torch

from .. import ops


# ---------------------------------------------------------------------------------------------------------------
# kinematic chains (pure Python structures, reference skinning.py:25-46, :111-131)
# ---------------------------------------------------------------------------------------------------------------
def build_kinematic_chain(n_bones, start_bone_idx):
    bones_to_joints, chain, dependents = [], [], []
    for i in range(n_bones):
        bones_to_joints.append((i + 1, i))
        chain = [(start_bone_idx + i, dependents)] + chain  # parents stay in front
        dependents = dependents + [start_bone_idx + i]
    return bones_to_joints, chain, dependents


def update_body_kinematic_chain(kinematic_chain, leg_kinematic_chain, body_bone_idx, leg_bone_idxs, attach_legs_to_body=True):
    if attach_legs_to_body:
        for bone_idx, dependents in kinematic_chain:
            if bone_idx == body_bone_idx or body_bone_idx in dependents:
                dependents += leg_bone_idxs
    return kinematic_chain + leg_kinematic_chain


def children_to_parents(kinematic_tree):
    return [(bone_id, [p for p, children in kinematic_tree if bone_id in children]) for bone_id, _ in kinematic_tree]


def _joints_to_bones(joints, bones_idxs):
    a = torch.tensor([i for i, _ in bones_idxs], device=joints.device)
    b = torch.tensor([j for _, j in bones_idxs], device=joints.device)
    return torch.stack([joints[:, :, a], joints[:, :, b]], dim=3)  # [B,F,K,2,3]


def _take_vertex(seq_shape, idx):
    return seq_shape.gather(2, idx[..., None, None].expand(-1, -1, 1, 3)).squeeze(2)


def _body_chain(n_body_bones):
    """Body part of the kinematic chain and bone->joint table (reference skinning.py:111-131)."""
    half = n_body_bones // 2
    bones_to_joints, kinematic_chain, dependents, bone_idx = [], [], [], 0
    for i in range(half):
        bones_to_joints.append((i + 1, i))
        kinematic_chain = [(bone_idx, dependents)] + kinematic_chain
        dependents = dependents + [bone_idx]
        bone_idx += 1
    dependents = []
    for i in range(n_body_bones - 1, half - 1, -1):
        bones_to_joints.append((i, i + 1))
        kinematic_chain = [(bone_idx, dependents)] + kinematic_chain
        dependents = dependents + [bone_idx]
        bone_idx += 1
    return bones_to_joints, kinematic_chain


_MODES = {"z_minmax": 0, "z_minmax_y+": 1}
# B2A_FUSED_FAUNA_BONES=0: 3D-Fauna's bone_y_threshold variant takes the device-side torch formulation instead of the kernel
FUSED_FAUNA_VARIANT = os.environ.get("B2A_FUSED_FAUNA_BONES", "1") != "0"


def _estimate_bones_fused(seq_shape, n_body_bones, n_leg_bones, body_bones_mode, compute_kinematic_chain, aux, attach_legs_to_body,
                          legs_to_body_joint_indices, bone_y_threshold=None):
    """The sm_100a path (csrc/estimate_bones.cu): steady state = 1 memset + 4 launches (7 for 3D-Fauna's bone_y_threshold
    variant), no host sync.  Building the
    Python kinematic-chain lists (once per epoch in MagicPony, InstancePredictorBase.py:319-335) needs the leg attachment
    indices on the host: one 16-byte read."""
    mode = _MODES[body_bones_mode]
    if not compute_kinematic_chain:
        attach = [leg["body_bone_idx"] for leg in aux["legs"]] if n_leg_bones > 0 else (-1, -1, -1, -1)
        return ops.estimate_bones(seq_shape, n_body_bones, n_leg_bones, mode, attach, bone_y_threshold=bone_y_threshold)
    aux = {}
    bones_to_joints, kinematic_chain = _body_chain(n_body_bones)
    aux["bones_to_joints"] = bones_to_joints
    if n_leg_bones > 0:
        cfg = legs_to_body_joint_indices if legs_to_body_joint_indices is not None else [None, None, None, None]
        first = [-1 if cfg[i] is None else int(cfg[i]) for i in range(2)]
        if -1 in first:
            _, att = ops.estimate_bones(seq_shape, n_body_bones, n_leg_bones, mode, first + [-1, -1], want_attach=True,
                                        bone_y_threshold=bone_y_threshold)
            first = att[:2].tolist()
        # legs 2 and 3 reuse the joints chosen for legs 1 and 0 (skinning.py:214-217)
        attach = [first[0], first[1], first[1], first[0]]
        for i in range(4):
            cfg[i] = attach[i]
        start_bone_idx, leg_auxs = n_body_bones, []
        for i in range(4):
            leg_b2j, leg_chain, leg_ids = build_kinematic_chain(n_leg_bones, start_bone_idx)
            kinematic_chain = update_body_kinematic_chain(kinematic_chain, leg_chain, attach[i], leg_ids, attach_legs_to_body)
            leg_auxs.append({"body_bone_idx": attach[i], "leg_bones_to_joints": leg_b2j})
            start_bone_idx += n_leg_bones
        aux["legs"] = leg_auxs
    else:
        attach = (-1, -1, -1, -1)
    aux["kinematic_chain"] = kinematic_chain
    bones = ops.estimate_bones(seq_shape, n_body_bones, n_leg_bones, mode, attach, bone_y_threshold=bone_y_threshold)
    return bones, kinematic_chain, aux


@torch.no_grad()
def estimate_bones(seq_shape, n_body_bones, resample=False, n_legs=4, n_leg_bones=0, body_bones_mode="z_minmax",
                   compute_kinematic_chain=True, aux=None, attach_legs_to_body=True, legs_to_body_joint_indices=None,
                   bone_y_threshold=None):
    """seq_shape [B,F,V,3] -> bones [B,F,K,2,3] (+ kinematic_chain, aux when compute_kinematic_chain)."""
    if resample:
        raise NotImplementedError("resample=True is never set by any caller of the reference (SURVEY.md §2 #5)")
    assert n_body_bones % 2 == 0
    if n_leg_bones > 0:
        assert n_legs == 4
    if not seq_shape.is_cuda:
        raise ops._lib.B2AError("estimate_bones needs CUDA tensors (the B200 hot path has no CPU fallback)")
    fused_ok = body_bones_mode in _MODES and (bone_y_threshold is None or (FUSED_FAUNA_VARIANT and 0.0 < float(bone_y_threshold) <= 1.0))
    if fused_ok:
        return _estimate_bones_fused(seq_shape, n_body_bones, n_leg_bones, body_bones_mode, compute_kinematic_chain, aux,
                                     attach_legs_to_body, legs_to_body_joint_indices, bone_y_threshold)
    return _estimate_bones_torch(seq_shape, n_body_bones, n_leg_bones, body_bones_mode, compute_kinematic_chain, aux, attach_legs_to_body,
                                 legs_to_body_joint_indices, bone_y_threshold)


def _estimate_bones_torch(seq_shape, n_body_bones, n_leg_bones, body_bones_mode, compute_kinematic_chain, aux, attach_legs_to_body,
                          legs_to_body_joint_indices, bone_y_threshold):
    """Device-side torch formulation of both variants (Fauna's bone_y_threshold one: InstancePredictorFauna.py:20,90, seven
    masked quantiles).  The fused kernel covers both; this formulation remains as the switchable alternative
    (B2A_FUSED_FAUNA_BONES=0) and for body_bones_mode values outside the kernel's.  The public entry point only passes CUDA tensors; the
    formulation itself is device-agnostic, which is how tests/test_host_logic.py pins it against the reference's goldens
    without a GPU."""
    n_legs = 4
    zs_all = seq_shape[..., 2]
    if body_bones_mode == "z_minmax":
        point_a = _take_vertex(seq_shape, zs_all.argmax(dim=2))
        point_b = _take_vertex(seq_shape, zs_all.argmin(dim=2))
    elif body_bones_mode == "z_minmax_y+":
        mid = seq_shape.mean(2)
        upper = (seq_shape[..., 1] > (mid[:, :, None, 1] - 0.5)).float()
        point_a = _take_vertex(seq_shape, (zs_all * upper + (-1e6) * (1 - upper)).argmax(2))
        point_b = _take_vertex(seq_shape, (zs_all * upper + 1e6 * (1 - upper)).argmin(2))
    else:
        raise NotImplementedError
    point_a = point_a.clone(); point_b = point_b.clone()
    point_a[..., 0] = 0
    point_b[..., 0] = 0
    mid_point = seq_shape.mean(2)
    mid_point[..., 0] = 0
    if n_leg_bones > 0:
        mid_point[..., 1] += 0.5
    assert n_body_bones % 2 == 0
    n_joints = n_body_bones + 1
    blend = torch.linspace(0., 1., math.ceil(n_joints / 2), device=seq_shape.device)[None, None, :, None]
    joints_a = point_a[:, :, None, :] * (1 - blend) + mid_point[:, :, None, :] * blend
    joints_b = point_b[:, :, None, :] * blend + mid_point[:, :, None, :] * (1 - blend)
    joints = torch.cat([joints_a[:, :, :-1], joints_b], 2)

    if compute_kinematic_chain:
        aux = {}
        half = n_body_bones // 2
        bones_to_joints, kinematic_chain, dependents, bone_idx = [], [], [], 0
        for i in range(half):
            bones_to_joints.append((i + 1, i))
            kinematic_chain = [(bone_idx, dependents)] + kinematic_chain
            dependents = dependents + [bone_idx]
            bone_idx += 1
        dependents = []
        for i in range(n_body_bones - 1, half - 1, -1):
            bones_to_joints.append((i, i + 1))
            kinematic_chain = [(bone_idx, dependents)] + kinematic_chain
            dependents = dependents + [bone_idx]
            bone_idx += 1
        aux["bones_to_joints"] = bones_to_joints
    else:
        bones_to_joints = aux["bones_to_joints"]
        kinematic_chain = aux["kinematic_chain"]
    bones_pred = _joints_to_bones(joints, bones_to_joints)

    if n_leg_bones > 0:
        assert n_legs == 4
        xs, ys, zs = seq_shape.unbind(-1)
        if bone_y_threshold is None:
            x_margin = (xs.quantile(0.95) - xs.quantile(0.05)) * 0.2
            quadrants = [(xs > x_margin) & (zs > 0), (xs > x_margin) & (zs < 0), (xs < -x_margin) & (zs < 0), (xs < -x_margin) & (zs > 0)]
        else:
            flags = ys < ys.quantile(bone_y_threshold)
            x0, z0 = xs[flags].quantile(0.5), zs[flags].quantile(0.5)
            x_margin = (xs[flags].quantile(0.95) - xs[flags].quantile(0.05)) * 0.2
            z_margin = (zs[flags].quantile(0.95) - zs[flags].quantile(0.05)) * 0.2
            quadrants = [(xs - x0 > x_margin) & (zs - z0 > z_margin), (xs - x0 > x_margin) & (zs < z0),
                         (xs - x0 < -x_margin) & (zs < z0), (xs - x0 < -x_margin) & (zs - z0 > z_margin)]
        leg_blend = torch.linspace(0., 1., n_leg_bones + 1, device=seq_shape.device)[None, None, :, None]

        def find_leg(quadrant, body_bone_idx):
            # lowest-y vertex of the quadrant per (b,f): masked arg-min (first minimum, like argmin over the subset)
            foot = _take_vertex(seq_shape, torch.where(quadrant, ys, torch.full_like(ys, float("inf"))).argmin(dim=2))
            if body_bone_idx is None:
                # the reference fixes the index on the first (b,f) it visits and reuses it (skinning.py:190-192)
                if not bool(quadrant[0, 0].any()):
                    raise RuntimeError("estimate_bones: no vertex in a leg quadrant")
                body_bone_idx = int(torch.argmin((bones_pred[0, 0, :, 1, 2] - foot[0, 0, 2]).abs()))
            body_joint = bones_pred[:, :, body_bone_idx, 1]
            return foot[:, :, None, :] * (1 - leg_blend) + body_joint[:, :, None, :] * leg_blend, body_bone_idx

        if legs_to_body_joint_indices is None:
            legs_to_body_joint_indices = [None, None, None, None]
        start_bone_idx = n_body_bones
        all_leg_bones = []
        leg_auxs = [] if compute_kinematic_chain else aux["legs"]
        for i, quadrant in enumerate(quadrants):
            if compute_kinematic_chain:
                body_bone_idx = legs_to_body_joint_indices[i]
                if i == 2:
                    body_bone_idx = legs_to_body_joint_indices[1]
                elif i == 3:
                    body_bone_idx = legs_to_body_joint_indices[0]
                leg_joints, body_bone_idx = find_leg(quadrant, body_bone_idx)
                legs_to_body_joint_indices[i] = body_bone_idx
                leg_b2j, leg_chain, leg_ids = build_kinematic_chain(n_leg_bones, start_bone_idx)
                kinematic_chain = update_body_kinematic_chain(kinematic_chain, leg_chain, body_bone_idx, leg_ids, attach_legs_to_body)
                leg_auxs.append({"body_bone_idx": body_bone_idx, "leg_bones_to_joints": leg_b2j})
                start_bone_idx += n_leg_bones
            else:
                body_bone_idx = leg_auxs[i]["body_bone_idx"]
                leg_joints, _ = find_leg(quadrant, body_bone_idx)
                leg_b2j = leg_auxs[i]["leg_bones_to_joints"]
            all_leg_bones.append(_joints_to_bones(leg_joints, leg_b2j))
        all_bones = torch.cat([bones_pred] + all_leg_bones, dim=2)
    else:
        all_bones = bones_pred

    if compute_kinematic_chain:
        aux["kinematic_chain"] = kinematic_chain
        if n_leg_bones > 0:
            aux["legs"] = leg_auxs
        return all_bones.detach(), kinematic_chain, aux
    return all_bones.detach()


# ---------------------------------------------------------------------------------------------------------------
# rotations (reference skinning.py:289-340; thin torch helpers kept for callers such as the articulation code)
# ---------------------------------------------------------------------------------------------------------------
def _axis_angle_rotation(axis, angle):
    c, s = torch.cos(angle), torch.sin(angle)
    o, z = torch.ones_like(angle), torch.zeros_like(angle)
    flat = {"X": (o, z, z, z, c, -s, z, s, c), "Y": (c, z, s, z, o, z, -s, z, c), "Z": (c, -s, z, s, c, z, z, z, o)}
    if axis not in flat:
        raise ValueError("letter must be either X, Y or Z.")
    return torch.stack(flat[axis], -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(euler_angles, convention):
    if euler_angles.dim() == 0 or euler_angles.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    if len(convention) != 3:
        raise ValueError("Convention must have 3 letters.")
    if convention[1] in (convention[0], convention[2]):
        raise ValueError(f"Invalid convention {convention}.")
    for letter in convention:
        if letter not in ("X", "Y", "Z"):
            raise ValueError(f"Invalid letter {letter} in convention string.")
    m = [_axis_angle_rotation(c, e) for c, e in zip(convention, torch.unbind(euler_angles, -1))]
    return torch.matmul(torch.matmul(m[0], m[1]), m[2])


# ---------------------------------------------------------------------------------------------------------------
# skinning
# ---------------------------------------------------------------------------------------------------------------
class _LazyWeights:
    """aux['vertices_to_bones'] is only read by visualisation code; compute it on first access."""

    def __init__(self, fn):
        self._fn, self._val = fn, None

    def get(self):
        if self._val is None:
            self._val = self._fn()
        return self._val


class _SkinningAux(dict):
    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        return v.get() if isinstance(v, _LazyWeights) else v

    def get(self, key, default=None):
        return self[key] if key in self else default


_chain_cache = {}


_chain_last = [None, None, None, None]   # (tree object, K, device, tables): the same list object comes back every step


def _chain_tables(kinematic_tree, K, device):
    if _chain_last[0] is kinematic_tree and _chain_last[1] == K and _chain_last[2] == device:
        return _chain_last[3]
    tables = _chain_tables_by_value(kinematic_tree, K, device)
    _chain_last[:] = [kinematic_tree, K, device, tables]
    return tables


def _chain_tables_by_value(kinematic_tree, K, device):
    key = (tuple((int(b), tuple(int(c) for c in ch)) for b, ch in kinematic_tree), K, str(device))
    if key not in _chain_cache:
        if len(_chain_cache) > 64:
            _chain_cache.clear()
        _chain_cache[key] = ops.chain_tables(kinematic_tree, K, device)
    return _chain_cache[key]


def skinning(v_pos, bones_pred, kinematic_tree, deform_params, output_posed_bones=False, temperature=1):
    """Reference skinning.py:369-439.  v_pos [B|1,F|1,V,3], bones_pred [B|1,F|1,K,2,3], deform_params [B,F,K,3] (rad)
    -> (posed verts [B,F,V,3], aux{bones_pred, vertices_to_bones [K,B',F',V], posed_bones [B,F,K,2,3]})."""
    B, Fr, K = deform_params.shape[:3]
    V = v_pos.shape[2]
    BF = B * Fr

    def flat(x, inner):
        # broadcast batch/frame dims only when they disagree with each other (kernels broadcast a leading 1 themselves)
        if x.shape[0] * x.shape[1] == 1:
            return x.reshape(1, *inner)
        if x.shape[0] != B or x.shape[1] != Fr:
            x = x.expand(B, Fr, *inner)
        return x.reshape(BF, *inner)

    chain_ptr, chain_ids = _chain_tables(kinematic_tree, K, deform_params.device)
    bones_in = bones_pred.detach() if not bones_pred.requires_grad else bones_pred

    def plain(x):       # batch/frame dims both 1, or both full: no broadcast to materialise
        return x.shape[0] * x.shape[1] == 1 or (x.shape[0] == B and x.shape[1] == Fr)

    if plain(v_pos) and plain(bones_in):
        v_flat = b_flat = None
        out4, posed4 = ops.lbs_bf(v_pos, bones_in, deform_params, chain_ptr, chain_ids, temperature)
    else:
        v_flat = flat(v_pos, (V, 3))
        b_flat = flat(bones_in, (K, 2, 3))
        out, posed, _ = ops.lbs(v_flat, b_flat, deform_params.reshape(BF, K, 3), chain_ptr, chain_ids, temperature, want_weights=False)
        out4, posed4 = out.reshape(B, Fr, V, 3), posed.reshape(B, Fr, K, 2, 3)

    def weights():
        nonlocal v_flat, b_flat
        if v_flat is None:
            v_flat, b_flat = flat(v_pos, (V, 3)), flat(bones_in, (K, 2, 3))
        with torch.no_grad():
            _, _, w = ops.lbs(v_flat.detach(), b_flat.detach(), deform_params.detach().reshape(BF, K, 3), chain_ptr, chain_ids,
                              temperature, want_weights=True)
        Bw = w.shape[1]
        if Bw == 1:
            return w.reshape(K, 1, 1, V)
        return w.reshape(K, B, Fr, V)

    aux = _SkinningAux()
    aux["bones_pred"] = bones_pred
    aux["vertices_to_bones"] = _LazyWeights(weights)
    if output_posed_bones:
        aux["posed_bones"] = posed4
    return out4, aux
