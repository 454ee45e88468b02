"""Seeded synthetic scene: the inputs of the hot path (not part of it).

The reference needs a dataset (SMPL-H body model, motion, cameras, env-maps) and trained
checkpoints, none of which exist offline.  This module fabricates a
scene with the same *shapes, conventions and value ranges* the reference dataset hands to
`renderer.render(batch)` (lib/datasets/pose_dataset.py:45-113, base_dataset.py:308-397):

* an SMPL-H-shaped articulated body: N=6890 vertices on a union of capsules around a
  52-joint skeleton (y-up "pose space", as SMPL), unit normals, sparse skinning weights;
* per-frame rigid bone transforms `A` / `big_A` (rest -> posed / rest -> "big pose", the
  legs-apart canonical pose of base_dataset.py:222-241), global `R`,`Th`;
* rays through the posed body's AABB with near/far (data_utils.py:827-875,925-938);
* HDR env-map probes (16x32x3);
* a network state-dict with the reference's key names and init schemes
  (net_utils.py:1242-1352, base_network.py:14-171, relight_network.py:45-72).

Everything is numpy float64 -> float32 and deterministic for a given seed.  It is used by
tests/, bench.py and the oracle alike; it never imports `oracle/`.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

N_VERTS = 6890
N_BONES = 52
ENV_H, ENV_W, ENV_R = 16, 32, 10.0

# ----------------------------------------------------------------------------- skeleton
# SMPL-H joint order (22 body joints + 15 per hand), y-up, x = subject's left, z = forward.
_BODY_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19]
_BODY_JOINTS = np.array([
    [0.00, 0.92, 0.00],   # 0 pelvis
    [0.09, 0.84, 0.00],   # 1 l_hip
    [-0.09, 0.84, 0.00],  # 2 r_hip
    [0.00, 1.04, -0.01],  # 3 spine1
    [0.10, 0.47, 0.00],   # 4 l_knee
    [-0.10, 0.47, 0.00],  # 5 r_knee
    [0.00, 1.17, 0.00],   # 6 spine2
    [0.10, 0.08, -0.02],  # 7 l_ankle
    [-0.10, 0.08, -0.02],  # 8 r_ankle
    [0.00, 1.28, 0.00],   # 9 spine3
    [0.11, 0.03, 0.10],   # 10 l_foot
    [-0.11, 0.03, 0.10],  # 11 r_foot
    [0.00, 1.46, -0.01],  # 12 neck
    [0.07, 1.39, 0.00],   # 13 l_collar
    [-0.07, 1.39, 0.00],  # 14 r_collar
    [0.00, 1.58, 0.01],   # 15 head
    [0.18, 1.40, 0.00],   # 16 l_shoulder
    [-0.18, 1.40, 0.00],  # 17 r_shoulder
    [0.44, 1.40, 0.00],   # 18 l_elbow
    [-0.44, 1.40, 0.00],  # 19 r_elbow
    [0.69, 1.40, 0.00],   # 20 l_wrist
    [-0.69, 1.40, 0.00],  # 21 r_wrist
], dtype=np.float64)


def _hand(sign: float, wrist: np.ndarray):
    """15 finger joints (5 fingers x 3 phalanges) beyond a wrist; returns joints, local parents."""
    joints, parents = [], []
    spread = [-0.035, -0.017, 0.0, 0.017, 0.034]          # along z
    length = [0.028, 0.03, 0.032, 0.03, 0.026]
    for f in range(5):
        base = wrist + np.array([sign * 0.085, 0.0, spread[f]])
        for k in range(3):
            joints.append(base + np.array([sign * length[f] * k, 0.0, 0.0]))
            parents.append(-1 if k == 0 else len(joints) - 2)
    return np.array(joints), parents


def make_skeleton():
    joints = [_BODY_JOINTS]
    parents = list(_BODY_PARENTS)
    for sign, wrist in ((1.0, 20), (-1.0, 21)):
        hj, hp = _hand(sign, _BODY_JOINTS[wrist])
        off = sum(len(j) for j in joints)
        joints.append(hj)
        parents += [wrist if p < 0 else off + p for p in hp]
    joints = np.concatenate(joints, 0)
    assert joints.shape == (N_BONES, 3) and len(parents) == N_BONES
    return joints, np.array(parents, dtype=np.int64)


def _bone_segments(joints: np.ndarray, parents: np.ndarray):
    """Capsules owned by joint j: one per child (j -> child); leaves get a short stub."""
    children = [[] for _ in range(len(joints))]
    for c, p in enumerate(parents):
        if p >= 0:
            children[p].append(c)
    radius_body = {0: 0.125, 3: 0.125, 6: 0.13, 9: 0.12, 12: 0.05, 15: 0.095,
                   1: 0.078, 2: 0.078, 4: 0.052, 5: 0.052, 7: 0.04, 8: 0.04, 10: 0.035, 11: 0.035,
                   13: 0.06, 14: 0.06, 16: 0.047, 17: 0.047, 18: 0.037, 19: 0.037, 20: 0.03, 21: 0.03}
    segs = []  # (owner, a, b, r)
    for j in range(len(joints)):
        r = radius_body.get(j, 0.0085)
        if j == 15:  # head: blob above the head joint
            segs.append((j, joints[j] + [0, 0.02, 0.0], joints[j] + [0, 0.10, 0.01], r))
            continue
        if j in (10, 11):  # toes stub forward
            segs.append((j, joints[j], joints[j] + [0, 0.0, 0.07], r))
            continue
        if j in (20, 21):  # palm
            s = 1.0 if j == 20 else -1.0
            segs.append((j, joints[j], joints[j] + [s * 0.07, 0, 0], r))
            continue
        if not children[j]:
            par = joints[parents[j]]
            d = joints[j] - par
            d = d / (np.linalg.norm(d) + 1e-9)
            segs.append((j, joints[j], joints[j] + d * 0.02, r))
            continue
        for c in children[j]:
            if j in (0, 9) and c in (1, 2, 13, 14):   # pelvis->hips, spine3->collars: keep torso capsule only
                continue
            if j >= 20 and c >= 22 and j < 22:         # wrist -> finger bases handled by the palm capsule
                continue
            segs.append((j, joints[j], joints[c], r))
    return segs


def _seg_dist(x: np.ndarray, a: np.ndarray, b: np.ndarray):
    ab = b - a
    t = np.clip(((x - a) @ ab) / (ab @ ab + 1e-12), 0.0, 1.0)
    cp = a + t[:, None] * ab
    return np.linalg.norm(x - cp, axis=-1), cp


def capsule_union_sdf(x: np.ndarray, segs) -> np.ndarray:
    d = np.full(len(x), 1e9)
    for _, a, b, r in segs:
        d = np.minimum(d, _seg_dist(x, np.asarray(a, float), np.asarray(b, float))[0] - r)
    return d


def _rodrigues(r: np.ndarray) -> np.ndarray:
    """axis-angle (...,3) -> rotation matrices (...,3,3)."""
    r = np.asarray(r, dtype=np.float64)
    th = np.linalg.norm(r, axis=-1, keepdims=True)
    k = r / np.maximum(th, 1e-12)
    K = np.zeros(r.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    th = th[..., None]
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def rigid_transform(poses: np.ndarray, joints: np.ndarray, parents: np.ndarray):
    """Forward kinematics. Returns posed joints (J,3) and A (J,4,4): rest -> posed per bone.
    Same contract as the reference's get_rigid_transform (net_utils.py:1163-1172 via smplx)."""
    R = _rodrigues(poses.reshape(-1, 3))
    J = len(joints)
    G = np.zeros((J, 4, 4))
    for j in range(J):
        L = np.eye(4)
        L[:3, :3] = R[j]
        L[:3, 3] = joints[j] - (joints[parents[j]] if parents[j] >= 0 else 0.0)
        G[j] = L if parents[j] < 0 else G[parents[j]] @ L
    posed = G[:, :3, 3].copy()
    A = G.copy()
    A[:, :3, 3] = G[:, :3, 3] - np.einsum('jab,jb->ja', G[:, :3, :3], joints)
    return posed, A


def _lbs(verts, norms, weights, A):
    Av = np.einsum('nj,jab->nab', weights, A)
    v = np.einsum('nab,nb->na', Av[:, :3, :3], verts) + Av[:, :3, 3]
    n = np.einsum('nab,nb->na', Av[:, :3, :3], norms)
    n = n / (np.linalg.norm(n, axis=-1, keepdims=True) + 1e-12)
    return v, n


@dataclass
class Body:
    joints: np.ndarray      # (52,3) rest joints
    parents: np.ndarray     # (52,)
    rverts: np.ndarray      # (N,3) rest-pose vertices
    rnorm: np.ndarray       # (N,3)
    weights: np.ndarray     # (N,52) float32 rows sum to 1
    big_A: np.ndarray       # (52,4,4)
    tverts: np.ndarray      # (N,3) big-pose vertices ("tverts" in the reference batch)
    tnorm: np.ndarray
    big_segs: list = field(default_factory=list)   # capsules in big pose (for the SDF fit)


_BODY_CACHE: Dict[int, Body] = {}


SMPL_BONES = 24     # the reference's SMPL subjects (configs/my_zju_mocap, configs/synthetic_human): cfg.n_bones 24, cond_dim 72


def _reduce_to_smpl24(body: Body) -> Body:
    """The same body on SMPL's 24-joint skeleton: the 15 finger joints of each hand collapse into one hand joint that
    follows the wrist.  Rest / big-pose vertices and normals are unchanged (fingers are not posed in the big pose)."""
    hands = [(22, 37, 20), (37, 52, 21)]                       # (first, end, wrist) of the finger blocks
    joints = np.concatenate([body.joints[:22]] + [body.joints[a + 6:a + 7] for a, _, _ in hands])       # middle-finger base
    parents = np.concatenate([body.parents[:22], [20, 21]]).astype(np.int64)
    w = np.concatenate([body.weights[:, :22]] + [body.weights[:, a:b].sum(1, keepdims=True) for a, b, _ in hands], axis=1)
    big_A = np.concatenate([body.big_A[:22]] + [body.big_A[wr:wr + 1] for _, _, wr in hands])
    return Body(joints, parents, body.rverts, body.rnorm, w.astype(np.float32), big_A, body.tverts, body.tnorm, body.big_segs)


def make_body(seed: int = 0, n_bones: int = N_BONES) -> Body:
    if n_bones == SMPL_BONES:
        if (seed, n_bones) not in _BODY_CACHE:
            _BODY_CACHE[(seed, n_bones)] = _reduce_to_smpl24(make_body(seed))
        return _BODY_CACHE[(seed, n_bones)]
    assert n_bones == N_BONES, 'n_bones must be 52 (SMPL-H) or 24 (SMPL)'
    if seed in _BODY_CACHE:
        return _BODY_CACHE[seed]
    rng = np.random.default_rng(1000 + seed)
    joints, parents = make_skeleton()
    segs = _bone_segments(joints, parents)
    # --- sample points on capsule surfaces, area-weighted, reject those inside another capsule
    areas = np.array([2 * math.pi * r * np.linalg.norm(np.asarray(b) - a) + 4 * math.pi * r * r for _, a, b, r in segs])
    # fingers are tiny: boost their share a little so every bone owns vertices
    boost = np.array([3.0 if o >= 22 else 1.0 for o, *_ in segs])
    share = areas * boost / (areas * boost).sum()
    pts, nrm = [], []
    total = int(N_VERTS * 3.0)
    for (o, a, b, r), s in zip(segs, share):
        m = max(int(total * s), 24)
        a, b = np.asarray(a, float), np.asarray(b, float)
        L = np.linalg.norm(b - a)
        axis = (b - a) / (L + 1e-12)
        # orthonormal frame
        tmp = np.array([1.0, 0, 0]) if abs(axis[0]) < 0.9 else np.array([0, 1.0, 0])
        u = np.cross(axis, tmp); u /= np.linalg.norm(u)
        v = np.cross(axis, u)
        # choose cylinder vs caps by area
        p_cyl = (2 * math.pi * r * L) / (2 * math.pi * r * L + 4 * math.pi * r * r)
        is_cyl = rng.random(m) < p_cyl
        ang = rng.random(m) * 2 * math.pi
        h = rng.random(m) * L
        zc = rng.random(m) * 2 - 1                      # sphere cap sampling
        rad = np.sqrt(np.maximum(1 - zc * zc, 0))
        n_cyl = np.cos(ang)[:, None] * u + np.sin(ang)[:, None] * v
        p_cyl_pts = a + h[:, None] * axis + r * n_cyl
        n_sph = rad[:, None] * (np.cos(ang)[:, None] * u + np.sin(ang)[:, None] * v) + zc[:, None] * axis
        centre = np.where((zc > 0)[:, None], b, a)
        p_sph = centre + r * n_sph
        pts.append(np.where(is_cyl[:, None], p_cyl_pts, p_sph))
        nrm.append(np.where(is_cyl[:, None], n_cyl, n_sph))
    pts, nrm = np.concatenate(pts), np.concatenate(nrm)
    keep = capsule_union_sdf(pts, segs) > -1e-6
    pts, nrm = pts[keep], nrm[keep]
    assert len(pts) >= N_VERTS, len(pts)
    sel = rng.permutation(len(pts))[:N_VERTS]
    rverts, rnorm = pts[sel], nrm[sel]
    # --- skinning weights: gaussian of the distance to the bones' capsules, top-4, normalised
    dist = np.full((N_VERTS, N_BONES), 1e9)
    for o, a, b, r in segs:
        d = np.maximum(_seg_dist(rverts, np.asarray(a, float), np.asarray(b, float))[0] - r, 0.0)
        dist[:, o] = np.minimum(dist[:, o], d)
    w = np.exp(-dist ** 2 / (2 * 0.035 ** 2))
    kth = np.sort(w, axis=1)[:, -4][:, None]
    w = np.where(w >= kth, w, 0.0)
    w = w / w.sum(1, keepdims=True)
    # --- big pose (legs +-30 deg about z), base_dataset.py:222-229
    big_poses = np.zeros((N_BONES, 3))
    big_poses.reshape(-1)[5] = np.deg2rad(30)
    big_poses.reshape(-1)[8] = np.deg2rad(-30)
    big_joints, big_A = rigid_transform(big_poses, joints, parents)
    tverts, tnorm = _lbs(rverts, rnorm, w, big_A)
    big_segs = []
    for o, a, b, r in segs:
        a4 = big_A[o, :3, :3] @ np.asarray(a, float) + big_A[o, :3, 3]
        b4 = big_A[o, :3, :3] @ np.asarray(b, float) + big_A[o, :3, 3]
        big_segs.append((o, a4, b4, r))
    body = Body(joints, parents, rverts, rnorm, w.astype(np.float32), big_A, tverts, tnorm, big_segs)
    _BODY_CACHE[seed] = body
    return body


# ----------------------------------------------------------------------------- motion
def make_motion(n_frames: int, seed: int = 1, n_bones: int = N_BONES):
    """Smooth random-walk SMPL-H motion: poses (T,156), Rh (T,3), Th (T,3) (motion.npz layout).
    n_bones = 24: the same body motion on SMPL's skeleton, poses (T,72) with the two hand joints at rest."""
    if n_bones == SMPL_BONES:
        poses, Rh, Th = make_motion(n_frames, seed)
        p = poses.reshape(n_frames, N_BONES, 3)[:, :SMPL_BONES].copy()
        p[:, 22:] = 0
        return p.reshape(n_frames, -1), Rh, Th
    rng = np.random.default_rng(2000 + seed)
    base = rng.normal(0, 0.2, (N_BONES, 3))
    base[0] = 0                     # root rotation lives in Rh
    base[22:] *= 0.5                # fingers
    base[[1, 2]] *= 0.6
    step = rng.normal(0, 0.03, (n_frames, N_BONES, 3)).cumsum(0)
    step -= step[:1]
    poses = (base[None] + step)
    poses[:, 0] = 0
    # pose space is y-up; the world is z-up: Rh ~ rotation by +90deg about x (+ noise, slow yaw drift)
    Rh = np.tile(np.array([math.pi / 2, 0.0, 0.0]), (n_frames, 1)) + rng.normal(0, 0.05, (1, 3)) \
        + rng.normal(0, 0.004, (n_frames, 3)).cumsum(0)
    Th = np.tile(np.array([0.0, 0.0, 0.0]), (n_frames, 1)) + rng.normal(0, 0.002, (n_frames, 3)).cumsum(0)
    return poses.reshape(n_frames, -1), Rh, Th


# ----------------------------------------------------------------------------- rays
def make_camera(H: int, W: int, target: np.ndarray, dist: float = 3.0, azim_deg: float = 20.0,
                elev_deg: float = 5.0, ixt_ratio: float = 0.8):
    """K as pose_dataset.py:58-64 (uses H for both focal and both principal coordinates);
    world z-up camera looking at `target` from `dist` metres.  Returns K, R (w2c), T (3,1)."""
    K = np.zeros((3, 3))
    K[0, 0] = K[1, 1] = H * ixt_ratio
    K[0, 2] = K[1, 2] = H / 2
    K[2, 2] = 1
    az, el = np.deg2rad(azim_deg), np.deg2rad(elev_deg)
    c = target + dist * np.array([math.cos(el) * math.sin(az), -math.cos(el) * math.cos(az), math.sin(el)])
    fwd = target - c; fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, [0, 0, 1.0]); right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], 0)          # rows: camera x,y,z in world
    T = -(R @ c)[:, None]
    return K, R, T


def get_rays(H, W, K, R, T):
    """data_utils.py:827-845 restated."""
    ray_o = -(R.T @ T).ravel()
    i, j = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing='ij')
    xy1 = np.stack([j, i, np.ones_like(i)], axis=2)
    pixel_camera = xy1 @ np.linalg.inv(K).T
    pixel_world = (pixel_camera - T.ravel()) @ R
    ray_d = pixel_world - ray_o[None, None]
    ray_d = ray_d / np.linalg.norm(ray_d, axis=2, keepdims=True)
    ray_o = np.broadcast_to(ray_o, ray_d.shape)
    return ray_o, ray_d


def get_near_far(bounds, ray_o, ray_d):
    """data_utils.py:848-875 restated (rays are unit length)."""
    norm_d = np.linalg.norm(ray_d, axis=-1, keepdims=True)
    viewdir = ray_d / norm_d
    viewdir[(viewdir < 1e-5) & (viewdir > -1e-10)] = 1e-5
    viewdir[(viewdir > -1e-5) & (viewdir < 1e-10)] = -1e-5
    tmin = (bounds[:1] - ray_o[:1]) / viewdir
    tmax = (bounds[1:2] - ray_o[:1]) / viewdir
    t1, t2 = np.minimum(tmin, tmax), np.maximum(tmin, tmax)
    near, far = t1.max(-1), t2.min(-1)
    mask = near < far
    near, far = near / norm_d[..., 0], far / norm_d[..., 0]
    return near[mask], far[mask], mask


# ----------------------------------------------------------------------------- env-maps
def make_envmaps(n: int = 8, seed: int = 10) -> Dict[str, np.ndarray]:
    """HDR probes (16,32,3): low-passed exp(N(0,1)) skies + one OLAT (base_dataset.py:136-143)."""
    out = {}
    for i in range(n):
        rng = np.random.default_rng(seed + i)
        if i == n - 1 and n > 1:
            p = np.full((ENV_H * ENV_W, 3), 0.25)
            p[4 * 32 + 7] = 100.0
            out[f'olat{4 * 32 + 7:04d}'] = p.reshape(ENV_H, ENV_W, 3).astype(np.float32)
            continue
        g = rng.normal(0, 1, (ENV_H, ENV_W, 3))
        for _ in range(2):   # separable 3-tap blur, wrap in longitude, clamp in latitude
            g = (np.roll(g, 1, 1) + 2 * g + np.roll(g, -1, 1)) / 4
            gp = np.concatenate([g[:1], g, g[-1:]], 0)
            g = (gp[:-2] + 2 * gp[1:-1] + gp[2:]) / 4
        p = np.exp(1.5 * g / g.std()) * 0.5
        sun = rng.integers(0, ENV_H // 2), rng.integers(0, ENV_W)
        p[sun] += 30.0 * rng.uniform(0.5, 1.0, 3)
        out[f'sky{i:02d}'] = p.astype(np.float32)
    return out


# ----------------------------------------------------------------------------- batch
def make_batch(H: int = 512, W: int = 512, frame: int = 0, n_frames: int = 1, seed: int = 0,
               n_env: int = 1, cam_dist: float = 3.0, azim_deg: float = 20.0, n_bones: int = N_BONES) -> Dict[str, np.ndarray]:
    """The `batch` dict of SURVEY.md 8b for one frame, as float32/int numpy arrays with the
    leading B=1 dimension (what DataLoader's default_collate would produce)."""
    body = make_body(seed, n_bones)
    poses, Rh, Th = make_motion(max(n_frames, frame + 1), seed + 1, n_bones)
    pose = poses[frame].reshape(-1, 3)
    _, A = rigid_transform(pose, body.joints, body.parents)
    pverts, pnorm = _lbs(body.rverts, body.rnorm, body.weights.astype(np.float64), A)
    R = _rodrigues(Rh[frame])
    th = Th[frame] + np.array([0.0, 0.0, 0.0])
    wverts = pverts @ R.T + th
    # put the feet on z = 0 (world), like a mocap stage
    th = th - np.array([0, 0, wverts[:, 2].min()])
    wverts = pverts @ R.T + th
    wnorm = pnorm @ R.T
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)

    def bounds(x, pad=0.05):
        return np.stack([x.min(0) - pad, x.max(0) + pad]).astype(np.float32)

    wb = bounds(wverts)
    target = (wb[0] + wb[1]) / 2
    K, cR, cT = make_camera(H, W, target.astype(np.float64), cam_dist, azim_deg)
    ro, rd = get_rays(H, W, K, cR, cT)
    ro, rd = f32(ro.reshape(-1, 3)), f32(rd.reshape(-1, 3))
    near, far, mask = get_near_far(wb, ro, rd)
    ro, rd = ro[mask], rd[mask]
    train_poses, _, _ = make_motion(4, seed + 7, n_bones)      # "training motion": material condition source
    b = dict(
        ray_o=ro[None], ray_d=rd[None], near=f32(near)[None], far=f32(far)[None],
        mask_at_box=mask.reshape(1, H, W),
        R=f32(R)[None], Th=f32(th)[None, None], Rh=f32(Rh[frame])[None],
        poses=f32(pose)[None], A=f32(A)[None], big_A=f32(body.big_A)[None],
        weights=f32(body.weights)[None],
        pverts=f32(pverts)[None], pnorm=f32(pnorm)[None],
        tverts=f32(body.tverts)[None], tnorm=f32(body.tnorm)[None],
        wverts=f32(wverts)[None], wnorm=f32(wnorm)[None],
        wbounds=wb[None], pbounds=bounds(pverts)[None], tbounds=bounds(body.tverts)[None],
        train_poses=f32(train_poses)[None],         # batch.train_motion.poses (1,T,156)
        cam_K=f32(K)[None], cam_R=f32(cR)[None], cam_T=f32(cT)[None],
        H=np.int64(H), W=np.int64(W), frame_index=np.int64(frame), view_index=np.int64(0),
        latent_index=np.int64(frame),
    )
    if n_env > 0:
        b['novel_lights'] = {k: v[None] for k, v in make_envmaps(n_env, 10 + seed).items()}
    return b


# ----------------------------------------------------------------------------- weights
_FIT_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'sdf_fit_seed0.npz')


def make_state_dict(seed: int = 0, relight: bool = True, fitted: bool = True, n_bones: int = N_BONES):
    """Network state-dict with the reference's key names and init schemes.

    residual_deformation_network.mlp.linears.{0..8}: nn.Linear default init, last bias 0 (base_network.py:31-32)
    signed_distance_network.mlp.lin{0..8}: geometric init + weight_norm g/v (net_utils.py:1276-1330)
    render_network.l{0..4}: weight-normed nn.Linear default (base_network.py:145-149)
    albedo/roughness_network.linears.{0..2}: kaiming_normal_ weights (relight_network.py:46-47)
    global_env_map_: rand(32,64,1)*0.2 (relight_network.py:64);  light grid buffers (relight_utils.py:423-465)

    With `fitted=True` the SDF MLP's (v,g,bias) are replaced by the committed fit to the synthetic
    body's big-pose capsule-union SDF (tools/fit_synthetic_sdf.py), so that the learned zero-set,
    not the SMPL proxy, terminates the rays -- the situation the trained reference model is in.
    """
    import torch
    g = torch.Generator().manual_seed(seed)

    def linear(i, o):
        bound = 1 / math.sqrt(i)
        w = (torch.rand(o, i, generator=g) * 2 - 1) * bound      # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(i), 1/sqrt(i))
        b = (torch.rand(o, generator=g) * 2 - 1) * bound
        return w, b

    sd = {}
    # residual deformation: 219 -> 256 x4 -> (256+219) -> 256 x3 -> 3   (219 = 63 + cond_dim, cond_dim = 3 * n_bones = 156)
    cond = 3 * n_bones
    dims_in = [63 + cond, 256, 256, 256, 256 + 63 + cond, 256, 256, 256, 256]
    dims_out = [256] * 8 + [3]
    for l, (i, o) in enumerate(zip(dims_in, dims_out)):
        w, b = linear(i, o)
        if l == 8:
            b = torch.zeros_like(b)
        sd[f'residual_deformation_network.mlp.linears.{l}.weight'] = w
        sd[f'residual_deformation_network.mlp.linears.{l}.bias'] = b
    # SDF: 51 -> 256,256,256,205 -> (205+51) -> 256,256,256 -> 257 ; geometric init
    d_in = 51
    dims = [d_in] + [256] * 8 + [257]
    for l in range(9):
        out_dim = dims[l + 1] - dims[0] if (l + 1) == 4 else dims[l + 1]
        w = torch.empty(out_dim, dims[l]); b = torch.zeros(out_dim)
        if l == 8:
            w.normal_(math.sqrt(math.pi) / math.sqrt(dims[l]), 0.0001, generator=g); b.fill_(-0.5)
        elif l == 0:
            w.zero_(); w[:, :3].normal_(0.0, math.sqrt(2) / math.sqrt(out_dim), generator=g)
        elif l == 4:
            w.normal_(0.0, math.sqrt(2) / math.sqrt(out_dim), generator=g); w[:, -(dims[0] - 3):] = 0
        else:
            w.normal_(0.0, math.sqrt(2) / math.sqrt(out_dim), generator=g)
        sd[f'signed_distance_network.mlp.lin{l}.weight_v'] = w
        sd[f'signed_distance_network.mlp.lin{l}.weight_g'] = w.norm(dim=1, keepdim=True)
        sd[f'signed_distance_network.mlp.lin{l}.bias'] = b
    sd['signed_distance_network._beta'] = torch.tensor(0.1)
    if fitted:
        if not os.path.exists(_FIT_PATH):
            raise FileNotFoundError(f'{_FIT_PATH} missing: run tools/fit_synthetic_sdf.py')
        fit = np.load(_FIT_PATH)
        for k in fit.files:
            sd[k] = torch.from_numpy(fit[k].astype(np.float32))
    # render network (AniSDF colour): 286 -> 256 x3 -> (256+156) -> 256 -> 3
    for l, (i, o) in enumerate([(286, 256), (256, 256), (256, 256), (256 + cond, 256), (256, 3)]):
        w, b = linear(i, o)
        sd[f'render_network.l{l}.weight_v'] = w
        sd[f'render_network.l{l}.weight_g'] = w.norm(dim=1, keepdim=True)
        sd[f'render_network.l{l}.bias'] = b
    if relight:
        for name, out in (('albedo_network', 3), ('roughness_network', 1)):
            for l, (i, o) in enumerate([(256, 128), (128, 128), (128, out)]):
                _, b = linear(i, o)
                w = torch.randn(o, i, generator=g) * math.sqrt(2.0 / i)     # kaiming_normal_
                sd[f'{name}.linears.{l}.weight'] = w
                sd[f'{name}.linears.{l}.bias'] = b
        sd['global_env_map_'] = torch.rand(ENV_H * 2, ENV_W * 2, 1, generator=g) * 0.2
        xyz, area = gen_light_xyz(ENV_H, ENV_W, ENV_R)
        sd['light_xyz_'] = torch.from_numpy(xyz)
        sd['light_area'] = torch.from_numpy(area)
        sd['light_sharp'] = torch.from_numpy((1 / np.sqrt(area / np.float32(math.pi))).astype(np.float32))
    return sd


def gen_light_xyz(h: int = ENV_H, w: int = ENV_W, r: float = ENV_R):
    """Lat-long light grid (relight_utils.py:423-465), float32 like the reference buffers."""
    import torch
    lat_half = torch.pi / h / 2
    lng_half = 2 * torch.pi / w / 2
    lats = torch.linspace(torch.pi / 2 - lat_half, -torch.pi / 2 + lat_half, h)
    lngs = torch.linspace(torch.pi - lng_half, -torch.pi + lng_half, w)
    lngs, lats = torch.meshgrid(lngs, lats, indexing='xy')
    x = r * torch.cos(lats) * torch.cos(lngs)
    y = r * torch.cos(lats) * torch.sin(lngs)
    z = r * torch.sin(lats)
    xyz = torch.stack([x, y, z], -1).reshape(h, w, 3)
    sin_colat = torch.sin(torch.pi / 2 - lats)
    areas = 4 * torch.pi * sin_colat / torch.sum(sin_colat)
    return xyz.numpy().astype(np.float32), areas.numpy().astype(np.float32)


class SyntheticNet:
    """Minimal stand-in for the reference `nn.Module` the renderer receives: exposes
    `state_dict()`, `training`, `dist_th`.  (`torch.nn.Module` instances work the same way.)"""

    def __init__(self, sd, relight=True):
        self._sd = sd
        self.training = False
        self.relight = relight
        self.dist_th = 0.125 if relight else 0.1

    def state_dict(self):
        return self._sd

    def eval(self):
        return self
