"""TEST INFRASTRUCTURE -- CPU/GPU *restatement* of the reference's inference hot path.

This file is the oracle of SURVEY.md 8c(ii): a from-scratch pure-PyTorch restatement of
zju3dv/RelightableAvatar's per-pixel rendering path, written from the algorithm as stated in
SURVEY.md Appendix A and pinned against the reference's own code executed under the import-shim
harness (`oracle/ref_harness.py` -> fixtures in tests/golden/, see tests/test_oracle_vs_reference.py).
Only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference` legs may
import it.  The product (`relightableavatar_b200/`) never does.

All tensors carry the reference's implicit batch B=1 squeezed away: points are (P,3).
Works in float32 (reference precision) or float64 (ceiling for PSNR comparisons).

Reference lines followed (relative to /root/reference):
  positional encoding ........ lib/networks/embedder.py:12-37
  MLP / SphereSDF ............ lib/utils/net_utils.py:1242-1273, 1276-1352
  sdf->occ ................... lib/utils/net_utils.py:852-893
  volume weights ............. lib/utils/net_utils.py:970-999
  AABB slab test ............. lib/utils/net_utils.py:1683-1719
  LBS ........................ lib/utils/blend_utils.py:125-165, 212-329
  geodesic 3-NN .............. lib/utils/sample_utils.py:103-162 (+ pytorch3d.ops.knn_points, external)
  HDQ query / forward ........ lib/networks/deform/base_network.py:238-336, 365-515
  relight heads .............. lib/networks/relight/relight_network.py:45-120
  sphere tracing / shadows ... lib/networks/renderer/sphere_tracing_renderer.py:20-216, 265-376
  render_human / render ...... lib/networks/renderer/sphere_tracing_renderer.py:551-784, 948-1115
  novel light ................ lib/networks/renderer/novel_light_sphere_tracing.py:21-66, 101-221
  volume renderer ............ lib/networks/renderer/base_renderer.py:15-129
  env-map / BRDF / sRGB ...... lib/utils/relight_utils.py:106-127, 179-192, 468-633
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
@dataclass
class Cfg:
    """Effective config values of xuzhen_12v_geo(_fix_mat) (SURVEY.md 8 notation paragraph)."""
    relight: bool = True
    dist_th: float = 0.125            # net.dist_th: 0.125 relight / 0.1 AniSDF
    blend_radius: float = 0.075
    resd_limit: float = 0.05
    xyz_res: int = 10
    sdf_res: int = 8
    view_res: int = 4
    tonemapping: bool = True          # cfg.tonemapping_rendering (config.py:417; False for .exr/.hdr output): main pass only
    n_samples: int = 3
    surf_sample_range: float = 0.005
    st_iter: int = 16
    st_tan_i: float = 1000.0
    st_relax: float = 0.0
    st_offset: float = 0.02
    st_eps: float = 1e-8
    st_skip: int = 1
    lv_iter: int = 4
    lv_offset: float = 0.01
    lv_relax: float = 0.0
    lv_near: float = 0.02
    lv_dist_th: float = 0.125
    env_r: float = 10.0
    bbox_margin: float = 0.25
    render_chunk: int = 65536
    fresnel_f0: float = 0.02
    albedo_slope: float = 1.0
    albedo_bias: float = 0.0
    rough_slope: float = 0.9
    rough_bias: float = 0.09
    albedo_multiplier: float = 1.0
    vis_specular_map: bool = False      # main pass also returns spec_map (sphere_tracing_renderer.py:739-748)
    no_visibility: bool = False         # lvis = 1                      sphere_tracing_renderer.py:296-298
    local_visibility: bool = False      # lvis = (n.l > 0)              :299-301
    lambert_only: bool = False          # relight_utils.py:563-568
    glossy_only: bool = False
    shading_albedo: float = 0.8
    fix_material: int = 0
    always_fix_material: bool = True  # base_network.py:501-503: cond = train_motion.poses[:, fix_material] if fix_material >= 0 or always
    knn_chunk: int = 16384
    # ground-plane shading (row f2): cfg.env_lvis (config.py:135-141), cfg.ground_* (:45, :104-107, :353)
    gl_iter: int = 16
    gl_offset: float = 0.01
    gl_relax: float = 0.0
    gl_near: float = 0.02
    gl_dist_th: float = 0.005
    ground_normal: tuple = (0.0, 0.0, 1.0)
    ground_origin: tuple = (0.0, 0.0, 0.0)
    ground_albedo: tuple = (0.05, 0.05, 0.05)
    ground_attach_envmap: bool = True
    ground_shading_multiplier: float = 1.0


def anisdf_cfg() -> Cfg:
    return Cfg(relight=False, dist_th=0.1)


# ---------------------------------------------------------------------------------------------- basics
def normalize(x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    return x / (x.norm(dim=-1, keepdim=True) + eps)                       # net_utils.py:1626-1628


def positional_encoding(x: torch.Tensor, L: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ...]; per-frequency (sin xyz, cos xyz). embedder.py:26-37"""
    freqs = (2.0 ** torch.linspace(0.0, L - 1, L, dtype=torch.float32)).to(x)
    xf = x[..., None, :] * freqs[:, None]                                  # (..., L, 3)
    enc = torch.stack([torch.sin(xf), torch.cos(xf)], dim=-2)              # (..., L, 2, 3)
    return torch.cat([x, enc.reshape(*x.shape[:-1], L * 6)], dim=-1)


def softplus100(x):
    return F.softplus(x, beta=100)                                         # threshold=20 default


class Weights:
    """State-dict view with weight-norm folded (w = g * v / ||v||_row), cast to dtype/device."""

    def __init__(self, sd: Dict[str, torch.Tensor], dtype=torch.float32, device='cpu'):
        c = lambda t: t.detach().to(device=device, dtype=dtype)
        self.resd = [(c(sd[f'residual_deformation_network.mlp.linears.{l}.weight']),
                      c(sd[f'residual_deformation_network.mlp.linears.{l}.bias'])) for l in range(9)]

        def wn(prefix):
            v, g = c(sd[prefix + '.weight_v']), c(sd[prefix + '.weight_g'])
            return g * v / v.norm(dim=1, keepdim=True), c(sd[prefix + '.bias'])

        self.sdf = [wn(f'signed_distance_network.mlp.lin{l}') for l in range(9)]
        self.beta = c(sd['signed_distance_network._beta']).clamp(1e-9, 1e6)
        self.render = [wn(f'render_network.l{l}') for l in range(5)] if 'render_network.l0.weight_v' in sd else None
        self.relight = 'albedo_network.linears.0.weight' in sd
        if self.relight:
            self.albedo = [(c(sd[f'albedo_network.linears.{l}.weight']), c(sd[f'albedo_network.linears.{l}.bias'])) for l in range(3)]
            self.rough = [(c(sd[f'roughness_network.linears.{l}.weight']), c(sd[f'roughness_network.linears.{l}.bias'])) for l in range(3)]
            env = c(sd['global_env_map_'])
            self.env_main = F.softplus(env.expand(*env.shape[:2], 3))        # relight_network.py:86-89
            self.light_xyz = c(sd['light_xyz_'])
            self.light_area = c(sd['light_area'])
            self.light_sharp = c(sd['light_sharp'])


def resd_mlp(W: Weights, x: torch.Tensor) -> torch.Tensor:
    """net_utils.py:1263-1273 with skips=[4], ReLU."""
    h = x
    for i, (w, b) in enumerate(W.resd):
        if i == 4:
            h = torch.cat([h, x], dim=-1)
        h = F.linear(h, w, b)
        if i < 8:
            h = F.relu(h)
    return h


def sdf_mlp(W: Weights, x: torch.Tensor) -> torch.Tensor:
    """net_utils.py:1337-1352: skip at 4 is cat([x, in])/sqrt(2); Softplus(100) between; 257 out."""
    h = x
    for l, (w, b) in enumerate(W.sdf):
        if l == 4:
            h = torch.cat([h, x], dim=-1) / math.sqrt(2)
        h = F.linear(h, w, b)
        if l < 8:
            h = softplus100(h)
    return h


def small_mlp(layers, x: torch.Tensor) -> torch.Tensor:
    """relight MLP(256->128->128->out), Softplus(100), no skip reached (D=2 < skip 4)."""
    h = x
    for i, (w, b) in enumerate(layers):
        h = F.linear(h, w, b)
        if i < len(layers) - 1:
            h = softplus100(h)
    return h


def sdf_to_occ(sdf: torch.Tensor, beta: torch.Tensor, dists: float = 0.005) -> torch.Tensor:
    """net_utils.py:867-893 (Laplace CDF density, alpha over the fixed 5 mm interval)."""
    x = -sdf
    ind0 = (x <= 0).to(sdf.dtype)
    ind1 = 1 - ind0
    val0 = 1 / beta * (0.5 * (x * ind0 / beta).exp()) * ind0
    val1 = 1 / beta * (1 - 0.5 * (-x * ind1 / beta).exp()) * ind1
    sigma = val0 + val1
    return 1.0 - torch.exp(-F.relu(sigma) * dists)


def inverse_3x3(R: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """Adjugate / (det + 1e-8), blend_utils.py:125-165."""
    r00, r01, r02 = R[..., 0, 0], R[..., 0, 1], R[..., 0, 2]
    r10, r11, r12 = R[..., 1, 0], R[..., 1, 1], R[..., 1, 2]
    r20, r21, r22 = R[..., 2, 0], R[..., 2, 1], R[..., 2, 2]
    M = torch.empty_like(R)
    M[..., 0, 0] = r11 * r22 - r21 * r12
    M[..., 1, 0] = -r10 * r22 + r20 * r12
    M[..., 2, 0] = r10 * r21 - r20 * r11
    M[..., 0, 1] = -r01 * r22 + r21 * r02
    M[..., 1, 1] = r00 * r22 - r20 * r02
    M[..., 2, 1] = -r00 * r21 + r20 * r01
    M[..., 0, 2] = r01 * r12 - r11 * r02
    M[..., 1, 2] = -r00 * r12 + r10 * r02
    M[..., 2, 2] = r00 * r11 - r10 * r01
    D = r00 * M[..., 0, 0] + r01 * M[..., 1, 0] + r02 * M[..., 2, 0]
    return M / (D[..., None, None] + eps)


# ---------------------------------------------------------------------------------------------- frame
@dataclass
class Frame:
    """Per-frame read-only state (SURVEY.md 8b batch entries), squeezed of B."""
    R: torch.Tensor; Th: torch.Tensor; poses: torch.Tensor
    A: torch.Tensor; big_A: torch.Tensor; weights: torch.Tensor
    pverts: torch.Tensor; pnorm: torch.Tensor; tverts: torch.Tensor
    wbounds: torch.Tensor; mat_cond: torch.Tensor

    @staticmethod
    def from_batch(b: dict, cfg: Cfg, dtype=torch.float32, device='cpu') -> 'Frame':
        t = lambda a: torch.as_tensor(a).to(device=device, dtype=dtype)
        return Frame(R=t(b['R'][0]), Th=t(b['Th']).reshape(3), poses=t(b['poses']).reshape(-1),
                     A=t(b['A'][0]), big_A=t(b['big_A'][0]), weights=t(b['weights'][0]),
                     pverts=t(b['pverts'][0]), pnorm=t(b['pnorm'][0]), tverts=t(b['tverts'][0]),
                     wbounds=t(b['wbounds'][0]).clone(),
                     mat_cond=(t(b['train_poses'][0, cfg.fix_material]) if (cfg.fix_material >= 0 or cfg.always_fix_material)
                               else t(b['poses'][0])).reshape(-1))      # python indexing: -1 = last training pose


def knn3(p: torch.Tensor, verts: torch.Tensor, chunk: int):
    """Exact 3 smallest squared-L2 distances, ascending (pytorch3d.ops.knn_points K=3,
    return_sorted=True; sample_utils.py:122).  Distances as (dx^2+dy^2)+dz^2, no mm shortcut."""
    d2s, ids = [], []
    for s in range(0, p.shape[0], chunk):
        q = p[s:s + chunk]
        d = (q[:, None, 0] - verts[None, :, 0]) ** 2
        d = d + (q[:, None, 1] - verts[None, :, 1]) ** 2
        d = d + (q[:, None, 2] - verts[None, :, 2]) ** 2
        v, i = torch.topk(d, 3, dim=1, largest=False, sorted=True)
        d2s.append(v); ids.append(i)
    if not d2s:
        return p.new_zeros(0, 3), torch.zeros(0, 3, dtype=torch.long, device=p.device)
    return torch.cat(d2s), torch.cat(ids)


def world_to_bigpose(x: torch.Tensor, v: Optional[torch.Tensor], fr: Frame, cfg: Cfg, th: float):
    """base_network.py:238-336 + sample_utils.py:103-162.  Returns a dict with the per-point
    SMPL distances for ALL points and the warp results for the in-shell subset (mask `ins`)."""
    p = (x - fr.Th) @ fr.R                                                  # blend_utils.py:252-261
    d2, nn = knn3(p, fr.pverts, cfg.knn_chunk)
    dist = d2.sqrt()
    dot = ((p[:, None] - fr.pverts[nn]) * fr.pnorm[nn]).sum(-1)
    sdf_batch = dist * dot.sign()                                           # (P,3)
    ins = d2[:, 0] < th ** 2
    # geodesic filter in canonical (big-pose) space, sample_utils.py:148-160
    tv = fr.tverts[nn]
    far = ((tv - tv[:, :1]) ** 2).sum(-1) >= th ** 2                        # ~msk
    sdf_batch = torch.where(far, sdf_batch[:, :1], sdf_batch)
    d2f = torch.where(far, d2[:, :1], d2)
    nnf = torch.where(far, nn[:, :1], nn)
    out = dict(sdf_batch=sdf_batch, ins=ins)
    # shell points only
    ps, d2s, nns = p[ins], d2f[ins], nnf[ins]
    w = (-d2s / (2 * cfg.blend_radius ** 2)).exp()
    w = w / (w.sum(-1, keepdim=True) + torch.finfo(w.dtype).eps)
    bw = (w[..., None] * fr.weights[nns]).sum(-2)                           # (S,52)
    big_A = (bw[:, :, None, None] * fr.big_A[None]).sum(1)                  # blend_transform
    A = (bw[:, :, None, None] * fr.A[None]).sum(1)
    big_Rinv = inverse_3x3(big_A[:, :3, :3])
    Rinv = inverse_3x3(A[:, :3, :3])
    tp = (Rinv * (ps - A[:, :3, 3])[:, None, :]).sum(-1)                    # pose -> tpose
    bp = (big_A[:, :3, :3] * tp[:, None, :]).sum(-1) + big_A[:, :3, 3]      # tpose -> bigpose
    out.update(bpts=bp, A=A, Rinv=Rinv, big_A=big_A, big_Rinv=big_Rinv)
    if v is not None:
        pv = (v @ fr.R)[ins]                                                # world_dirs_to_pose_dirs
        tv_ = (A[:, :3, :3].mT * pv[:, None, :]).sum(-1)                    # pose_dirs_to_tpose_dirs: R^T
        bv = (big_Rinv.mT * tv_[:, None, :]).sum(-1)                        # tpose_dirs_to_pose_dirs: Rinv^T
        out['bvds'] = bv
    return out


def hdq_distance(x: torch.Tensor, fr: Frame, W: Weights, cfg: Cfg, th: float, smooth: bool = True) -> torch.Tensor:
    """inference_world_distance_field (base_network.py:365-387): (P,3) -> (P,1)."""
    o = world_to_bigpose(x, None, fr, cfg, th)
    bp = o['bpts']
    cond = fr.poses[None].expand(bp.shape[0], -1)
    resd = torch.tanh(resd_mlp(W, torch.cat([positional_encoding(bp, cfg.xyz_res), cond], -1))) * cfg.resd_limit
    net = sdf_mlp(W, positional_encoding(bp + resd, cfg.sdf_res))[:, :1]
    smpl = o['sdf_batch'].mean(-1, keepdim=True)
    smpl = torch.where(smpl < -th, smpl, smpl.abs())
    if smooth:
        d1 = smpl[o['ins']]
        r = (net.abs() / th).clip(0, 1)
        net = d1 * r + net * (1 - r)
    sdf = smpl.clone()
    sdf[o['ins']] = net
    return sdf


# ---------------------------------------------------------------------------------------------- tracing
def sphere_tracing(ray_o, ray_d, near, far, sdf_fn: Callable, iters: int, tan_i, relax: float, offset: float,
                   eps: float, skip: int, soft: bool, claybook: bool = True):
    """sphere_tracing_renderer.py:103-216 (mode 'hdq').  near/far (R,1); tan_i scalar or (R,1)."""
    ones = torch.ones_like(ray_o[:, :1])
    near, far = ones * near, ones * far
    tan = ones / tan_i
    off, rlx = ones * offset, ones * relax
    occ = ones.clone()
    d0 = ones * 1e9; cd = ones * 1e9; dt = ones * 1e9
    st, ot, t = far, far, near
    for i in range(iters):
        d1 = sdf_fn(ray_o + t * ray_d)
        if soft and claybook and i >= skip:
            dx0 = d0 + rlx * d0 + off
            dx1 = d1 + rlx * d1 + off
            dy = (dx1 ** 2) / (2 * dx0)
            dx = ((dx1 ** 2 - dy ** 2).sqrt() - off) / (1 + rlx)
            cls = dx.clip(0) / (t - dy).clip(near).clip(eps) / (tan * 2)
            msk = (cls < occ) & (dy < t) & (dx1 > 0) & (dx0 > 0) & (dx > 0) & (dy > 0) & (dy < dx0)
            ot = torch.where(msk, t - dy, ot)
            occ = torch.where(msk, cls, occ)
        if i >= skip:
            cls = d1.clip(0) / t.clip(near).clip(eps) / (tan * 2)
            msk = cls < occ
            ot = torch.where(msk, t, ot)
            occ = torch.where(msk, cls, occ)
        if not soft:
            d1u, d0u = d1.abs(), d0.abs()
            msk = d0.sign() != d1.sign()
            st = torch.where(msk, t - dt * (d1u / (d0u + d1u + eps)).clip(0, 1), st)
            off = torch.where(msk, torch.zeros_like(off), off)
            rlx = torch.where(msk, torch.zeros_like(rlx), rlx)
            msk = d1u < cd
            cd = torch.where(msk, d1u, cd)
            st = torch.where(msk, t, st)
        dt = d1 + rlx * d1 + off
        t = torch.maximum(torch.minimum(t + dt, far), near)
        d0 = d1
    return ray_o + st * ray_d, ray_o + ot * ray_d, occ, st, ot


def near_far_aabb(bounds: torch.Tensor, ray_o: torch.Tensor, ray_d: torch.Tensor, epsilon: float = 1e-8):
    """get_near_far_aabb(return_raw=True), net_utils.py:1683-1719 (works on a copy of ray_d)."""
    d = ray_d.clone()
    d[(d < epsilon) & (d > -epsilon ** 2)] = epsilon
    d[(d > -epsilon ** 2) & (d < epsilon)] = -epsilon
    tmin = (bounds[:1] - ray_o) / d
    tmax = (bounds[1:] - ray_o) / d
    t1, t2 = torch.minimum(tmin, tmax), torch.maximum(tmin, tmax)
    return t1.max(-1)[0], t2.min(-1)[0]


def light_visibility(surf, norm, acc, W: Weights, fr: Frame, cfg: Cfg, bbox: torch.Tensor, lv: Optional[dict] = None):
    """sphere_tracing_renderer.py:265-344 -> lvis, ldot of shape (512, S).  `lv` overrides the obj_lvis tracing
    parameters (iter, offset, relax, near, dist_th) -- render_ground passes cfg.env_lvis."""
    lv = lv or dict(iter=cfg.lv_iter, offset=cfg.lv_offset, relax=cfg.lv_relax, near=cfg.lv_near, dist_th=cfg.lv_dist_th)
    S = surf.shape[0]
    L = W.light_xyz.reshape(-1, 3)
    ldir = normalize(L)                                                     # (512,3)
    ldot = (ldir[:, None] * norm[None]).sum(-1)                             # (512,S)
    if cfg.no_visibility:
        return torch.ones_like(ldot), ldot
    if cfg.local_visibility:
        return (ldot > 0).to(ldot), ldot
    lfrt = (ldot > 0) & (acc[None] > 0)
    li, pi = lfrt.nonzero(as_tuple=True)
    ro, rd = surf[pi], ldir[li]
    near, far = near_far_aabb(bbox, ro, rd)
    near, far = near[:, None].clip(lv['near']), far[:, None].clip(lv['near'])
    lbox_sub = (near < far)[:, 0]
    ro, rd, near, far = ro[lbox_sub], rd[lbox_sub], near[lbox_sub], far[lbox_sub]
    tan_i = W.light_sharp.reshape(-1)[li][lbox_sub][:, None]
    sdf_fn = lambda x: hdq_distance(x, fr, W, cfg, lv['dist_th'], True)
    occ_parts = []
    for s0 in range(0, ro.shape[0], 1 << 20):                               # bounded memory; rays are independent
        sl = slice(s0, s0 + (1 << 20))
        _, _, occ, _, _ = sphere_tracing(ro[sl], rd[sl], near[sl], far[sl], sdf_fn, lv['iter'], tan_i[sl], lv['relax'], lv['offset'],
                                         cfg.st_eps, cfg.st_skip, soft=True)
        occ_parts.append(occ)
    occ = torch.cat(occ_parts) if occ_parts else ro.new_zeros(0, 1)
    lvis = torch.zeros_like(ldot)
    lbox = torch.zeros_like(lfrt)
    lbox[li[lbox_sub], pi[lbox_sub]] = True
    lvis[li[lbox_sub], pi[lbox_sub]] = occ[:, 0]
    lvis = lvis * lbox + 1.0 * (~lbox)
    lvis = lvis * lfrt
    return lvis, ldot


# ---------------------------------------------------------------------------------------------- shading
def sample_envmap(image: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """relight_utils.py:106-127: image (H,W,3), unit dirs (...,3) -> (...,3)."""
    sh = d.shape
    d = d.reshape(-1, 3)
    img = image.permute(2, 0, 1)[None]
    theta = torch.arccos(d[:, 2]) - 1e-6
    phi = torch.atan2(d[:, 1], d[:, 0])
    qy = (theta / torch.pi) * 2 - 1
    qx = -phi / torch.pi
    grid = torch.stack((qx, qy), -1)[None, None]
    rgb = F.grid_sample(img, grid, align_corners=False, padding_mode='border')
    return rgb[0, :, 0].permute(1, 0).reshape(sh)


def safe_divide(a: torch.Tensor, b: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """relight_utils.py:618-633.  NOTE: clamps `a` and `b` IN PLACE like the reference (callers rely on it)."""
    a[(a < eps) & (a >= 0)] = eps
    a[(a > -eps) & (a <= 0)] = -eps
    b[(b < eps) & (b >= 0)] = eps
    b[(b > -eps) & (b <= 0)] = -eps
    div = a / b
    div[div != div] = 0.0
    div[(div == torch.inf) | (div == -torch.inf)] = 0.0
    return div.clip(-1e10, 1e10)


def microfacet(pts2l, pts2c, normal, albedo, rough, f0: float, lambert_only: bool = False, glossy_only: bool = False):
    """relight_utils.py:484-615 with cancel_cosine=True.  pts2l (N,L,3), others (N,3)/(N,1) -> (N,L,3)."""
    pts2l = F.normalize(pts2l, p=2, dim=-1, eps=1e-7)
    pts2c = F.normalize(pts2c, p=2, dim=-1, eps=1e-7)
    normal = F.normalize(normal, p=2, dim=-1, eps=1e-7)
    l_dot_n = torch.einsum('ijk,ik->ij', pts2l, normal).clip(1e-4, 1)
    v_dot_n = torch.einsum('ij,ij->i', pts2c, normal).clip(1e-4, 1)
    lambert = albedo[:, None, :].repeat(1, pts2l.shape[1], 1) / torch.pi * l_dot_n[:, :, None]
    h = F.normalize(pts2l + pts2c[:, None, :], p=2, dim=-1, eps=1e-7)
    f = f0 + (1 - f0) * (1 - torch.einsum('ijk,ijk->ij', pts2l, h)) ** 5
    alpha = rough ** 2
    # _get_d
    cm = torch.einsum('ijk,ik->ij', h, normal)
    chi = torch.where(cm > 0, 1., 0.).to(cm)
    cm2 = torch.square(cm)
    tm2 = safe_divide(1 - cm2, cm2)                                          # clamps cm2 in place
    d = safe_divide(alpha ** 2 * chi, torch.pi * torch.square(cm2) * torch.square(alpha ** 2 + tm2))
    # _get_g
    cv = torch.einsum('ij,ij->i', normal, pts2c)
    ct = torch.einsum('ijk,ik->ij', h, pts2c)
    div = safe_divide(ct, cv[:, None])                                       # clamps cv in place (view)
    chi_g = torch.where(div > 0, 1., 0.).to(cm)
    cv2 = torch.clip(torch.square(cv), 0., 1.)
    tv2 = torch.clip(safe_divide(1 - cv2, cv2), 0., 1e10)
    g = safe_divide(chi_g * 2, 1 + torch.sqrt(1 + alpha ** 2 * tv2[:, None]))
    denom = 4 * torch.abs(torch.ones_like(l_dot_n)) * torch.abs(v_dot_n)[:, None]
    spec = safe_divide(f * g * d, denom)
    if lambert_only:
        return lambert
    if glossy_only:
        return spec[:, :, None].repeat(1, 1, 3)
    return spec[:, :, None].repeat(1, 1, 3) + lambert


def linear2srgb(x: torch.Tensor) -> torch.Tensor:
    x = x.clip(0., 1.)                                                       # relight_utils.py:179-192
    return torch.where(x <= 0.0031308, x * 12.92, 1.055 * torch.pow(x + 1e-7, 1 / 2.4) - 0.055)


def shade_pixels(ray_o, surf, norm, albedo, rough, lvis, ldot, probe, W: Weights, cfg: Cfg, chunk: int = 4096, tonemap: bool = True):
    """A.7 + A.8: env fetch, BRDF, 512-light sum.  lvis/ldot (512,S).  Returns rgb(sRGB), shade, spec (S,3).
    `tonemap`: the main pass honours cfg.tonemapping_rendering (sphere_tracing_renderer.py:731); the novel-light re-shade applies
    linear2srgb unconditionally (novel_light_sphere_tracing.py:48)."""
    L = W.light_xyz.reshape(-1, 3)
    area = W.light_area.reshape(-1)
    rgbs, shades, specs = [], [], []
    for s in range(0, surf.shape[0], chunk):
        sl = slice(s, s + chunk)
        s2l = normalize(L[None] - surf[sl, None])                           # (n,512,3)
        s2c = normalize(ray_o[sl] - surf[sl])
        light = sample_envmap(probe, s2l)                                   # (n,512,3)
        lv, ld = lvis[:, sl].T, ldot[:, sl].T                               # (n,512)
        brdf = microfacet(s2l, s2c, norm[sl], albedo[sl], rough[sl], cfg.fresnel_f0, cfg.lambert_only, cfg.glossy_only)
        shade = lv[..., None] * 1.0 * area[None, :, None] * light           # ldot := 1 (cancel_cosine)
        rgbs.append(linear2srgb((brdf * shade).sum(1)) if tonemap else (brdf * shade).sum(1))
        sb = microfacet(s2l, s2c, norm[sl], torch.zeros_like(albedo[sl]), rough[sl], cfg.fresnel_f0, cfg.lambert_only, cfg.glossy_only)
        # spec: lvis=1, ldot := 1/(|ones|+1e-8)   (sphere_tracing_renderer.py:739-749)
        specs.append((sb * (1.0 / (1.0 + 1e-8)) * area[None, :, None] * light).sum(1))
        shades.append((lv[..., None] * ld[..., None] * area[None, :, None] * light).sum(1) * cfg.shading_albedo / math.pi)
    cat = lambda l: torch.cat(l) if l else surf.new_zeros(0, 3)
    return cat(rgbs), cat(shades), cat(specs)


# ---------------------------------------------------------------------------------------------- network forward
def forward_geometry(x, v, fr: Frame, W: Weights, cfg: Cfg):
    """base_network.py:456-494: warp + resd + sdf/occ/feat + autograd normal, for in-shell points."""
    o = world_to_bigpose(x, v, fr, cfg, cfg.dist_th)
    bp = o['bpts'].detach().requires_grad_(True)
    cond = fr.poses[None].expand(bp.shape[0], -1)
    with torch.enable_grad():
        resd = torch.tanh(resd_mlp(W, torch.cat([positional_encoding(bp, cfg.xyz_res), cond], -1))) * cfg.resd_limit
        cp = bp + resd
        out = sdf_mlp(W, positional_encoding(cp, cfg.sdf_res))
        sdf, feat = out[:, :1], out[:, 1:]
        grad = torch.autograd.grad(sdf, bp, torch.ones_like(sdf))[0] if bp.shape[0] else torch.zeros_like(bp)
    occ = sdf_to_occ(sdf.detach(), W.beta)
    n = normalize(grad)
    n = (o['big_A'][:, :3, :3].mT * n[:, None, :]).sum(-1)                  # pose_dirs_to_tpose_dirs (big)
    n = (o['Rinv'].mT * n[:, None, :]).sum(-1)                              # tpose_dirs_to_pose_dirs
    n = n @ fr.R.mT                                                         # pose_dirs_to_world_dirs
    n = normalize(n)
    return dict(o, cpts=cp.detach(), bpts=bp.detach(), resd=resd.detach(), norm=n, feat=feat.detach(), occ=occ, sdf=sdf.detach())


def network_forward(x, v, fr: Frame, W: Weights, cfg: Cfg) -> torch.Tensor:
    """`net(x, v, d, batch).raw` (eval): relight 17 ch (relight_network.py:91-104) / AniSDF 16 ch
    (base_network.py:496-515); zero rows for out-of-shell points."""
    with torch.no_grad():
        pass
    g = forward_geometry(x, None if cfg.relight else v, fr, W, cfg)
    if cfg.relight:
        albedo = cfg.albedo_slope * torch.sigmoid(small_mlp(W.albedo, g['feat'])) + cfg.albedo_bias
        rough = cfg.rough_slope * torch.sigmoid(small_mlp(W.rough, g['feat'])) + cfg.rough_bias
        raw = torch.cat([g['cpts'], g['bpts'], g['resd'], albedo, rough, g['norm'], g['occ']], -1)
    else:
        cond = fr.mat_cond[None].expand(g['bpts'].shape[0], -1)
        h = torch.cat([positional_encoding(g['bvds'], cfg.view_res), g['norm'], g['feat']], -1)
        for l, (w, b) in enumerate(W.render):
            if l == 3:
                h = torch.cat([h, cond], -1)
            h = F.linear(h, w, b)
            if l < 4:
                h = F.relu(h)
        rgb = torch.sigmoid(h)
        raw = torch.cat([g['cpts'], g['bpts'], g['resd'], g['norm'], rgb, g['occ']], -1)
    full = raw.new_zeros(x.shape[0], raw.shape[-1])
    full[g['ins']] = raw
    return full


def volume_weights(alpha: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """net_utils.py:987-991: w_i = a_i * prod_{j<i}(1 - a_j + eps); alpha (P,S)."""
    ex = torch.cat([alpha.new_ones(alpha.shape[0], 1), 1. - alpha + eps], -1)
    return alpha * torch.cumprod(ex, -1)[:, :-1]


# ---------------------------------------------------------------------------------------------- renderers
def render_human(ray_o, ray_d, near, far, fr: Frame, W: Weights, cfg: Cfg, bbox, probe, want_lvis=True) -> dict:
    """sphere_tracing_renderer.py:551-784 for one pixel chunk (eval, rendering map on)."""
    P = ray_o.shape[0]
    sdf_fn = lambda x: hdq_distance(x, fr, W, cfg, cfg.dist_th, True)
    surf, edge, occ, st, ot = sphere_tracing(ray_o, ray_d, near[:, None], far[:, None], sdf_fn, cfg.st_iter, cfg.st_tan_i,
                                             cfg.st_relax, cfg.st_offset, cfg.st_eps, cfg.st_skip, soft=False)
    depth = (surf[:, 0] - ray_o[:, 0]) / ray_d[:, 0]
    acc_all = 1 - occ[:, 0]
    fg = acc_all > 0
    acc, surf, view, ro, depth = acc_all[fg], surf[fg], ray_d[fg], ray_o[fg], depth[fg]
    S = surf.shape[0]
    z = torch.linspace(0., 1., cfg.n_samples).to(surf) * (2 * cfg.surf_sample_range) - cfg.surf_sample_range
    pts = surf[:, None] + z[None, :, None] * view[:, None]
    raw = network_forward(pts.reshape(-1, 3), view[:, None].expand(-1, cfg.n_samples, -1).reshape(-1, 3), fr, W, cfg)
    raw = raw.reshape(S, cfg.n_samples, -1)
    val, a = raw[..., :-1], raw[..., -1]
    w = volume_weights(a)
    occ_s = w.sum(-1)
    val = (w[..., None] * val).sum(1) / (occ_s[:, None] + 1e-8)
    ret = dict(acc_map=acc, ray_o=ro, surf_map=surf, depth_map=depth)
    if cfg.relight:
        cpts, bpts, resd, albedo, rough, norm = val.split([3, 3, 3, 3, 1, 3], -1)
    else:
        cpts, bpts, resd, norm, rgb = val.split([3, 3, 3, 3, 3], -1)
    norm = torch.where(norm.sum(-1, keepdim=True) == 0, torch.ones_like(norm), norm)
    norm = normalize(norm)
    ret.update(cpts_map=cpts, bpts_map=bpts, resd_map=resd, norm_map=norm)
    if cfg.relight:
        albedo = albedo.clip(cfg.albedo_bias, cfg.albedo_bias + cfg.albedo_slope)
        if cfg.albedo_multiplier > 0:          # sphere_tracing_renderer.py:653-655: a non-positive multiplier means 'off'
            albedo = albedo * cfg.albedo_multiplier
        rough = rough.clip(cfg.rough_bias, cfg.rough_bias + cfg.rough_slope)
        ret.update(albedo_map=albedo, roughness_map=rough[:, 0])
        lvis, ldot = light_visibility(surf, norm, acc, W, fr, cfg, bbox)
        rgb, shade, spec = shade_pixels(ro, surf, norm, albedo, rough, lvis, ldot, probe, W, cfg, tonemap=cfg.tonemapping)
        ret.update(rgb_map=rgb, shade_map=shade)
        if cfg.vis_specular_map:
            ret.update(spec_map=spec)
        if want_lvis:
            ret.update(lvis_map=lvis.T.contiguous(), ldot_map=ldot.T.contiguous())     # (S,512)
    else:
        ret.update(rgb_map=rgb)
    # multi_scatter_zeros back to all P rays (:679-686)
    out = {}
    for k, t in ret.items():
        full = t.new_zeros(P, *t.shape[1:])
        full[fg] = t
        out[k] = full
    return out


_BLEND_KEYS = ['rgb_map', 'surf_map', 'albedo_map', 'roughness_map', 'norm_map', 'cpts_map', 'bpts_map', 'spec_map',
               'depth_map', 'lvis_map', 'ldot_map', 'shade_map']


# ---------------------------------------------------------------------------------------------- ground plane (row f2)
def get_rays_full(H: int, W: int, K: torch.Tensor, R: torch.Tensor, T: torch.Tensor):
    """net_utils.get_rays (:403-420): every pixel of the H x W image -> ray_o, ray_d (H*W, 3)."""
    ray_o = -(R.mT @ T).ravel()
    i, j = torch.meshgrid(torch.arange(H, dtype=R.dtype, device=R.device), torch.arange(W, dtype=R.dtype, device=R.device), indexing='ij')
    xy1 = torch.stack([j, i, torch.ones_like(i)], dim=2)
    pixel_camera = xy1 @ torch.inverse(K).mT
    pixel_world = (pixel_camera - T.ravel()) @ R
    ray_o = ray_o[None, None].expand(pixel_world.shape)
    ray_d = normalize(pixel_world - ray_o)
    return ray_o.reshape(-1, 3), ray_d.reshape(-1, 3)


def ground_plane_t(ray_o, ray_d, orig, norm, tangent_scale: float = 1.0, eps: float = 1e-8):
    """t of moller_trumbore (mesh_utils.py:710-738) against compute_ground_tris(orig, norm) (net_utils.py:392-396).
    The triangle is (o, o + a, o + b) with a = d x n, b = d x a for a RANDOM unit vector n, so E1 x E2 = |a|^2 d and
    t = -((ray_o - o) . N) / (ray_d . N + eps) with N = |a|^2 d: the draw only scales how much the 1e-8 matters
    (`tangent_scale` = |a|^2 in (0, 1]; relative effect on t <= 1e-8 / (|a|^2 |ray_d . d|))."""
    N = norm * tangent_scale
    invdet = 1.0 / -((ray_d * N).sum(-1) + eps)
    return ((ray_o - orig) * N).sum(-1) * invdet


def render_ground(ray_o, ray_d, acc, fr: Frame, W: Weights, cfg: Cfg, bbox, probe, image=None, tangent_scale: float = 1.0) -> dict:
    """sphere_tracing_renderer.render_ground (:463-548) for one pixel chunk; every map is (P, .)."""
    P = ray_o.shape[0]
    norm = normalize(ray_o.new_tensor(cfg.ground_normal))
    orig = ray_o.new_tensor(cfg.ground_origin)
    t = ground_plane_t(ray_o, ray_d, orig, norm, tangent_scale)[:, None]
    surf = ray_o + t * ray_d
    normP = norm[None].expand(P, 3)
    lv = dict(iter=cfg.gl_iter, offset=cfg.gl_offset, relax=cfg.gl_relax, near=cfg.gl_near, dist_th=cfg.gl_dist_th)
    lvis, ldot = light_visibility(surf, normP, acc, W, fr, cfg, bbox, lv)                     # (512, P)
    if cfg.ground_attach_envmap:
        albedo = sample_envmap(image if image is not None else probe, ray_d)
    else:
        albedo = torch.ones_like(surf) * ray_o.new_tensor(cfg.ground_albedo)
    dist = torch.where(t[:, 0] <= 0, torch.full_like(t[:, 0], 1e9), (surf - orig).norm(dim=-1))
    weight = ((dist - cfg.env_r) / cfg.env_r).clip(0, 1)[None]                                # (1, P)
    L = W.light_xyz.reshape(-1, 3)
    ldot = (normalize(L)[:, None] * normP[None]).sum(-1)
    lvis = lvis * (1 - weight) + torch.ones_like(lvis) * weight
    light = sample_envmap(probe, normalize(L))                                                # (512, 3): the same for every pixel
    area = W.light_area.reshape(-1)
    shade = lvis[..., None] * ldot[..., None] * area[:, None, None] * light[:, None]          # evaluate_shade :369-376
    rgb = ((albedo[None] / math.pi) * shade).sum(0)
    if cfg.tonemapping:                                                                       # sphere_tracing_renderer.py:523
        rgb = linear2srgb(rgb)
    shade = shade.sum(0) * cfg.shading_albedo / math.pi
    return dict(rgb_map=rgb, surf_map=surf, albedo_map=albedo, roughness_map=torch.ones_like(albedo[:, 0]), spec_map=shade / 20,
                norm_map=normP.clone(), shade_map=shade * cfg.ground_shading_multiplier, cpts_map=torch.zeros_like(surf),
                bpts_map=torch.zeros_like(surf), depth_map=t[:, 0].clip(-cfg.env_r, cfg.env_r),
                lvis_map=lvis.T.contiguous(), ldot_map=ldot.T.contiguous())


def render_ground_novel(ray_d, albedo_map, lvis_map, ldot_map, W: Weights, cfg: Cfg, probe, image=None):
    """novel_light_sphere_tracing.render_ground (:69-98): re-shade the floor from its stored (F,512) visibility maps."""
    L = W.light_xyz.reshape(-1, 3)
    light = sample_envmap(probe, normalize(L))
    albedo = sample_envmap(image if image is not None else probe, ray_d) if cfg.ground_attach_envmap else albedo_map
    area = W.light_area.reshape(-1)
    shade = lvis_map.T[..., None] * ldot_map.T[..., None] * area[:, None, None] * light[:, None]
    rgb = linear2srgb(((albedo[None] / math.pi) * shade).sum(0))                              # unconditional, :94
    shade = shade.sum(0) / math.pi
    return rgb, albedo, shade, shade / 20


def blend_output(acc, inds, grd: dict, ret: dict) -> dict:
    """blend_output_ (:433-451): image-sized maps = ground * acc_g + scatter(human) * (1 - acc_g); acc (F,), inds (P,)."""
    out = dict(ret)
    for k in _BLEND_KEYS:
        if k in grd:
            a = acc if grd[k].ndim == 1 else acc[:, None]
            if k in ret:
                sc = torch.zeros_like(grd[k])
                sc[inds] = ret[k]
                out[k] = grd[k] * a + sc * (1 - a)
            else:
                out[k] = grd[k] * a
    sc = torch.zeros_like(acc)
    sc[inds] = ret['acc_map']
    out['acc_map'] = torch.zeros_like(acc) * acc + sc * (1 - acc)
    return out


def render_ground_pass(batch: dict, ret: dict, fr: Frame, W: Weights, cfg: Cfg, probe, dtype, device, tangent_scale=1.0,
                       inds_fn=None) -> dict:
    """The vis_ground_shading branch of Renderer.render (:1079-1101): all H*W pixels, chunked like the human pass
    (the in-place wbounds growth simply continues).
    `inds` (ray k -> image pixel): the reference takes it from batch_aware_indexing = topk(S, sorted=False) of the 0/1 mask
    (net_utils.py:381-389), i.e. it RELIES on torch returning tied values in index order.  CPU torch does not (the
    golden vectors made under the CPU harness are scrambled accordingly); the intended order -- and the product's -- is
    mask.nonzero().  `inds_fn(mask) -> inds` lets the pinning test replay the CPU order."""
    t = lambda a: torch.as_tensor(a).to(device=device, dtype=dtype)
    H, Wd = int(batch['H']), int(batch['W'])
    F_ = H * Wd
    mask = torch.as_tensor(batch['mask_at_box'][0]).reshape(-1).to(device)
    inds = mask.nonzero()[:, 0] if inds_fn is None else inds_fn(mask)
    ray_o, ray_d = get_rays_full(H, Wd, t(batch['cam_K']), t(batch['cam_R']), t(batch['cam_T']))
    acc = torch.ones(F_, dtype=dtype, device=device)
    acc[inds] = 1 - ret['acc_map']
    n_chunks = max(math.ceil(F_ / cfg.render_chunk), 1)
    actual = math.ceil(F_ / n_chunks)
    parts = []
    for c in range(n_chunks):
        sl = slice(c * actual, (c + 1) * actual)
        fr.wbounds[0] -= cfg.bbox_margin
        fr.wbounds[1] += cfg.bbox_margin
        with torch.no_grad():
            parts.append(render_ground(ray_o[sl], ray_d[sl], acc[sl], fr, W, cfg, fr.wbounds, probe, None, tangent_scale))
    ground = {k: torch.cat([p[k] for p in parts]) for k in parts[0]}
    ground.update(ray_o=ray_o, ray_d=ray_d, acc_map=acc, inds=inds)
    return ground


def render_sphere_tracing(batch: dict, sd: dict, cfg: Cfg, dtype=torch.float32, device='cpu', want_lvis=True, ground=False,
                          tangent_scale: float = 1.0, inds_fn=None, main_probe=None) -> dict:
    """sphere_tracing_renderer.Renderer.render (:1066-1115): chunked get_pixel_value with the in-place wbounds growth, then
    alpha_output_ -- or, with ground=True (cfg.vis_ground_shading, vis_novel_light on), the UN-premultiplied human maps plus
    ret['ground'] (the floor pass over all H*W pixels)."""
    W = Weights(sd, dtype, device)
    fr = Frame.from_batch(batch, cfg, dtype, device)
    t = lambda a: torch.as_tensor(a).to(device=device, dtype=dtype)
    ray_o, ray_d, near, far = t(batch['ray_o'][0]), t(batch['ray_d'][0]), t(batch['near'][0]), t(batch['far'][0])
    P = ray_o.shape[0]
    probe = W.env_main if cfg.relight else None
    if main_probe is not None:                                                # cfg.replace_light (:1068-1069)
        probe = torch.as_tensor(main_probe).to(device=device, dtype=dtype)
    n_chunks = max(math.ceil(P / cfg.render_chunk), 1)
    actual = math.ceil(P / n_chunks) if P else 1                              # net_utils.py:323
    rets = []
    for c in range(n_chunks):
        sl = slice(c * actual, (c + 1) * actual)
        fr.wbounds[0] -= cfg.bbox_margin                                      # :1020-1022, in place, per chunk
        fr.wbounds[1] += cfg.bbox_margin
        with torch.no_grad():
            rets.append(render_human(ray_o[sl], ray_d[sl], near[sl], far[sl], fr, W, cfg, fr.wbounds, probe, want_lvis))
    ret = {k: torch.cat([r[k] for r in rets]) for k in rets[0]}
    if ground:
        ret['ground'] = render_ground_pass(batch, ret, fr, W, cfg, probe, dtype, device, tangent_scale, inds_fn)
    else:
        acc = ret['acc_map']
        for k in _BLEND_KEYS:                                                 # alpha_output_ :454-460
            if k in ret:
                ret[k] = ret[k] * (acc if ret[k].ndim == 1 else acc[:, None])
    ret['wbounds_after'] = fr.wbounds.clone()
    return ret


_VISUAL = ['rgb_map', 'acc_map', 'norm_map', 'surf_map', 'bpts_map', 'cpts_map', 'spec_map', 'shade_map', 'depth_map', 'albedo_map',
           'roughness_map']


def render_novel_light(batch: dict, sd: dict, cfg: Cfg, probes: Dict[str, torch.Tensor], dtype=torch.float32,
                       device='cpu', include_main=True, ground=False, tangent_scale: float = 1.0, inds_fn=None) -> Dict[str, dict]:
    """novel_light_sphere_tracing.Renderer.render (:101-221), no rotation: main pass, then one cheap re-shade per env-map
    from the stored maps (acc-premultiplied without ground shading, raw with it) and, with ground=True, the floor
    re-shade + blend_output_ into image-sized (H*W) maps."""
    main = render_sphere_tracing(batch, sd, cfg, dtype, device, want_lvis=True, ground=ground, tangent_scale=tangent_scale, inds_fn=inds_fn)
    W = Weights(sd, dtype, device)
    out = {}
    grd = main.get('ground')
    if include_main:
        if ground:
            out['main'] = blend_output(grd['acc_map'], grd['inds'], grd, {k: main[k] for k in _VISUAL if k in main})
        else:
            out['main'] = {k: main[k] for k in main if k not in ('lvis_map', 'ldot_map', 'ray_o', 'wbounds_after', 'resd_map')}
    for name, probe in probes.items():
        probe = torch.as_tensor(probe).to(device=device, dtype=dtype).reshape(16, 32, 3)
        with torch.no_grad():
            rgb, shade, spec = shade_pixels(main['ray_o'], main['surf_map'], main['norm_map'], main['albedo_map'],
                                            main['roughness_map'][:, None], main['lvis_map'].T, main['ldot_map'].T,
                                            probe, W, cfg)
        out[name] = dict(rgb_map=rgb, shade_map=shade, spec_map=spec)
        if ground:
            human = {k: main[k] for k in _VISUAL if k in main}
            human.update(out[name])
            with torch.no_grad():
                g_rgb, g_alb, g_shade, g_spec = render_ground_novel(grd['ray_d'], grd['albedo_map'], grd['lvis_map'], grd['ldot_map'], W, cfg, probe)
            g = {k: grd[k] for k in _VISUAL if k in grd}
            g.update(rgb_map=g_rgb, albedo_map=g_alb, shade_map=g_shade, spec_map=g_spec)
            out[name] = blend_output(grd['acc_map'], grd['inds'], g, human)
    out['_main_full'] = main
    return out


def render_volume(batch: dict, sd: dict, cfg: Cfg, dtype=torch.float32, device='cpu', n_samples: int = 128,
                  chunk: int = 8192) -> dict:
    """base_renderer.Renderer.render (:115-129): 128 uniform samples, AniSDF raw (16 ch)."""
    W = Weights(sd, dtype, device)
    fr = Frame.from_batch(batch, cfg, dtype, device)
    t = lambda a: torch.as_tensor(a).to(device=device, dtype=dtype)
    ray_o, ray_d = t(batch['ray_o'][0]), t(batch['ray_d'][0])
    near, far = t(batch['near'][0]).clip(min=0.02), t(batch['far'][0]).clip(max=10.0)
    tv = torch.linspace(0., 1., n_samples).to(near)
    outs = []
    for s in range(0, ray_o.shape[0], chunk):
        sl = slice(s, s + chunk)
        z = near[sl, None] * (1. - tv) + far[sl, None] * tv
        pts = ray_o[sl, None] + ray_d[sl, None] * z[..., None]
        n = pts.shape[0]
        raw = network_forward(pts.reshape(-1, 3), ray_d[sl, None].expand(-1, n_samples, -1).reshape(-1, 3), fr, W, cfg)
        raw = raw.reshape(n, n_samples, -1)
        w = volume_weights(raw[..., -1])
        val = (w[..., None] * raw[..., :-1]).sum(1)
        outs.append(dict(cpts_map=val[:, 0:3], bpts_map=val[:, 3:6], resd_map=val[:, 6:9], norm_map=val[:, 9:12],
                         rgb_map=val[:, 12:15], acc_map=w.sum(-1), depth_map=(w * z).sum(-1)))
    return {k: torch.cat([o[k] for o in outs]) for k in outs[0]}


# ---------------------------------------------------------------------------------------------- batch preparation (row f1)
def rodrigues(r: torch.Tensor) -> torch.Tensor:
    """smplx.lbs.batch_rodrigues (smplx is an unpinned dependency, requirements.txt:19; called at net_utils.py:1163-1172):
    angle = |r + 1e-8|, axis = r / angle, R = I + sin K + (1 - cos) K K.   (n,3) -> (n,3,3)"""
    angle = torch.norm(r + 1e-8, dim=1, keepdim=True)
    d = r / angle
    cos, sin = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    rx, ry, rz = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    z = torch.zeros_like(rx)
    K = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).view(-1, 3, 3)
    return torch.eye(3, dtype=r.dtype)[None] + sin * K + (1 - cos) * torch.bmm(K, K)


def rigid_transform(poses: torch.Tensor, joints: torch.Tensor, parents) -> torch.Tensor:
    """get_rigid_transform (net_utils.py:1163-1172) = smplx.lbs.batch_rigid_transform: chain of [R_j | t_j - t_parent] and
    A_j = G_j with the rest joint taken out.  poses (J,3), joints (J,3) -> A (J,4,4)."""
    R = rodrigues(poses.reshape(-1, 3))
    J = joints.shape[0]
    G = []
    for j in range(J):
        p = int(parents[j])
        L = torch.eye(4, dtype=poses.dtype)
        L[:3, :3] = R[j]
        L[:3, 3] = joints[j] - (joints[p] if p >= 0 else 0.0)
        G.append(L if p < 0 else G[p] @ L)
    G = torch.stack(G)
    A = G.clone()
    A[:, :3, 3] = G[:, :3, 3] - torch.einsum('jab,jb->ja', G[:, :3, :3], joints)
    return A


def vertex_normals(verts: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """pytorch3d Meshes.verts_normals_packed (called at base_dataset.py:380-381): area-weighted face normals accumulated
    per corner, F.normalize(eps=1e-6)."""
    n = torch.zeros_like(verts)
    vf = verts[faces]
    n = n.index_add(0, faces[:, 1], torch.cross(vf[:, 2] - vf[:, 1], vf[:, 0] - vf[:, 1], dim=1))
    n = n.index_add(0, faces[:, 2], torch.cross(vf[:, 0] - vf[:, 2], vf[:, 1] - vf[:, 2], dim=1))
    n = n.index_add(0, faces[:, 0], torch.cross(vf[:, 1] - vf[:, 0], vf[:, 2] - vf[:, 0], dim=1))
    return F.normalize(n, eps=1e-6, dim=1)


def prepare_pose(poses, Rh, Th, tjoints, parents, rverts, weights, rnorm=None, faces=None, pad: float = 0.05) -> dict:
    """get_lbs_params / get_blend (base_dataset.py:308-397, the LBS branch :337-343): float32 torch on the CPU."""
    t = lambda a: torch.as_tensor(a, dtype=torch.float32)
    poses, Rh, Th, tjoints, rverts, weights = t(poses).reshape(-1, 3), t(Rh).reshape(1, 3), t(Th).reshape(1, 3), t(tjoints), t(rverts), t(weights)
    A = rigid_transform(poses, tjoints, parents)
    R = rodrigues(Rh)[0]
    A_bw = torch.einsum('nj,jab->nab', weights, A)                                   # blend_transform  blend_utils.py:212-218
    pverts = torch.sum(A_bw[:, :3, :3] * rverts[:, None], dim=-1) + A_bw[:, :3, 3]   # tpose_points_to_pose_points :303-313
    wverts = pverts @ R.mT + Th                                                      # pose_points_to_world_points :264-273
    if faces is not None:
        pnorm = vertex_normals(pverts, torch.as_tensor(faces, dtype=torch.long))
        wnorm = vertex_normals(wverts, torch.as_tensor(faces, dtype=torch.long))
    else:
        pn = torch.sum(A_bw[:, :3, :3] * t(rnorm)[:, None], dim=-1)
        pnorm = pn / (pn.norm(dim=-1, keepdim=True) + 1e-12)
        wnorm = pnorm @ R.mT
    bounds = lambda x: torch.stack([x.min(0)[0] - pad, x.max(0)[0] + pad])            # get_bounds  data_utils.py:1241-1248
    return dict(A=A, R=R, pverts=pverts, pnorm=pnorm, wverts=wverts, wnorm=wnorm, pbounds=bounds(pverts), wbounds=bounds(wverts))


def rays_within_bounds(H: int, W: int, K, R, T, bounds) -> dict:
    """get_rays_within_bounds (data_utils.py:925-938) = get_rays (:827-845) + get_near_far / get_full_near_far (:848-875),
    in numpy float32 like the reference."""
    import numpy as np
    K, R, T, bounds = (np.asarray(a, np.float32) for a in (K, R, T, bounds))
    ray_o = -np.dot(R.T, T).ravel()
    i, j = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing='ij')
    xy1 = np.stack([j, i, np.ones_like(i)], axis=2)
    pixel_camera = np.dot(xy1, np.linalg.inv(K).T)
    pixel_world = np.dot(pixel_camera - T.ravel(), R)
    ray_d = pixel_world - ray_o[None, None]
    ray_d = ray_d / np.linalg.norm(ray_d, axis=2, keepdims=True)
    ray_o = np.broadcast_to(ray_o, ray_d.shape).reshape(-1, 3).astype(np.float32)
    ray_d = ray_d.reshape(-1, 3).astype(np.float32)
    norm_d = np.linalg.norm(ray_d, axis=-1, keepdims=True)
    viewdir = ray_d / norm_d
    viewdir[(viewdir < 1e-5) & (viewdir > -1e-10)] = 1e-5
    viewdir[(viewdir > -1e-5) & (viewdir < 1e-10)] = -1e-5
    tmin = (bounds[:1] - ray_o[:1]) / viewdir
    tmax = (bounds[1:2] - ray_o[:1]) / viewdir
    near = np.max(np.minimum(tmin, tmax), axis=-1)
    far = np.min(np.maximum(tmin, tmax), axis=-1)
    mask = near < far
    near, far = near / norm_d[..., 0], far / norm_d[..., 0]
    near, far = near[mask] / norm_d[mask, 0], far[mask] / norm_d[mask, 0]
    return dict(ray_o=ray_o[mask], ray_d=ray_d[mask], near=near.astype(np.float32), far=far.astype(np.float32), mask_at_box=mask.reshape(H, W))


def rotate_probe(probe: torch.Tensor, j: int, repeat: int, probe_width: Optional[int] = None) -> torch.Tensor:
    """rotate_envmap's shift_image applied to the probe (relight_utils.py:55-103): (eH,eW,3) -> (eH,eW,3); with `probe_width` = eW it
    is the shift of the attached env-map IMAGE (iH,iW,3): image_shift = iW / (eW * repeat) * j  (:74-75)."""
    image = probe[None]
    B, H, W = image.shape[:3]
    shift = W / ((probe_width or W) * repeat) * j
    i, jj = torch.meshgrid(torch.arange(0, H, device=image.device), torch.arange(0, W, device=image.device), indexing='ij')
    grid = torch.stack([jj, i], dim=-1)[None].expand(B, H, W, 2).float() + 0.5
    grid = grid.clone()
    grid[..., 0] = grid[..., 0] + shift
    grid[..., 0] = grid[..., 0] % W
    grid[..., 0] = grid[..., 0] / W * 2 - 1
    grid[..., 1] = grid[..., 1] / H * 2 - 1
    return F.grid_sample(image.permute(0, 3, 1, 2), grid, align_corners=False, mode='bilinear', padding_mode='border').permute(0, 2, 3, 1)[0]


def assemble_image(batch: dict, rgb_map: torch.Tensor) -> torch.Tensor:
    """Ray -> image scatter (base_visualizer.py:182-202), background 0."""
    mask = torch.as_tensor(batch['mask_at_box'][0])
    img = torch.zeros(*mask.shape, rgb_map.shape[-1], dtype=rgb_map.dtype)
    img[mask] = rgb_map.cpu()
    return img


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    """base_evaluator.py:26-29."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float('inf') if mse == 0 else -10 * math.log10(mse)


# ---------------------------------------------------------------------------------------------- row f3: image assembly
@dataclass
class VisCfg:
    """cfg values Visualizer.generate_image reads (lib/config/config.py:41-46,92,354,395-398,416)."""
    bg_brightness: float = 0.0
    store_alpha_channel: bool = True
    probe_size_ratio: float = 0.2
    min_clip: float = 1.0
    normalize_shading: bool = False
    normalize_specular: bool = True
    tonemapping_albedo: bool = True
    env_h: int = 16
    env_w: int = 32


def gen_light_xyz_dirs(h: int, w: int, r: float = 1e2) -> torch.Tensor:
    """gen_light_xyz (relight_utils.py:423-452), positions only."""
    lat_half, lng_half = torch.pi / h / 2, 2 * torch.pi / w / 2
    lats = torch.linspace(torch.pi / 2 - lat_half, -torch.pi / 2 + lat_half, h)
    lngs = torch.linspace(torch.pi - lng_half, -torch.pi + lng_half, w)
    lngs, lats = torch.meshgrid(lngs, lats, indexing='xy')
    return torch.stack([r * torch.cos(lats) * torch.cos(lngs), r * torch.cos(lats) * torch.sin(lngs), r * torch.sin(lats)], -1)


def gen_light_dir(H: int, W: int, cam_R: torch.Tensor) -> torch.Tensor:
    """relight_utils.py:9-35: overlay pixel -> world direction; cam_R (3,3) world -> camera."""
    R = cam_R.reshape(3, 3).clone().mT.clone()
    front = R[:, 2]
    down = torch.zeros_like(R[:, 1])
    down[2] = torch.sign(R[:, 1][2])
    right = normalize(torch.linalg.cross(down, front))
    front = normalize(torch.linalg.cross(right, down))
    R[:, 0], R[:, 1], R[:, 2] = right, down, front
    R[:, 1], R[:, 2] = -R[:, 2].clone(), -R[:, 1].clone()
    return normalize(gen_light_xyz_dirs(H, W)) @ R.mT


def kth_percentile(x: torch.Tensor, k: int, largest: bool) -> torch.Tensor:
    """The reference's "simple version of percentile": topk(k)[0].max() / .min()   base_visualizer.py:107-108."""
    v = x.ravel().topk(k, largest=largest)[0]
    return v.min() if largest else v.max()


def generate_image(output: dict, batch: dict, vtype: str, vc: VisCfg) -> torch.Tensor:
    """Visualizer.generate_image (base_visualizer.py:55-231) for one light's maps ((P,C) tensors, B squeezed): (H,W,3|4) float32.
    `output['envmap']` (eh,ew,3) or None; batch: mask_at_box (1,H,W), cam_R (1,3,3), tbounds (1,2,3)."""
    mask = torch.as_tensor(batch['mask_at_box'][0]).bool()
    H, W = mask.shape
    g = lambda k: torch.as_tensor(output[k]).float().cpu()
    acc = g('acc_map')
    t = vtype.lower()
    if t == 'normal':
        n = normalize(g('norm_map')) @ torch.as_tensor(batch['cam_R'][0]).float().mT
        n = n * torch.tensor([1.0, -1.0, -1.0])
        rgb = (n * 0.5 + 0.5) * acc[:, None]
    elif t == 'alpha':
        rgb = acc[:, None].expand(-1, 3)
    elif t == 'depth':
        d = g('depth_map')
        k = int(0.01 * d.numel())
        lo = kth_percentile(d[acc.bool()], k, False).clip(None, vc.min_clip)
        hi = kth_percentile(d[acc.bool()], k, True)
        rgb = ((d - lo) / (hi - lo)).clip(0, 1)[:, None].expand(-1, 3)
    elif t in ('shading', 'specular'):
        rgb = g('shade_map' if t == 'shading' else 'spec_map')
        if vc.normalize_shading if t == 'shading' else vc.normalize_specular:
            rgb = rgb / kth_percentile(rgb, int(0.005 * rgb.numel()), True)
    elif t == 'albedo':
        rgb = linear2srgb(g('albedo_map')) if vc.tonemapping_albedo else g('albedo_map')
    elif t == 'roughness':
        rgb = g('roughness_map')[:, None].expand(-1, 3)
    elif t == 'surface':
        tb = torch.as_tensor(batch['tbounds'][0]).float()
        rgb = g('cpts_map') if 'cpts_map' in output else g('surf_map')
        rgb = acc[:, None] * ((rgb - tb[0:1]) / (tb[1:2] - tb[0:1]))
    elif t == 'residual':
        d = g('cpts_map') - g('bpts_map')
        rgb = acc[:, None] * (d / kth_percentile(d, int(0.005 * d.numel()), True))
    elif t == 'rendering':
        rgb = g('rgb_map')
    else:
        raise NotImplementedError(vtype)
    img = torch.ones(H, W, 3) * vc.bg_brightness
    img[mask] = rgb
    env = output.get('envmap')
    if vc.probe_size_ratio > 0 and env is not None:                      # add_light_probe (relight_utils.py:38-52)
        uW = int(W * vc.probe_size_ratio)
        uH = int(uW * vc.env_h / vc.env_w)
        dirs = gen_light_dir(uH, uW, torch.as_tensor(batch['cam_R'][0]).float())
        img[:uH, :uW] = sample_envmap(torch.as_tensor(env).float().cpu(), dirs.reshape(-1, 3)).reshape(uH, uW, 3)
    if vc.store_alpha_channel:
        alpha = torch.zeros(H, W, 1)
        alpha[mask] = acc[:, None]
        img = torch.cat([img, alpha], -1)
    return img


def save_image_pixels(img: torch.Tensor, ext: str):
    """What save_image (data_utils.py:689-709) hands to cv2.imwrite: BGR order; .png -> uint16, .jpg -> 3-channel uint8."""
    import numpy as np
    a = img.numpy().copy()
    a[..., :3] = a[..., [2, 1, 0]]
    if ext == '.png':
        return (a * 65535).clip(0, 65535).astype(np.uint16)
    if ext == '.jpg':
        return (a[..., :3] * 255).clip(0, 255).astype(np.uint8)
    return a[..., :3] if ext == '.hdr' else a
