"""GPU per-frame batch preparation: the device-side replacement of what
`pose_dataset.__getitem__` computes on the CPU for every frame (lib/datasets/pose_dataset.py:45-113):

    get_lbs_params / get_blend   base_dataset.py:308-397   Rodrigues + kinematic chain (A), LBS, pose -> world, normals, bounds
    get_rays_within_bounds       data_utils.py:925-938     all-pixel rays, AABB near/far, mask_at_box, compaction

The subject's static data (rest joints, kinematic tree, rest vertices, skinning weights, faces or rest normals, big-pose
vertices and transforms) is uploaded once; per frame only `poses`, `Rh`, `Th` and the camera cross PCIe (< 1 KB instead of
~3.9 MB).  `FramePreparer.make_batch` returns the `batch` dict `Renderer.render` consumes, every tensor already on the device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from ._lib import POSE_OUTPUTS, ra_body, ra_pose_outputs
from .renderer import Engine, _fptr, _ptr


class FramePreparer:
    def __init__(self, engine: Engine, tjoints, parents, rverts, weights, big_A, tverts, rnorm=None, faces=None, tnorm=None,
                 bounds_pad: float = 0.05):
        self.eng = engine
        dev = engine.device
        f = lambda a: torch.as_tensor(np.asarray(a)).to(device=dev, dtype=torch.float32).contiguous()
        i32 = lambda a: torch.as_tensor(np.asarray(a)).to(device=dev, dtype=torch.int32).contiguous()
        self.N, self.J = engine.config['n_verts'], engine.config['n_bones']
        self.static = dict(tjoints=f(tjoints).reshape(self.J, 3), parents=i32(parents).reshape(self.J), rverts=f(rverts).reshape(self.N, 3),
                           weights=f(weights).reshape(self.N, self.J), big_A=f(big_A).reshape(self.J, 4, 4), tverts=f(tverts).reshape(self.N, 3))
        if tnorm is not None:
            self.static['tnorm'] = f(tnorm).reshape(self.N, 3)
        self.rnorm = f(rnorm).reshape(self.N, 3) if rnorm is not None else None
        self.faces = i32(faces).reshape(-1, 3) if faces is not None else None
        self.bounds_pad = float(bounds_pad)
        b = ra_body()
        b.tjoints = _fptr(self.static['tjoints']); b.rverts = _fptr(self.static['rverts']); b.weights = _fptr(self.static['weights'])
        b.parents = C.cast(C.c_void_p(self.static['parents'].data_ptr()), C.POINTER(C.c_int32))
        b.rnorm = _fptr(self.rnorm)
        b.faces = C.cast(C.c_void_p(self.faces.data_ptr() if self.faces is not None else 0), C.POINTER(C.c_int32))
        b.n_faces = int(self.faces.shape[0]) if self.faces is not None else 0
        with torch.cuda.device(dev):
            engine._check(engine.lib.ra_upload_body(engine.h, C.byref(b), engine._stream()), 'ra_upload_body')

    def pose(self, poses, Rh, Th) -> Dict[str, torch.Tensor]:
        """get_lbs_params + get_blend for one frame: poses (J,3) axis-angle, Rh (3), Th (3)."""
        eng, dev = self.eng, self.eng.device
        up = torch.as_tensor(np.concatenate([np.asarray(poses, np.float32).reshape(-1), np.asarray(Rh, np.float32).reshape(-1),
                                             np.asarray(Th, np.float32).reshape(-1)])).to(dev, non_blocking=True)   # the per-frame upload: 165 floats
        n = self.J * 3
        out = dict(A=torch.empty(self.J, 4, 4, device=dev), R=torch.empty(3, 3, device=dev), pverts=torch.empty(self.N, 3, device=dev),
                   pnorm=torch.empty(self.N, 3, device=dev), wverts=torch.empty(self.N, 3, device=dev), wnorm=torch.empty(self.N, 3, device=dev),
                   pbounds=torch.empty(2, 3, device=dev), wbounds=torch.empty(2, 3, device=dev))
        o = ra_pose_outputs()
        for k in POSE_OUTPUTS:
            setattr(o, k, _fptr(out[k]))
        with torch.cuda.device(dev):
            eng._check(eng.lib.ra_prepare_pose(eng.h, _ptr(up[:n]), _ptr(up[n:n + 3]), _ptr(up[n + 3:n + 6]), self.bounds_pad, C.byref(o), eng._stream()),
                       'ra_prepare_pose')
        out['poses'] = up[:n].reshape(self.J, 3)
        out['Th'] = up[n + 3:n + 6].reshape(1, 3)
        out['_keep'] = up
        return out

    def rays(self, K, R, T, H: int, W: int, wbounds: torch.Tensor) -> Dict[str, torch.Tensor]:
        """get_rays_within_bounds: camera K, R (3,3), T (3,1) on the host; returns compacted rays + mask_at_box (H,W) bool."""
        eng, dev = self.eng, self.eng.device
        Kh, Rh_, Th_ = (np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1)) for a in (K, R, T))
        n = H * W
        ro, rd = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
        near, far = torch.empty(n, device=dev), torch.empty(n, device=dev)
        mask = torch.empty(n, device=dev, dtype=torch.uint8)
        cnt = torch.zeros(1, device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            eng._check(eng.lib.ra_prepare_rays(eng.h, Kh.ctypes.data_as(C.c_void_p), Rh_.ctypes.data_as(C.c_void_p), Th_.ctypes.data_as(C.c_void_p), H, W,
                                               _ptr(wbounds.contiguous()), _ptr(ro), _ptr(rd), _ptr(near), _ptr(far), _ptr(mask), _ptr(cnt), eng._stream()),
                       'ra_prepare_rays')
        P = int(cnt.item())            # the one host read of the frame: how many rays hit the box
        return dict(ray_o=ro[:P], ray_d=rd[:P], near=near[:P], far=far[:P], mask_at_box=mask.reshape(H, W).bool())

    def make_batch(self, poses, Rh, Th, K, R, T, H: int, W: int, extra: Optional[Dict] = None) -> Dict:
        """The `batch` of SURVEY.md 8b for one frame (leading B = 1), all tensors resident on the device."""
        p = self.pose(poses, Rh, Th)
        r = self.rays(K, R, T, H, W, p['wbounds'])
        dev = self.eng.device
        b = {k: v[None] for k, v in r.items()}
        for k in ('A', 'R', 'pverts', 'pnorm', 'wverts', 'wnorm', 'pbounds', 'wbounds', 'poses', 'Th'):
            b[k] = p[k][None]
        for k in ('big_A', 'weights', 'tverts', 'tnorm'):
            if k in self.static:
                b[k] = self.static[k][None]
        f = lambda a: torch.as_tensor(np.asarray(a, np.float32)).to(dev)
        b.update(cam_K=f(K).reshape(1, 3, 3), cam_R=f(R).reshape(1, 3, 3), cam_T=f(T).reshape(1, 3, 1), H=H, W=W)
        if extra:
            b.update(extra)
        return b
