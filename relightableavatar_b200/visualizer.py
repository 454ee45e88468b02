"""Device-side mirror of the reference's image assembly.

    reference                                                        here
    Visualizer.generate_image(output, batch, type)                   Visualizer.generate_image(output, batch, type)  (same return: (H,W,C) float32 numpy)
      lib/visualizers/base_visualizer.py:55-231
    add_light_probe (lib/utils/relight_utils.py:38-52)               the overlay of ra_assemble_visual
    save_image's BGR swap + 16-bit png / 8-bit jpg quantisation      Visualizer.encode / encode_frame -> finished uint8 / uint16 pixels
      lib/utils/data_utils.py:689-709
    per-light `to_cpu(human)` of fp32 maps                           ONE device -> host copy of the finished pixels of a whole frame
      novel_light_sphere_tracing.py:216                              (every light x every output type), into pinned memory, asynchronously

Everything per pixel runs in the CUDA library (ra_visual_map, ra_assemble_visual); torch provides memory and streams.
File writing / video encoding (cv2.imwrite, ffmpeg) stay with the caller: they are the consumers of this path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import VIS_TYPES, VISUAL_INPUTS, ra_image_config, ra_visual_config, ra_visual_inputs
from .renderer import _fptr, _ptr

DEFAULTS = dict(bg_brightness=0.0, store_alpha_channel=True, probe_size_ratio=0.2, min_clip=1.0, normalize_shading=False,
                normalize_specular=True, tonemapping_albedo=True, vis_ext='.jpg', env_h=16, env_w=32)      # lib/config/config.py:41-46,354,395-398,416


def _normalize(x: torch.Tensor) -> torch.Tensor:
    return x / (x.norm(dim=-1, keepdim=True) + 1e-8)          # net_utils.py:1626-1628


def gen_light_dir(uH: int, uW: int, cam_R: torch.Tensor) -> torch.Tensor:
    """World-space directions of the light-probe overlay pixels (relight_utils.py:9-35): the lat-long grid of gen_light_xyz
    (:423-452) seen from a camera that keeps only its horizontal rotation.  cam_R: (3,3) world -> camera.  Host-side, uH*uW*3 floats."""
    R = cam_R.detach().to('cpu', torch.float32).reshape(3, 3).clone().mT.clone()       # c2w: columns = camera axes
    front = R[:, 2]
    down = torch.zeros(3)
    down[2] = torch.sign(R[:, 1][2])
    right = _normalize(torch.linalg.cross(down, front))
    front = _normalize(torch.linalg.cross(right, down))
    R[:, 0], R[:, 1], R[:, 2] = right, down, front
    R[:, 1], R[:, 2] = -R[:, 2].clone(), -R[:, 1].clone()
    lat_half, lng_half = torch.pi / uH / 2, 2 * torch.pi / uW / 2
    lats = torch.linspace(torch.pi / 2 - lat_half, -torch.pi / 2 + lat_half, uH)
    lngs = torch.linspace(torch.pi - lng_half, -torch.pi + lng_half, uW)
    lngs, lats = torch.meshgrid(lngs, lats, indexing='xy')
    r = 1e2
    xyz = torch.stack([r * torch.cos(lats) * torch.cos(lngs), r * torch.cos(lats) * torch.sin(lngs), r * torch.sin(lats)], -1)
    return (_normalize(xyz) @ R.mT).contiguous()


class Visualizer:
    """`Visualizer(engine, cfg)`: `cfg` is the reference's cfg object (or None for its defaults; keyword overrides win)."""

    def __init__(self, engine, cfg=None, **over):
        self.engine = engine
        v = dict(DEFAULTS)
        if cfg is not None:
            for k in DEFAULTS:
                try:
                    if k in cfg:
                        v[k] = cfg[k]
                except TypeError:
                    if hasattr(cfg, k):
                        v[k] = getattr(cfg, k)
        v.update(over)
        self.v = v
        self._dirs: Dict[tuple, torch.Tensor] = {}
        self._pinned: Dict[tuple, torch.Tensor] = {}
        self._copy_stream: Optional[torch.cuda.Stream] = None

    # ------------------------------------------------------------------ helpers
    def _dev(self, t, dtype=torch.float32):
        return torch.as_tensor(t).to(device=self.engine.device, dtype=dtype).contiguous()

    @staticmethod
    def _hw(batch) -> Tuple[int, int]:
        meta = batch.get('meta') or {}
        H = meta['H'] if 'H' in meta else batch['H']
        W = meta['W'] if 'W' in meta else batch['W']
        return int(torch.as_tensor(H).reshape(-1)[0]), int(torch.as_tensor(W).reshape(-1)[0])

    def _probe_dirs(self, uH, uW, cam_R) -> torch.Tensor:
        key = (uH, uW, tuple(np.asarray(torch.as_tensor(cam_R).detach().cpu(), np.float32).ravel().tolist()))
        d = self._dirs.get(key)
        if d is None:
            if len(self._dirs) > 64:
                self._dirs.clear()
            d = self._dirs[key] = gen_light_dir(uH, uW, torch.as_tensor(cam_R)).to(self.engine.device)
        return d

    def visual_map(self, output, batch, type: str) -> torch.Tensor:
        """generate_image's per-type `rgb_map` in ray order, (n,3) on the device (base_visualizer.py:58-176)."""
        eng = self.engine
        t = type.lower()
        if t not in VIS_TYPES:
            raise NotImplementedError(f'Not implemented output type: {type}')            # Semantic / Feature: deprecated in the reference
        vin, keep = ra_visual_inputs(), []
        n = None
        for k in VISUAL_INPUTS[:-2]:
            if k in output and output[k] is not None:
                m = self._dev(output[k])[0].contiguous()
                keep.append(m)
                setattr(vin, k, _fptr(m))
                n = m.shape[0] if n is None else n
        if t == 'normal':
            m = self._dev(batch['cam_R']).reshape(3, 3).contiguous(); keep.append(m); vin.cam_R = _fptr(m)
        if t == 'surface':
            m = self._dev(batch['tbounds']).reshape(2, 3).contiguous(); keep.append(m); vin.tbounds = _fptr(m)
        vc = ra_visual_config(float(self.v['min_clip']), int(bool(self.v['normalize_shading'])), int(bool(self.v['normalize_specular'])),
                              int(bool(self.v['tonemapping_albedo'])))
        out = torch.empty(n or 0, 3, device=eng.device)
        with torch.cuda.device(eng.device):
            eng._check(eng.lib.ra_visual_map(eng.h, VIS_TYPES[t], C.byref(vin), int(n or 0), C.byref(vc), _ptr(out), eng._stream()), 'ra_visual_map')
        return out

    def _assemble(self, vmap, output, batch, type: str, fmt: str, bgr: bool, alpha: bool = True):
        eng = self.engine
        H, W = self._hw(batch)
        mask = torch.as_tensor(batch['mask_at_box']).to(eng.device).reshape(H, W).to(torch.uint8).contiguous()
        acc = self._dev(output['acc_map'])[0].contiguous() if (self.v['store_alpha_channel'] and alpha) else None
        ic = ra_image_config()
        ic.bg_brightness, ic.channels, ic.bgr = float(self.v['bg_brightness']), 4 if acc is not None else 3, int(bgr)
        keep = []
        env = output.get('envmap')
        if self.v['probe_size_ratio'] > 0 and env is not None:                              # add_light_probe (relight_utils.py:38-52)
            probe = self._dev(env['probe'])
            probe = (probe[0] if probe.ndim == 4 else probe).contiguous()
            uW = int(W * self.v['probe_size_ratio'])
            uH = int(uW * self.v['env_h'] / self.v['env_w'])
            dirs = self._probe_dirs(uH, uW, batch['cam_R'])
            keep += [probe, dirs]
            ic.probe, ic.eh, ic.ew, ic.probe_dirs, ic.uH, ic.uW = _fptr(probe), probe.shape[0], probe.shape[1], _fptr(dirs), uH, uW
        outs = {'f32': None, 'u8': None, 'u16': None}
        # 16-bit pixels live in int16 storage (same bits; torch's uint16 has almost no CUDA operators) and are viewed as uint16 at the end
        outs[fmt] = torch.empty(H, W, ic.channels, device=eng.device, dtype={'f32': torch.float32, 'u8': torch.uint8, 'u16': torch.int16}[fmt])
        with torch.cuda.device(eng.device):
            eng._check(eng.lib.ra_assemble_visual(eng.h, _ptr(vmap), _ptr(acc), _ptr(mask), H, W, C.byref(ic), _ptr(outs['f32']), _ptr(outs['u8']),
                                                  _ptr(outs['u16']), eng._stream()), 'ra_assemble_visual')
        return outs[fmt]

    # ------------------------------------------------------------------ the reference's surface
    def generate_image_device(self, output, batch, type: str = 'rendering', fmt: str = 'f32', bgr: bool = False, alpha: bool = True) -> torch.Tensor:
        """(H, W, 3|4) image on the device: fp32 (`fmt='f32'`, what generate_image returns) or quantised like save_image
        (`'u8'`: (v*255).clip(0,255); `'u16'`: (v*65535).clip(0,65535), held in int16 storage); `bgr` applies save_image's channel
        swap; `alpha=False` drops the alpha channel (save_image does for .jpg / .hdr)."""
        if type.lower() == 'envmap':                                                        # :162-163: the probe itself, no overlay / alpha
            img = self._dev(output['envmap']['probe'])
            img = img[0] if img.ndim == 4 else img
            if bgr:
                img = img[..., [2, 1, 0]]
            if fmt == 'f32':
                return img.contiguous()
            q = 255.0 if fmt == 'u8' else 65535.0
            qi = (img * q).clip(0, q).to(torch.int32)
            return qi.to(torch.uint8) if fmt == 'u8' else qi.to(torch.int16)
        img = self._assemble(self.visual_map(output, batch, type), output, batch, type, fmt, bgr, alpha)
        if 'orig_H' in batch and 'orig_W' in batch:                                         # fill_image (:222-229): paste the crop into the full frame
            oH, oW = int(torch.as_tensor(batch['orig_H']).reshape(-1)[0]), int(torch.as_tensor(batch['orig_W']).reshape(-1)[0])
            bb = torch.as_tensor(batch['crop_bbox'])[0].to('cpu', torch.int64)
            if img.shape[-1] != 3:
                raise ValueError('fill_image pastes 3-channel images (cfg.store_alpha_channel must be off with cropped rendering)')
            bgv = float(self.v['bg_brightness']) * {'f32': 1.0, 'u8': 255.0, 'u16': 65535.0}[fmt]
            full = torch.full((oH, oW, 3), bgv, device=img.device, dtype=torch.float32).to(img.dtype)
            h, w = int(bb[1, 1] - bb[0, 1]), int(bb[1, 0] - bb[0, 0])
            full[int(bb[0, 1]):int(bb[1, 1]), int(bb[0, 0]):int(bb[1, 0])] = img[:h, :w]
            img = full
        return img

    def generate_image(self, output, batch, type: str = 'rendering') -> np.ndarray:
        """Same contract as the reference's static method: the predicted image as an (H, W, C) float32 numpy array."""
        return self.generate_image_device(output, batch, type, 'f32').cpu().numpy()

    def encode(self, output, batch, type: str = 'rendering', ext: Optional[str] = None) -> torch.Tensor:
        """What save_image hands to cv2.imwrite, on the device: BGR(A) uint16 for .png, BGR uint8 for .jpg, BGR(A) fp32 otherwise."""
        ext = self.v['vis_ext'] if ext is None else ext
        if ext == '.png':
            return self.generate_image_device(output, batch, type, 'u16', bgr=True)          # int16 storage of the uint16 pixels
        if ext == '.jpg':
            return self.generate_image_device(output, batch, type, 'u8', bgr=True, alpha=False)
        return self.generate_image_device(output, batch, type, 'f32', bgr=True, alpha=ext != '.hdr')

    def encode_frame(self, outputs: Dict[str, dict], batch, types: Sequence[str] = ('rendering',), ext: Optional[str] = None,
                     names: Optional[Iterable[str]] = None, wait: bool = True) -> Dict[Tuple[str, str], np.ndarray]:
        """Every light x every output type of one frame -> finished pixels in pinned host memory with ONE device -> host copy
        (replaces the per-light `to_cpu` of fp32 maps, novel_light_sphere_tracing.py:216, and the per-image numpy quantisation).
        `outputs`: what Renderer.render returned (light name -> maps; the float 'diff' entry is skipped).  Returns views into a
        pinned buffer that is reused by the next call with the same layout; with `wait=False` the copy is only enqueued
        (on a side stream) and `self.copy_done.synchronize()` must precede the first read."""
        eng = self.engine
        ext = self.v['vis_ext'] if ext is None else ext
        names = [n for n in (names if names is not None else outputs.keys()) if isinstance(outputs.get(n), dict)]
        imgs, keys = [], []
        for n in names:
            for t in types:
                imgs.append(self.encode(outputs[n], batch, t, ext).contiguous())
                keys.append((n, t))
        if not imgs:
            return {}
        dt = imgs[0].dtype
        sizes = [int(i.numel()) for i in imgs]
        flat = torch.cat([i.reshape(-1) for i in imgs])                                      # one staging buffer on the device
        key = (dt, flat.numel())
        host = self._pinned.get(key)
        if host is None:
            if len(self._pinned) > 8:
                self._pinned.clear()
            host = self._pinned[key] = torch.empty(flat.numel(), dtype=dt).pin_memory()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(eng.device)
        ready = torch.cuda.Event(); ready.record(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            host.copy_(flat, non_blocking=True)
            flat.record_stream(self._copy_stream)
            self.copy_done = torch.cuda.Event(); self.copy_done.record(self._copy_stream)
        self.d2h_bytes = flat.numel() * flat.element_size()
        if wait:
            self.copy_done.synchronize()
        out, o = {}, 0
        hn = host.numpy().view(np.uint16) if dt == torch.int16 else host.numpy()
        for k, i, s in zip(keys, imgs, sizes):
            out[k] = hn[o:o + s].reshape(tuple(i.shape))
            o += s
        return out
