"""Multi-GPU sharding of the rendering path: one process per GPU, torch.distributed (NCCL over NVLink) as plumbing.

Every ray is independent given the per-frame read-only state, so there is no data-path collective
inside the renderer.  Two partitions are offered, both ending in ONE all-gather of finished pixels:

* frame sharding (BASELINE config 5; weak scaling): rank r renders frames r, r+W, r+2W, ... of a sequence;
* tile sharding (BASELINE config 4; strong scaling): the P in-box rays of one frame are dealt to the ranks in
  interleaved blocks of 32 rays (foreground/background and lit/unlit work balance), every rank renders its rays for
  all env-maps, and the fixed-size padded pixel blocks are all-gathered and de-interleaved.

The partition logic is pure index arithmetic and is covered on CPU with the gloo backend (tests/test_parallel.py).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

BLOCK = 32


def frame_indices(n_frames: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_frames, world))


_PART_CACHE: Dict[tuple, torch.Tensor] = {}


def tile_partition(P: int, rank: int, world: int, block: int = BLOCK, device=None) -> torch.Tensor:
    """Indices (ascending) of the rays owned by `rank`: blocks of `block` consecutive rays dealt round-robin.
    Cached per (P, rank, world, block, device): a sequence re-uses the same partition every frame."""
    key = (P, rank, world, block, str(device))
    t = _PART_CACHE.get(key)
    if t is None:
        idx = torch.arange(P)
        t = idx[(idx // block) % world == rank]
        if device is not None:
            t = t.to(device)
        if len(_PART_CACHE) > 256:
            _PART_CACHE.clear()
        _PART_CACHE[key] = t
    return t


def padded_count(P: int, world: int, block: int = BLOCK) -> int:
    """Upper bound of any rank's ray count (the all-gather block size)."""
    n_blocks = (P + block - 1) // block
    return ((n_blocks + world - 1) // world) * block


def shard_batch_rays(batch: Dict, rank: int, world: int) -> Tuple[Dict, torch.Tensor]:
    """Returns a shallow copy of `batch` whose ray tensors hold only this rank's rays, and the owned indices."""
    P = batch['ray_o'].shape[1]
    b = dict(batch)
    own = None
    for k in ('ray_o', 'ray_d', 'near', 'far'):
        t = torch.as_tensor(batch[k])
        own = tile_partition(P, rank, world, device=t.device)
        b[k] = t[:, own]
    return b, own


def allgather_pixels(local: torch.Tensor, n_pad: int, group=None) -> torch.Tensor:
    """local (n_local, C) -> (world, n_pad, C) on every rank with ONE collective (rows beyond n_local are zero)."""
    world = dist.get_world_size(group)
    buf = torch.zeros(n_pad, local.shape[1], dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty(world * n_pad, local.shape[1], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out.view(world, n_pad, local.shape[1])


def deinterleave(gathered: torch.Tensor, P: int, world: int, block: int = BLOCK) -> torch.Tensor:
    """(world, n_pad, C) blocks -> (P, C) in original ray order."""
    out = torch.empty(P, gathered.shape[-1], dtype=gathered.dtype, device=gathered.device)
    for r in range(world):
        own = tile_partition(P, r, world, block, device=gathered.device)
        out[own] = gathered[r, : own.numel()]
    return out


def render_tile_sharded(render_fn, batch: Dict, keys=('rgb_map', 'acc_map'), group=None, engine=None) -> Dict[str, torch.Tensor]:
    """`render_fn(batch) -> {key: (1, P_local, C) or (1, P_local)}` on this rank's rays; returns full-frame (1, P, C) maps.
    `engine` (the renderer's Engine): told the ray layout, so that the one frame-position dependent quantity of the reference
    (the per-chunk wbounds growth of frames with more than render_chunk rays) uses global ray indices."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    P = batch['ray_o'].shape[1]
    local_batch, own = shard_batch_rays(batch, rank, world)
    if engine is not None:
        engine.set_ray_layout(P, BLOCK, world, rank)
    try:
        out = render_fn(local_batch)
    finally:
        if engine is not None:
            engine.set_ray_layout(0, BLOCK, 1, 0)
    cols, widths = [], []
    for k in keys:
        t = out[k][0]
        t = t[:, None] if t.ndim == 1 else t
        cols.append(t); widths.append(t.shape[1])
    packed = torch.cat(cols, dim=1).contiguous()
    full = deinterleave(allgather_pixels(packed, padded_count(P, world), group), P, world)
    res, c0 = {}, 0
    for k, w in zip(keys, widths):
        res[k] = full[:, c0:c0 + w][None] if w > 1 else full[:, c0][None]
        c0 += w
    return res


def render_sequence_sharded(render_frame, n_frames: int, n_pad: int, channels: int = 4, group=None, device=None):
    """Frame sharding (BASELINE config 5): frame f is rendered by rank f % world; every step (world frames) ends with ONE all-gather
    of the ranks' finished pixel blocks, padded to `n_pad` rows.  `render_frame(f) -> (P_f, channels)` pixels of frame f (P_f <= n_pad,
    P_f < 2**24).  Yields `(f, pixels (P_f, channels))` in frame order on every rank.  The ray count travels in an extra leading row
    of the block, so ranks need not know each other's P_f and the step stays a single collective."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    for f0 in range(0, n_frames, world):
        f = f0 + rank
        px = render_frame(f) if f < n_frames else None
        dev = px.device if px is not None else device
        if dev is None:        # an idle rank of the ragged last step still joins the collective: NCCL needs a CUDA buffer
            dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu')
        buf = torch.zeros(n_pad + 1, channels, dtype=torch.float32, device=dev)
        if px is not None:
            if px.shape[0] > n_pad:
                raise ValueError(f'frame {f} has {px.shape[0]} rays, more than n_pad={n_pad}')
            buf[0, 0] = float(px.shape[0])
            buf[1:1 + px.shape[0]] = px
        out = torch.empty(world * (n_pad + 1), channels, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(out, buf, group=group)
        out = out.view(world, n_pad + 1, channels)
        for r in range(min(world, n_frames - f0)):
            n = int(out[r, 0, 0].item())
            yield f0 + r, out[r, 1:1 + n]


class PixelGather:
    """The one collective of a sharded step, taken off the rendering stream.

    Every rank's finished, padded pixel block (n_pad x channels fp32) is all-gathered through the C-ABI's own entry point
    `ra_allgather` (NCCL over NVLink; the communicator is created here through ctypes from a unique id that travels over
    torch.distributed) on a SIDE stream into pre-allocated, double-buffered device memory: step s is gathered while step s+1
    renders, so a rank never stalls on the slowest frame of the current step and no per-step allocation / zero-fill / concat
    kernels run.  `submit(px)` -> slot; `result(slot)` -> (world, n_pad, channels) view, valid until the slot is reused two
    submits later.  Unused rows of a block keep whatever an earlier step left there: callers slice by their own ray counts."""

    def __init__(self, engine, n_pad: int, channels: int = 4, group=None, slots: int = 2):
        import ctypes as C
        self.C = C
        self.engine, self.n_pad, self.channels = engine, int(n_pad), int(channels)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = engine.device
        self.dev = dev
        self.stream = torch.cuda.Stream(dev)
        self.send = [torch.zeros(self.n_pad, channels, device=dev) for _ in range(slots)]
        self.recv = [torch.zeros(self.world * self.n_pad, channels, device=dev) for _ in range(slots)]
        self.done = [None] * slots
        self.step = 0
        self.nccl = C.CDLL('libnccl.so.2')                 # the copy torch already loaded into this process
        uid = (C.c_char * 128)()
        if self.rank == 0:
            rc = self.nccl.ncclGetUniqueId(C.byref(uid))
            if rc != 0:
                raise RuntimeError(f'ncclGetUniqueId returned {rc}')
        t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0, group=group)

        class _Uid(C.Structure):
            _fields_ = [('internal', C.c_char * 128)]

        u = _Uid()
        C.memmove(C.byref(u), bytes(t.cpu().tolist()), 128)
        self.comm = C.c_void_p()
        self.nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _Uid, C.c_int]
        with torch.cuda.device(dev):
            rc = self.nccl.ncclCommInitRank(C.byref(self.comm), self.world, u, self.rank)
        if rc != 0:
            raise RuntimeError(f'ncclCommInitRank returned {rc}')

    def submit(self, px: torch.Tensor) -> int:
        """px (n <= n_pad, channels) on the engine's device, produced on the current stream."""
        C = self.C
        slot = self.step % len(self.send)
        self.step += 1
        if self.done[slot] is not None:
            torch.cuda.current_stream(self.dev).wait_event(self.done[slot])       # the slot's previous gather has drained
        self.send[slot][: px.shape[0]].copy_(px, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.dev))
        self.stream.wait_event(ready)
        eng = self.engine
        with torch.cuda.device(self.dev):
            eng._check(eng.lib.ra_allgather(eng.h, self.comm, C.c_void_p(self.send[slot].data_ptr()), self.n_pad * self.channels,
                                            C.c_void_p(self.recv[slot].data_ptr()), C.c_void_p(self.stream.cuda_stream)), 'ra_allgather')
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self.done[slot] = ev
        return slot

    def result(self, slot: int) -> torch.Tensor:
        """Makes the current stream wait for the slot's gather and returns (world, n_pad, channels)."""
        torch.cuda.current_stream(self.dev).wait_event(self.done[slot])
        return self.recv[slot].view(self.world, self.n_pad, self.channels)

    def drain(self) -> None:
        for ev in self.done:
            if ev is not None:
                torch.cuda.current_stream(self.dev).wait_event(ev)

    def close(self) -> None:
        if getattr(self, 'comm', None):
            torch.cuda.synchronize(self.dev)
            self.nccl.ncclCommDestroy.argtypes = [self.C.c_void_p]
            self.nccl.ncclCommDestroy(self.comm)
            self.comm = None


class FramesInFlight:
    """Several independent frames in flight on ONE GPU (sequence rendering: the loop of `run.py -t visualize`, run.py:68-85).

    A frame's surface-tracing stage is 17 dependent tracer / MLP iterations over <= 70 k rays and its attribute pass a chain of 48 small
    GEMMs: both are latency-bound and leave most SMs idle, while the visibility stage of another frame is tensor-bound.  `n` handles
    (each with its own workspaces) on `n` streams render consecutive frames concurrently; results come back in submission order.
    Measured on the 512^2 relight workload: 45.0 -> 47.5 frames/s with two frames in flight, 47.7 with three
    (tools/two_frames_in_flight.py); every frame equals the one-at-a-time rendering bit for bit (a frame never shares state with another).

    `make_renderer()` -> a fresh `Renderer` (its own Engine); `submit(batch)` -> ticket; `result(ticket)` makes the CURRENT stream wait
    for that frame and returns its outputs."""

    def __init__(self, make_renderer, n: int = 2):
        self.renderers = [make_renderer() for _ in range(max(int(n), 1))]
        dev = self.renderers[0].engine.device
        self.device = dev
        self.streams = [torch.cuda.Stream(dev) for _ in self.renderers]
        self.pending = {}
        self.count = 0

    def submit(self, batch):
        k = self.count % len(self.renderers)
        st = self.streams[k]
        st.wait_stream(torch.cuda.current_stream(self.device))        # inputs produced on the caller's stream
        with torch.cuda.stream(st):
            out = self.renderers[k].render(batch)
            ev = torch.cuda.Event()
            ev.record(st)
        ticket = self.count
        self.pending[ticket] = (out, ev)
        self.count += 1
        return ticket

    def result(self, ticket):
        out, ev = self.pending.pop(ticket)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)

        def mark(o):        # the outputs were allocated on the frame's side stream: tell the caching allocator who reads them now
            if torch.is_tensor(o):
                if o.is_cuda:
                    o.record_stream(cur)
            elif isinstance(o, dict):
                for v in o.values():
                    mark(v)
        mark(out)
        return out

    def render_sequence(self, batches):
        """Yields the outputs of `batches` in order, keeping len(self.renderers) frames in flight."""
        tickets = []
        for b in batches:
            tickets.append(self.submit(b))
            if len(tickets) >= len(self.renderers):
                yield self.result(tickets.pop(0))
        for t in tickets:
            yield self.result(t)

    def close(self):
        torch.cuda.synchronize(self.device)
        for r in self.renderers:
            r.engine.close()
        self.renderers = []
