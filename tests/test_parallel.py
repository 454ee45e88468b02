"""N>1 path on CPU: world_size-2 gloo processes exercise the ray partition, the single all-gather and the
de-interleave with a stand-in render function (the GPU renderer itself is covered by -m gpu tests)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relightableavatar_b200 import parallel


def _free_port() -> int:
    import socket
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        return sk.getsockname()[1]


def test_partition_covers_every_ray_once():
    for P in (0, 1, 31, 32, 33, 1000, 69137):
        for world in (1, 2, 4, 8):
            parts = [parallel.tile_partition(P, r, world) for r in range(world)]
            allidx = torch.cat(parts)
            assert allidx.numel() == P and torch.equal(torch.sort(allidx)[0], torch.arange(P))
            assert max([p.numel() for p in parts] + [0]) <= parallel.padded_count(P, world)
    assert parallel.frame_indices(10, 1, 4) == [1, 5, 9]


def _worker(rank, world, port, P):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    batch = dict(ray_o=torch.randn(1, P, 3, generator=g), ray_d=torch.randn(1, P, 3, generator=g),
                 near=torch.rand(1, P, generator=g), far=torch.rand(1, P, generator=g))

    def fake_render(b):      # per-ray function of the inputs: independent of the partition
        return dict(rgb_map=b['ray_o'] * 2 + b['ray_d'], acc_map=b['near'] - b['far'])

    out = parallel.render_tile_sharded(fake_render, batch)
    ref = fake_render(batch)
    assert torch.equal(out['rgb_map'], ref['rgb_map']) and torch.equal(out['acc_map'], ref['acc_map'])
    dist.destroy_process_group()


def test_tile_sharded_render_world2_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 1000), nprocs=2, join=True)


def _seq_worker(rank, world, port, n_frames):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rendered = []

    def render_frame(f):          # frame f has 10 + 3 f rays; pixel values identify (frame, ray)
        rendered.append(f)
        n = 10 + 3 * f
        return torch.arange(n, dtype=torch.float32)[:, None] * torch.ones(1, 4) + 1000 * f

    got = list(parallel.render_sequence_sharded(render_frame, n_frames, n_pad=64, device=torch.device('cpu')))
    assert rendered == parallel.frame_indices(n_frames, rank, world)          # each rank renders only its own frames
    assert [f for f, _ in got] == list(range(n_frames))                        # and every rank ends up with all of them, in order
    for f, px in got:
        assert px.shape == (10 + 3 * f, 4) and float(px[0, 0]) == 1000 * f and float(px[-1, 3]) == 1000 * f + 9 + 3 * f
    dist.destroy_process_group()


def test_frame_sharded_sequence_world2_gloo():
    """Config 5 host logic: 5 frames over 2 ranks (the last step has one idle rank), one all-gather per step."""
    mp.spawn(_seq_worker, args=(2, _free_port(), 5), nprocs=2, join=True)
