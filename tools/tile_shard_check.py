"""torchrun helper: a 128x128 relight frame tile-sharded over the ranks (one all-gather) equals the single-GPU frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from relightableavatar_b200 import parallel, scene
from relightableavatar_b200.renderer import Renderer

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
dev = torch.device(f'cuda:{local}')
torch.cuda.set_device(dev)
dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
sd = scene.make_state_dict(0, relight=True, fitted=True)
ok = True
# 128^2: one chunk.  512^2: P > 65536 rays = two reference chunks -- the shadow-ray box depends on the ray's chunk
# (in-place wbounds growth); ra_set_ray_layout makes the shards use global ray indices, so they agree bit for bit as well.
for H in (128, 512):
    b = scene.make_batch(H, H, seed=0, n_env=0)
    bt = {k: (torch.from_numpy(v).to(dev) if hasattr(v, 'shape') and getattr(v, 'ndim', 0) > 0 else v) for k, v in b.items() if k != 'novel_lights'}
    r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=dev, precision='tc', max_rays=b['ray_o'].shape[1] + 8, test_light=('main',), sync_timing=False)
    full = r.render(bt)['main']
    sharded = parallel.render_tile_sharded(lambda bb: r.render(bb)['main'], bt, keys=('rgb_map', 'acc_map'), engine=r.engine)
    ok = ok and torch.equal(sharded['rgb_map'], full['rgb_map']) and torch.equal(sharded['acc_map'], full['acc_map'])
    if H == 512:       # strong-scaling timing of the sharded frame (config 4 shape)
        for _ in range(2):
            parallel.render_tile_sharded(lambda bb: r.render(bb)['main'], bt, keys=('rgb_map', 'acc_map'), engine=r.engine)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            parallel.render_tile_sharded(lambda bb: r.render(bb)['main'], bt, keys=('rgb_map', 'acc_map'), engine=r.engine)
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            print(f'TILE_SHARD_TIMING world={world} 512x512 frame: {e0.elapsed_time(e1) / 5:.2f} ms')
    r.engine.close()
# ra_allgather (the C-ABI's own collective entry point) against torch.distributed on the same buffers: an NCCL communicator
# is created here through ctypes (unique id from rank 0, shared with a torch broadcast)
import ctypes
from relightableavatar_b200 import _lib
from relightableavatar_b200.renderer import Engine, default_config
nccl = ctypes.CDLL('libnccl.so.2')
uid = (ctypes.c_char * 128)()
if rank == 0:
    assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
t_uid = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=dev)
dist.broadcast(t_uid, 0)
uid = (ctypes.c_char * 128).from_buffer_copy(bytes(t_uid.cpu().tolist()))
comm = ctypes.c_void_p()


class _Uid(ctypes.Structure):
    _fields_ = [('internal', ctypes.c_char * 128)]


nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _Uid, ctypes.c_int]
u = _Uid(); ctypes.memmove(ctypes.byref(u), uid, 128)
assert nccl.ncclCommInitRank(ctypes.byref(comm), world, u, rank) == 0
eng = Engine(default_config(True, precision=1, max_rays=1024), dev)
send = torch.arange(4096, device=dev, dtype=torch.float32) + 10000 * rank
recv = torch.empty(world * 4096, device=dev)
lib = _lib.load()
rc = lib.ra_allgather(eng.h, comm, ctypes.c_void_p(send.data_ptr()), 4096, ctypes.c_void_p(recv.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
ref = torch.empty_like(recv)
dist.all_gather_into_tensor(ref, send)
ok_ag = rc == 0 and torch.equal(recv, ref)
if rank == 0:
    print('RA_ALLGATHER_OK' if ok_ag else f'RA_ALLGATHER_MISMATCH rc={rc}')
nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
nccl.ncclCommDestroy(comm)
eng.close()
ok = ok and ok_ag
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print('TILE_SHARD_OK' if flag.item() == 1 else 'TILE_SHARD_MISMATCH', float((sharded['rgb_map'] - full['rgb_map']).abs().max()))
dist.destroy_process_group()
