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
b = scene.make_batch(128, 128, seed=0, n_env=0)
bt = {k: (torch.from_numpy(v).to(dev) if hasattr(v, 'shape') and getattr(v, 'ndim', 0) > 0 else v) for k, v in b.items() if k != 'novel_lights'}
sd = scene.make_state_dict(0, relight=True, fitted=True)
r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=dev, precision='tc', max_rays=8192, test_light=('main',), sync_timing=False)
full = r.render(bt)['main']
sharded = parallel.render_tile_sharded(lambda bb: r.render(bb)['main'], bt, keys=('rgb_map', 'acc_map'))
ok = torch.equal(sharded['rgb_map'], full['rgb_map']) and torch.equal(sharded['acc_map'], full['acc_map'])
# shadow-ray far distance depends on the reference's per-chunk box growth, which is the constant 0.25 for P <= 65536 rays, so shards agree bit for bit
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print('TILE_SHARD_OK' if flag.item() == 1 else 'TILE_SHARD_MISMATCH', float((sharded['rgb_map'] - full['rgb_map']).abs().max()))
dist.destroy_process_group()
