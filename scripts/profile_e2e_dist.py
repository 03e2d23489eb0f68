"""cProfile of the end-to-end front-end call under torchrun (rank 0 prints).
torchrun --nproc-per-node 2 scripts/profile_e2e_dist.py"""
import cProfile, pstats, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import pypyls_b200 as pyls

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
lr = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
rs = np.random.RandomState(1234)
X, Y = rs.rand(80, 10000), rs.rand(80, 10)
Xh, Yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(Y).pin_memory()
kw = dict(groups=[20, 20], n_cond=2, n_perm=5000 * world, n_boot=5000 * world, verbose=False,
          device=lr)
for i in range(2):
    pyls.behavioral_pls(Xh, Yh, seed=i, **kw)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
pyls.behavioral_pls(Xh, Yh, seed=7, **kw)
torch.cuda.synchronize()
pr.disable()
if rank == 0:
    print('one call: %.1f ms' % (1e3 * (time.perf_counter() - t0)))
    pstats.Stats(pr).sort_stats('tottime').print_stats(18)
if world > 1:
    dist.destroy_process_group()
