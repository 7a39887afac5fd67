"""2+ GPU check of ShardedSweep.run_to_root (gather fused into the epilogue, uneven shards, five families per sample
through the row-per-lane kernel) against ShardedSweep.run (NCCL all-gather)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
dist.init_process_group('cpu:gloo,cuda:nccl')
from pyhalomodel_b200 import workloads, sweep, synthetic as syn
G = 2*world+1                                    # uneven shards


def build(lo, hi):
    return workloads.build(hi-lo, 96, 256, ('m', 'g'), syn.FAMILY_NAMES, seed=77+lo, fiducial_first=False)


sw = sweep.ShardedSweep(G, build, models_per_sample=5)
fused = sw.run_to_root()
ref = sw.run()
if rank == 0:
    print('shape', tuple(fused.shape), 'equal to the NCCL gather:', bool(torch.equal(fused, ref)),
          'finite', bool(torch.isfinite(fused).all()))
dist.barrier()
dist.destroy_process_group()
