"""2+ GPU check of the fused gather (peer stores into a symmetric buffer on rank 0) against NCCL all_gather."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
from pyhalomodel_b200 import workloads, sweep
batch = workloads.build(4, 64, 256, ('m', 'g', 'p'), seed=100+rank, fiducial_first=False)
plan = batch.plan()
peer = sweep.PeerGatherBuffer(plan.out.shape, 3)
for i in range(3):
    plan.run(out=peer.slot(i), peer=True)
root = peer.finish()
mine = plan.run()[0].clone(); torch.cuda.synchronize()
ref_all = sweep.gather_spectra(mine, world*4)
if rank == 0:
    got = root[:, 2].reshape(ref_all.shape)
    d = (got-ref_all).abs()
    print('nan got', int(torch.isnan(got).sum()), 'nan ref', int(torch.isnan(ref_all).sum()), 'max abs diff', float(d[~torch.isnan(d)].max()) if d.numel() else 0.,
          'equal', bool(torch.equal(got, ref_all)), 'per-rank equal', [bool(torch.equal(root[r, 2], ref_all.reshape(world, *mine.shape)[r])) for r in range(world)])
dist.destroy_process_group()
