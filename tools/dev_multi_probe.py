#!/usr/bin/env python
"""Where does the streaming kernel lose time in multi-rank runs?  Times plan.quadrature(1) at several points of the
multi-GPU set-up.  MODE (argv[1]): nccl (eager NCCL init), gloo (gloo only + symmetric buffer), lazy (NCCL group
without eager communicator; barrier on a gloo subgroup)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
from pyhalomodel_b200 import workloads, sweep

mode = sys.argv[1] if len(sys.argv) > 1 else 'nccl'
B = 64
batch = workloads.build(B, 512, 2048, ('m', 'g', 'p'), seed=1234+rank)
plan = batch.plan()
plan.run()


def t(label):
    for _ in range(3):
        plan.quadrature(1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        plan.quadrature(1)
    b.record()
    torch.cuda.synchronize()
    if rank == 0:
        print('[%s] %-44s %.2f us' % (mode, label, 1e3*a.elapsed_time(b)/20), flush=True)


t('before init_process_group')
os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
if mode == 'nccl':
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dist.barrier()
    torch.cuda.synchronize()
    t('after eager NCCL init + NCCL barrier')
elif mode == 'gloo':
    dist.init_process_group('gloo')
    dist.barrier()
    t('after gloo init + barrier')
elif mode == 'lazy':
    dist.init_process_group('cpu:gloo,cuda:nccl')
    dist.barrier()            # no tensor: which backend?  time it anyway
    torch.cuda.synchronize()
    t('after cpu:gloo,cuda:nccl init + barrier()')
    x = torch.zeros(1)
    dist.all_reduce(x)
    t('after gloo all_reduce on a CPU tensor')
try:
    peer = sweep.PeerGatherBuffer(plan.out.shape, 4)
    torch.cuda.synchronize()
    dist.barrier()
    t('after symmetric buffer')
    out, _ = plan.run(out=peer.slot(0), peer=True)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        mine = plan.run()[0]
        print('[%s] peer slot of rank 0 equals local result: %s' % (mode, bool(torch.equal(peer.local[0, 0], mine))), flush=True)
except Exception as e:
    if rank == 0:
        print('[%s] symmetric buffer failed: %r' % (mode, e), flush=True)
if mode != 'nccl':
    try:
        y = torch.ones(1, device='cuda')
        dist.all_reduce(y)
        torch.cuda.synchronize()
        t('after first NCCL collective on a CUDA tensor')
    except Exception as e:
        if rank == 0:
            print('[%s] cuda collective failed: %r' % (mode, e), flush=True)
dist.barrier()
dist.destroy_process_group()
