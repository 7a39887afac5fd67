import os, torch, torch.distributed as dist
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty((world, 1024), dtype=torch.float64, device='cuda')
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    t.zero_()
    dist.barrier(); torch.cuda.synchronize()
    dst = hdl.get_buffer(0, (world, 1024), torch.float64)     # rank 0's buffer, mapped here
    dst[rank].copy_(torch.full((1024,), float(rank+1), dtype=torch.float64, device='cuda'))
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        print('symm ok:', t[:, 0].tolist(), 'multicast' , getattr(hdl, 'has_multicast_support', None))
except Exception as e:
    print('rank', rank, 'symm failed:', type(e).__name__, str(e)[:300])
dist.destroy_process_group()
