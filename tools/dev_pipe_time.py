"""Developer timing: consecutive steps software-pipelined over two CUDA streams (two plans, double-buffered)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyhalomodel_b200 import workloads
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
batch = workloads.build(B, 512, 2048)
plans = [batch.plan(), batch.plan()]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
main = torch.cuda.current_stream()
def run(K):
    for s in streams: s.wait_stream(main)
    for i in range(K):
        with torch.cuda.stream(streams[i & 1]):
            plans[i & 1].run()
    for s in streams: main.wait_stream(s)
run(6); torch.cuda.synchronize()
for K in (20, 40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(K); e1.record(); torch.cuda.synchronize()
    print('two streams: K=%d  %.2f us/step' % (K, e0.elapsed_time(e1)/K*1e3))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): plans[0].run()
e1.record(); torch.cuda.synchronize()
print('one stream : %.2f us/step' % (e0.elapsed_time(e1)/20*1e3))
