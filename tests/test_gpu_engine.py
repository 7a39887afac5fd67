"""Engine-level behaviour on the GPU that the benchmark relies on: software-pipelined steps over two streams give
the same bits as plain calls, the per-stage entry points compose to the full call, and the `then` hook runs on
the step's stream."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipelined_plans_equal_plain_runs():
    from pyhalomodel_b200 import engine, workloads
    batch = workloads.build(5, 70, 192, ('m', 'g', 'p'), seed=21)
    plan = batch.plan()
    want, lowmass = plan.run()
    want, lowmass = want.clone(), lowmass.clone()
    pipe = engine.PipelinedPlans(batch.plan)
    kept = []
    for _ in range(5):          # odd count: both plans and both streams are used, the last step on stream 0
        out, low = pipe.submit(then=lambda o, a: kept.append((o.clone(), a.clone())))
    pipe.drain()
    torch.cuda.synchronize()
    assert len(kept) == 5
    for o, a in kept:
        assert torch.equal(o, want) and torch.equal(a, lowmass)
    assert torch.equal(out, want)
    # an explicit output buffer is honoured
    mine = torch.empty_like(want)
    pipe.submit(out=mine)
    pipe.drain()
    torch.cuda.synchronize()
    assert torch.equal(mine, want)


def test_stage_entry_points_compose():
    from pyhalomodel_b200 import workloads
    batch = workloads.build(3, 41, 128, ('m', 'g'), seed=22)
    plan = batch.plan()
    want = plan.run()[0].clone()
    plan.out.zero_()
    plan.weights()
    plan.quadrature(1)          # streaming kernel only: partial sums
    plan.quadrature(2)          # epilogue only
    torch.cuda.synchronize()
    assert torch.equal(plan.out, want)
    assert plan.stream_kernel_bytes() == 8*(3*2*41*128+3*5*128) and plan.algorithmic_bytes() > plan.stream_kernel_bytes()
