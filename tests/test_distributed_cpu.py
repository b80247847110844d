"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo processes."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imagined_speech_translation_b200 import distributed as dp
from imagined_speech_translation_b200 import trainer as tr


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = dp.init_from_env("gloo")
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3),
                                torch.nn.Linear(3, 3))      # last layer never used: grad stays None
    if rank == 1:                                            # replicas must start identical
        for p in model.parameters():
            p.data.add_(1.0)
    dp.broadcast_parameters(model, src=0)
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    idx = dp.shard_range(8, r, w)
    xs, ys = x[idx.start:idx.stop], y[idx.start:idx.stop]
    # sum-reduced loss so that averaging per-rank grads of the mean loss equals the full-batch grad
    loss = ((model[2](model[1](model[0](xs))) - ys) ** 2).mean()
    loss.backward()
    calls = dp.allreduce_gradients(list(model.parameters()), average=True, bucket_bytes=64)
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    buf = [flat.clone()]
    dp.allreduce_sum_(buf)                                   # flat-buffer path: SUM
    # numpy arrays travel by value; tensors would travel as shared-memory handles that die with this process
    q.put((rank, flat.numpy().copy(), calls, buf[0].numpy().copy(), [p.grad is None for p in model.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_matches_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    res = [(r, torch.from_numpy(f), c, torch.from_numpy(b), n) for r, f, c, b, n in res]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # single-process reference on the full batch
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 3))
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    ((model[2](model[1](model[0](x))) - y) ** 2).mean().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    for rank, flat, calls, summed, none_mask in res:
        torch.testing.assert_close(flat, ref, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(summed, 2 * ref, rtol=1e-5, atol=1e-6)
        assert calls >= 2                                    # tiny bucket size forces several buckets
        assert none_mask == [False, False, False, False, True, True]


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 256, 1000):
        for w in (1, 2, 4, 8):
            parts = [dp.shard_range(n, r, w) for r in range(w)]
            covered = [i for p in parts for i in p]
            assert covered == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= math.ceil(n / w)


def test_cosine_schedule_matches_transformers():
    from transformers import get_cosine_schedule_with_warmup
    def mk():
        p = torch.nn.Parameter(torch.zeros(1))
        return torch.optim.SGD([p], lr=3e-4)
    a, b = mk(), mk()
    sa = tr.cosine_schedule_with_warmup(a, 5, 40)
    sb = get_cosine_schedule_with_warmup(b, num_warmup_steps=5, num_training_steps=40)
    assert a.param_groups[0]['lr'] == 0.0            # lambda(0) = 0: first optimizer step is a no-op
    for _ in range(45):
        assert a.param_groups[0]['lr'] == pytest.approx(b.param_groups[0]['lr'], rel=1e-12, abs=1e-18)
        a.step(); b.step(); sa.step(); sb.step()


def test_optimizer_groups_route_by_name():
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.brain_encoder = torch.nn.Linear(2, 2)
            self.bart_decoder = torch.nn.ModuleDict({'eeg_to_bart': torch.nn.Linear(2, 2), 'bart': torch.nn.Linear(2, 2)})
    groups = tr.get_optimizer_groups(M())
    assert [len(g['params']) for g in groups] == [2, 2, 2]
    assert [g['lr'] for g in groups] == [3e-4, 1e-4, 3e-5]


def test_grad_boundary_fires_after_everything_downstream_has_its_gradient():
    """The overlap of the all-reduce with backward rests on one property of `fused.grad_boundary`: when its backward
    runs, every parameter used DOWNSTREAM of it (created later in forward) already holds its final gradient, while
    parameters upstream do not yet.  Checked on CPU with plain modules, including a branch that re-joins (the
    cross-scale pattern: a later layer also consumes an earlier activation directly)."""
    from imagined_speech_translation_b200 import fused
    torch.manual_seed(0)
    up = torch.nn.Linear(8, 8)
    mid = torch.nn.Linear(8, 8)
    down1 = torch.nn.Linear(8, 8)
    down2 = torch.nn.Linear(16, 4)
    seen = {}

    def cb(key):
        seen[key] = {n: (m.weight.grad is not None) for n, m in
                     dict(up=up, mid=mid, down1=down1, down2=down2).items()}

    x = torch.randn(5, 8)
    fused.set_grad_boundary_callback(cb)
    try:
        a = up(x)
        early = a                                           # consumed again after the boundary, bypassing it
        b = fused.grad_boundary(mid(torch.tanh(a)), 'B')
        c = down1(torch.tanh(b))
        out = down2(torch.cat([c, early], dim=1))
        out.square().mean().backward()
    finally:
        fused.set_grad_boundary_callback(None)
    assert seen['B'] == {'up': False, 'mid': False, 'down1': True, 'down2': True}
    # without a callback the boundary is a no-op (same tensor object, nothing on the tape)
    t = torch.randn(3, requires_grad=True)
    assert fused.grad_boundary(t, 'x') is t


def test_grad_runs_merge_adjacent_slices():
    """FlatAdamW.grad_runs: contiguous runs of the flat gradient buffer for a parameter subset (offsets stubbed)."""
    from imagined_speech_translation_b200.optim import FlatAdamW
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (8, 16, 8, 24)]
    opt = FlatAdamW.__new__(FlatAdamW)
    opt._flat = [object()]
    opt._offsets = {id(ps[0]): (0, 8), id(ps[1]): (8, 16), id(ps[2]): (24, 8), id(ps[3]): (32, 24)}
    assert opt.grad_runs(ps) == [(0, 56)]
    assert opt.grad_runs([ps[0], ps[2], ps[3]]) == [(0, 8), (24, 56)]
    assert opt.grad_runs([ps[3], ps[1]]) == [(8, 24), (32, 56)]
    assert opt.grad_runs([torch.nn.Parameter(torch.zeros(3))]) == []      # no gradient slot: skipped


def test_registered_gradient_buffer_is_optional():
    """Outside a multi-rank NCCL job (single process, gloo, or EEGX_NCCL_REGISTER=0) the NCCL-registered allocation
    is simply not used: the helper returns None and the optimizer falls back to plain device memory."""
    from imagined_speech_translation_b200 import distributed as dp
    assert dp.nccl_registered_zeros(16, "cpu") is None
    dp.release_registered_buffers()          # nothing registered: a no-op
