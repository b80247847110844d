"""Data-parallel plumbing: one process per GPU, trials shard by batch, one gradient all-reduce
per optimizer step.  The reference has no distributed code; this is the
single collective the hot path needs.  Preprocessing needs none (every stage is per trial)."""
from __future__ import annotations

import os
from typing import Iterable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_from_env(backend: Optional[str] = None):
    """torchrun-style initialisation (MASTER_ADDR / MASTER_PORT / RANK / WORLD_SIZE from the env)."""
    rank, world, local_rank = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local_rank


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous shard of a global batch: rank r owns [r*ceil(n/w), min(n, (r+1)*ceil(n/w)))."""
    per = (n_items + world - 1) // world
    return range(min(n_items, rank * per), min(n_items, (rank + 1) * per))


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Replicas start identical: parameters AND buffers (BatchNorm running statistics) from src."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def allreduce_sum_(buffers: Iterable[torch.Tensor], group=None) -> None:
    """In-place SUM all-reduce of a few large flat buffers (FlatAdamW.flat_grads())."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for b in buffers:
        dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)


def allreduce_gradients(params: Sequence[torch.nn.Parameter], group=None, average: bool = True,
                        bucket_bytes: int = 256 << 20) -> int:
    """Bucketed gradient all-reduce for optimizers without flat buffers.  Parameters whose grad
    is None (e.g. the BART encoder, which never runs) are skipped on every rank alike.
    Returns the number of collectives issued."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    calls, bucket, size = 0, [], 0

    def flush():
        nonlocal calls, bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        calls += 1
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return calls


_registered_pools: list = []      # (backend, pool): NCCL-registered allocations that must be deregistered before teardown


def nccl_registered_zeros(n: int, device) -> Optional[torch.Tensor]:
    """A zero-filled fp32 buffer of n elements allocated through NCCL's own allocator (``ncclMemAlloc``) and
    registered with the communicator, so that all-reduces on it (and on slices of it) run zero-copy: with NVLink
    SHARP (NVLS) the reduction happens in the switch on the user buffer itself, with fewer NCCL CTAs taking SMs from
    the backward kernels they overlap.  Returns None -- the caller allocates normally -- outside a multi-rank NCCL
    job, when ``EEGX_NCCL_REGISTER=0``, or when this torch / NCCL build lacks the hooks."""
    if os.environ.get("EEGX_NCCL_REGISTER", "1") == "0":
        return None
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return None
    if dist.get_backend() != "nccl" or not hasattr(torch.cuda, "MemPool"):
        return None
    try:
        backend = dist.distributed_c10d._get_default_group()._get_backend(torch.device(device))
        pool = torch.cuda.MemPool(backend.mem_allocator)
        with torch.cuda.use_mem_pool(pool):
            t = torch.zeros(n, device=device, dtype=torch.float32)
        backend.register_mem_pool(pool)
    except Exception as exc:       # older NCCL / no allocator support: plain memory works, only slower
        import warnings
        warnings.warn(f"NCCL buffer registration unavailable ({exc!r}); using an unregistered gradient buffer")
        return None
    _registered_pools.append((backend, pool))
    return t


def release_registered_buffers() -> None:
    """Deregister the NCCL-allocated pools (before the process group is destroyed)."""
    while _registered_pools:
        backend, pool = _registered_pools.pop()
        try:
            backend.deregister_mem_pool(pool)
        except Exception:
            pass


SHUTDOWN_STALLED_EXIT = 75      # EX_TEMPFAIL: teardown (graph release / barrier / destroy) did not finish in time


def shutdown(trainer=None, grace_s: float = 30.0) -> None:
    """Orderly end of a data-parallel job: release captured graphs (they may hold NCCL kernels), barrier,
    destroy the process group.  If the teardown itself stalls for ``grace_s`` seconds -- a hung NCCL
    communicator, or a peer rank that died and left the others at the barrier -- a watchdog says so on
    stderr and ends the process with the NON-ZERO status ``SHUTDOWN_STALLED_EXIT``, so the launcher sees a
    failure instead of waiting forever or recording a success."""
    import os
    import sys
    import threading
    sys.stdout.flush()
    sys.stderr.flush()

    def stalled():
        try:
            sys.stderr.write(f"eegx.distributed.shutdown: teardown stalled for {grace_s:.0f} s "
                             f"(rank {os.environ.get('RANK', '0')}); exiting with status {SHUTDOWN_STALLED_EXIT}\n")
            sys.stderr.flush()
        finally:
            os._exit(SHUTDOWN_STALLED_EXIT)

    timer = threading.Timer(grace_s, stalled)
    timer.daemon = True
    timer.start()
    if trainer is not None:
        trainer.release_graph()
    if dist.is_available() and dist.is_initialized():
        try:
            torch.cuda.synchronize() if torch.cuda.is_available() else None
            dist.barrier()
            release_registered_buffers()
        finally:
            dist.destroy_process_group()
    timer.cancel()
