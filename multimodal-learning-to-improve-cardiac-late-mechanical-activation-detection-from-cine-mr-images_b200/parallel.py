"""Multi-GPU layer: one process per GPU, slices sharded across ranks.

Every (slice, frame-pair) registration is independent
(/root/reference/modules/data/__init__.py:108-119 folds frames into the batch) and
the sector reduction is per slice, so the forward path shards by contiguous
blocks of SLICES with no data-path collective.  Training adds ONE all-reduce of
the parameter gradients per step (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_slices(n_slices: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous [start, stop) block of slices owned by ``rank`` (sizes differ by at most one)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n_slices, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_gradients(params, world_size: int | None = None, bucket_bytes: int = 32 << 20) -> int:
    """Average ``.grad`` of ``params`` over all ranks with bucketed flat all-reduces.

    The gradient volume of this model is tiny (a few hundred KB), so the collective is
    latency bound: buckets are sized for launch count, not link bandwidth.  Returns the
    number of collectives issued.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    ws = world_size or dist.get_world_size()
    if ws == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    n_coll, bucket, size = 0, [], 0

    def flush():
        nonlocal n_coll, bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(ws)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_coll += 1
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return n_coll


class GradientAllReducer:
    """Gradient all-reduce overlapped with the rest of the backward pass.

    ``buckets``: lists of parameters in the order their gradients become ready during ``backward()`` (for the joint
    training step: the LMA net first, the registration net last).  A post-accumulate-grad hook counts the ready
    gradients of each bucket; when a bucket is complete its gradients are flattened and ONE asynchronous
    all-reduce is launched at once - NCCL runs it on its own stream, so the LMA bucket travels while the EPDiff
    adjoint kernel is still running.  ``finish()`` (after ``backward()``) waits for the collectives, averages and
    scatters the results back into ``.grad``; it also launches any bucket whose hooks did not all fire (parameters
    without gradient).  Without an initialised process group (or world size 1) everything is a no-op.
    """

    def __init__(self, buckets):
        self.buckets = [[p for p in b if p.requires_grad] for b in buckets]
        self.buckets = [b for b in self.buckets if b]
        self._ready = [0] * len(self.buckets)
        self._pending = [None] * len(self.buckets)
        self._hooks = []
        for bi, b in enumerate(self.buckets):
            for p in b:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _make_hook(self, bi):
        def hook(_param):
            self._ready[bi] += 1
            if self._ready[bi] == len(self.buckets[bi]):
                self._launch(bi)
        return hook

    def _launch(self, bi):
        if self._pending[bi] is not None or not self._active():
            return
        grads = [p.grad for p in self.buckets[bi] if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
        self._pending[bi] = (work, flat, grads)

    def finish(self) -> int:
        """Wait for the launched buckets (launching the incomplete ones first); returns the number of collectives."""
        n = 0
        if self._active():
            ws = dist.get_world_size()
            for bi in range(len(self.buckets)):
                self._launch(bi)
            for bi, pend in enumerate(self._pending):
                if pend is None:
                    continue
                work, flat, grads = pend
                work.wait()
                flat.div_(ws)
                off = 0
                for g in grads:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
                n += 1
        self._ready = [0] * len(self.buckets)
        self._pending = [None] * len(self.buckets)
        return n

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def gather_strain_matrices(S_local: torch.Tensor, n_slices_total: int) -> torch.Tensor | None:
    """Gather the per-rank (B_r,1,K,F) strain matrices on rank 0 in slice order (inference only)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return S_local
    ws, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_slices(n_slices_total, r, ws) for r in range(ws)]
    maxb = max(b - a for a, b in sizes)
    pad = torch.zeros((maxb,) + tuple(S_local.shape[1:]), dtype=S_local.dtype, device=S_local.device)
    pad[: S_local.shape[0]] = S_local
    outs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(outs, pad)
    if rank != 0:
        return None
    return torch.cat([o[: b - a] for o, (a, b) in zip(outs, sizes)], dim=0)
