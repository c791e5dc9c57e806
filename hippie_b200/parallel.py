"""Data-parallel plumbing (one process per GPU, torch.distributed): batch sharding and the gradient exchange.

The reference has no explicit parallelism; under Lightning's implicit DDP every rank runs the step on its own
shard, BatchNorm statistics stay per rank, gradients are summed and divided by the world size before clipping
(SURVEY.md section 8e).  Here the exchange is ONE all-reduce of the engine's flat fp32 gradient buffer; the
1/world factor is folded into the fused clip + AdamW kernel (`grad_scale`)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_batch_indices(perm: Sequence[int], step: int, rank: int, world: int, per_rank_batch: int) -> List[int]:
    """Global step `step` consumes world*per_rank_batch consecutive indices of the epoch permutation; rank r takes
    the r-th contiguous slice (ragged tail allowed)."""
    g0 = step * world * per_rank_batch
    lo = g0 + rank * per_rank_batch
    return list(perm[lo:min(lo + per_rank_batch, min(len(perm), g0 + world * per_rank_batch))])


def shard_units(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n independent units for communication-free embedding inference; concatenating
    the shards in rank order restores the row order of the CSV."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Inverse of `shard_units`: every rank passes the rows it computed for its contiguous shard of `n_total` units and
    gets back all rows in unit order (rank order = row order of the CSV the reference writes).  The only communication of
    the embedding path, after the compute; one all-gather of equally padded blocks."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in (shard_units(n_total, r, world) for r in range(world))]
    assert local.shape[0] == sizes[dist.get_rank(group)], "rows do not match this rank's shard"
    pad = local.new_zeros((max(sizes),) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)])


def init_from_env(backend: str = "nccl"):
    """torchrun plumbing for the command-line tools: one process per GPU (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from
    the environment).  Returns (rank, world); (0, 1) when not launched under torchrun."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, **nccl_group_options(backend))
    return dist.get_rank(), dist.get_world_size()


def nccl_group_options(backend: str = "nccl") -> dict:
    """Keyword arguments for `dist.init_process_group`: NCCL's streams at high priority, so that the few CTAs of an
    all-reduce get SM slots ahead of the weight-gradient GEMMs the exchange overlaps with -- opt-in
    (`HIPPIE_B200_NCCL_HIPRIO=1`): on two GPUs it changes nothing (3.10 ms per step either way, profiles/r02_exp43_n2.txt)."""
    import os
    if backend != "nccl" or os.environ.get("HIPPIE_B200_NCCL_HIPRIO", "0") != "1":
        return {}
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        return {"pg_options": opts}
    except Exception:  # pragma: no cover - a torch build without the NCCL backend
        return {}


def all_reduce_gradients(flat_grads: torch.Tensor, group=None) -> float:
    """Sums the flat gradient buffer over the ranks (NCCL over NVLink on GPUs, gloo in the CPU tests) and returns
    the factor the optimizer kernel must apply (1/world)."""
    if not dist.is_available() or not dist.is_initialized():
        return 1.0
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def broadcast_state(tensors: Sequence[torch.Tensor], src: int = 0, group=None):
    """Rank `src`'s parameters / BatchNorm buffers / AdamW state to everybody (DDP's initial broadcast)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.broadcast(t, src=src, group=group)


_exchange_streams = {}


def _exchange_stream(device) -> "torch.cuda.Stream":
    """One side stream per device on which the slice waits and the all-reduces of the overlapped exchange are issued."""
    key = torch.device(device).index
    if key not in _exchange_streams:
        _exchange_streams[key] = torch.cuda.Stream(device=device, priority=-1)
    return _exchange_streams[key]


def train_step_overlapped(engine, x1, x2, src, cls, eps, beta, w1=1.0, w2=1.0, scalars=None, group=None) -> float:
    """One data-parallel forward + backward with the gradient exchange overlapped with the backward pass, in the order
    the gradients become final: the latent-head + decoder half of the buffer (52 %) -- its all-reduce runs on NCCL's
    stream under the encoders' backward pass --, the deep half of each encoder (layer3, layer4, Linear: 45 %), under the
    backward pass of the wide shallow layers, and only the shallow remainder (3 %, < 2 MB) after the step with nothing to
    hide under.  The step itself stays ONE launch (one CUDA graph, `hippie_train_fwd_bwd_part` part 4): the engine
    publishes the two points as events (`hippie_slice_wait`) that an exchange stream waits on, so the backward chain is
    never cut (round 1 cut the step into three graphs at these points: every cut drained the weight-gradient streams).
    `HIPPIE_B200_DP_PARTS=1` selects that older three-part variant.  Returns the factor for the optimizer (1/world).
    Without an initialised process group this is a plain train_fwd_bwd."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        engine.train_fwd_bwd(x1, x2, src, cls, eps, beta, w1, w2, scalars=scalars)
        return 1.0
    import os
    if os.environ.get("HIPPIE_B200_DP_PARTS", "0") == "1" or not hasattr(engine, "slice_wait"):
        return _train_step_overlapped_parts(engine, x1, x2, src, cls, eps, beta, w1, w2, scalars, group)
    g = engine.flat_grads
    reduce = lambda lo, hi: dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True)
    main = torch.cuda.current_stream(g.device)
    xs = _exchange_stream(g.device)
    engine.train_fwd_bwd_part(4, x1, x2, src, cls, eps, beta, w1, w2, scalars=scalars)
    bounds = engine.grad_bounds
    with torch.cuda.stream(xs):
        engine.slice_wait(0, xs)
        work = [reduce(engine.grad_split, g.numel())]
        engine.slice_wait(1, xs)
        work += [reduce(deep, end) for _, deep, end in bounds]
    work += [reduce(begin, deep) for begin, deep, _ in bounds]  # on the step's own stream: final when it gets there
    for w in work:
        w.wait()  # the current stream waits for NCCL's
    main.wait_stream(xs)
    return 1.0 / dist.get_world_size(group)


def _train_step_overlapped_parts(engine, x1, x2, src, cls, eps, beta, w1=1.0, w2=1.0, scalars=None, group=None) -> float:
    """Round-1 variant of the overlapped exchange: the step cut into three stream-ordered parts (0, 2, 3)."""
    g = engine.flat_grads
    reduce = lambda lo, hi: dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True)
    engine.train_fwd_bwd_part(0, x1, x2, src, cls, eps, beta, w1, w2, scalars=scalars)
    work = [reduce(engine.grad_split, g.numel())]
    bounds = engine.grad_bounds
    engine.train_fwd_bwd_part(2, x1, x2, src, cls, eps, beta, w1, w2)
    work += [reduce(deep, end) for _, deep, end in bounds]
    engine.train_fwd_bwd_part(3, x1, x2, src, cls, eps, beta, w1, w2)
    work += [reduce(begin, deep) for begin, deep, _ in bounds]
    for w in work:
        w.wait()
    return 1.0 / dist.get_world_size(group)
