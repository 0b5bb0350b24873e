"""Multi-GPU plumbing: one process per GPU, candidates sharded over ranks, ONE broadcast of the factor.

The path shards over the candidate axis B: every candidate is an independent H-step
recursion and the ranks share only the read-only model.  So

* ``shard_range`` splits B into contiguous per-rank shards;
* rank ``src`` factorises the model (K0), every other rank uploads the same data, allocates the factor
  buffers and receives them with one ``torch.distributed.broadcast`` per buffer (NCCL over NVLink on GPUs);
* there is NO collective inside the rollout loop;
* optionally, ``argmin_across_ranks`` picks the best candidate of all ranks (one tiny all-gather), which
  is the only exchange a sampling-MPC driver needs.

``torch.distributed`` is used purely as the launcher / communicator; with the ``gloo`` backend the same
helpers run on CPU tensors (tests/test_distributed_cpu.py, world_size 2).
"""
import os

import numpy as np

__all__ = ["shard_range", "init_from_env", "broadcast_buffers", "broadcast_factor", "build_replicated_model",
           "argmin_across_ranks", "max_across_ranks"]


def shard_range(total, rank, world_size):
    """Contiguous shard [start, stop) of `total` items for `rank`; the first total % world ranks get one extra."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(total), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).
    Returns (rank, world_size, local_rank).  A single process (no env) returns (0, 1, 0) without a group."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return 0, 1, local_rank
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        if backend == "nccl":
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def broadcast_buffers(tensors, src=0, group=None):
    """One broadcast per tensor from `src` (in place)."""
    import torch.distributed as dist
    for t in tensors:
        dist.broadcast(t, src=src, group=group)


def broadcast_factor(gp, src=0, group=None):
    """Broadcast the factorised state of a BatchedGPSSM from `src`.  Rank `src` must have trained the model; the
    others must have called ``set_data_only``.

    ONE broadcast: buffer 0 is the whole factorised state of the int8 path (a single device allocation: beta,
    log-determinants, digit planes, row factors, error-model weights, the probe's decisions).  Only a model that runs
    the float64 contraction has a second buffer (the DMMA operand); whether it does is known to every rank once
    buffer 0 has arrived (``mark_factorized`` refuses a model that still misses it), so the ranks agree on it with a
    one-integer MAX all-reduce and, only then, broadcast buffer 1.

    `gp` needs ``factor_views()`` (uint8 tensors aliasing the buffers), ``alloc_fp64_operand()``, ``mark_factorized()``
    and ``get_option("fp64_operand_needed")`` -- BatchedGPSSM, or a stand-in in the CPU tests."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)

    def sync():
        if gp.device.type == "cuda":
            torch.cuda.synchronize(gp.device)

    views = gp.factor_views()
    sync()
    broadcast_buffers(views[:1], src, group)
    sync()
    if rank != src:
        try:
            gp.mark_factorized()
            need = False
        except ValueError:
            need = True
    else:
        need = len(views) > 1 and bool(gp.get_option("fp64_operand_needed"))
    flag = torch.tensor([1 if need else 0], dtype=torch.int32, device=gp.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    sent = [views[0]]
    if int(flag.item()):
        if rank != src:
            gp.alloc_fp64_operand()
        views = gp.factor_views()
        if len(views) < 2:
            raise RuntimeError("the factorising rank did not keep the float64 operand (set tri_mode / keep_fp64 "
                               "before training)")
        broadcast_buffers(views[1:2], src, group)
        sync()
        sent.append(views[1])
        if rank != src:
            gp.mark_factorized()
    return sum(v.numel() for v in sent)


def build_replicated_model(n_s_out, n_s_in, n_u, x, y, kern_types, hyp, rank, world_size, src=0, device=None,
                           redundant=False):
    """Model on every rank: factorise on `src` and broadcast (default), or factorise redundantly on every rank
    (zero communication; SURVEY.md section 8e 'Alternative')."""
    from .ssm import BatchedGPSSM
    if world_size == 1 or redundant or rank == src:
        gp = BatchedGPSSM(n_s_out, n_s_in, n_u, x, y, kern_types=kern_types, hyp=hyp, device=device)
    else:
        gp = BatchedGPSSM(n_s_out, n_s_in, n_u, kern_types=kern_types, hyp=hyp, device=device)
        gp.set_data_only(x, y)
    if world_size > 1 and not redundant:
        broadcast_factor(gp, src)
    return gp


def argmin_across_ranks(local_cost, local_index, group=None):
    """(cost, global index, rank) of the smallest cost over all ranks; `local_index` is already global.
    Works with any backend (tiny all-gather of two float64 per rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(local_cost), int(local_index), 0
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([float(local_cost), float(local_index)], dtype=torch.float64, device=dev)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    costs = np.array([float(v[0]) for v in allv])
    costs = np.where(np.isnan(costs), np.inf, costs)
    r = int(np.argmin(costs))
    return float(costs[r]), int(allv[r][1]), r


def max_across_ranks(value, group=None):
    """MAX all-reduce of a scalar (device timings are reported as the max over ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
