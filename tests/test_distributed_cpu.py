"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, the factor broadcast helper and the
arg-best exchange.  The data path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from safe_exploration_b200 import distributed as sd


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [sd.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sd.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = sd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # "factor" buffers: rank 0 holds the truth, the others receive it with one broadcast per buffer
    rng = np.random.RandomState(0)
    truth = [rng.randn(64, 64), rng.randn(128), rng.randn(2)]
    bufs = [torch.from_numpy(t.copy()) if rank == 0 else torch.zeros(t.shape, dtype=torch.float64) for t in truth]
    sd.broadcast_buffers(bufs, src=0)
    ok = all(np.array_equal(b.numpy(), t) for b, t in zip(bufs, truth))
    # sharded candidates: every rank scores its shard, the best of all ranks is agreed on
    total = 1001
    cost = np.random.RandomState(1).rand(total)
    cost[3] = np.nan                                        # a diverged candidate must never win
    s0, s1 = sd.shard_range(total, rank, world)
    local = cost[s0:s1]
    j = int(np.nanargmin(local))
    best_cost, best_idx, best_rank = sd.argmin_across_ranks(local[j], s0 + j)
    ok = ok and best_idx == int(np.nanargmin(cost)) and abs(best_cost - np.nanmin(cost)) == 0.0
    ok = ok and sd.max_across_ranks(float(rank + 1)) == float(world)
    dist.barrier()
    with open(os.path.join(out_dir, "ok%d" % rank), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_broadcast_and_argmin_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(os.path.join(str(tmp_path), "ok%d" % r)).read() == "1"


def _best_worker(rank, world, port, out_dir):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from safe_exploration_b200.safempc_sampling import _best_across_ranks
    # rank 0: infeasible but tiny violation; rank 1: feasible with a larger cost -> feasible wins
    local = (3, -1.0, 0.2, False) if rank == 0 else (8192 + 5, 7.0, -0.1, True)
    res1 = _best_across_ranks(local)
    # both feasible: lowest cost wins; equal cost: lowest index
    res2 = _best_across_ranks((10 + rank, 2.0, -0.5, True))
    # nobody has a candidate
    res3 = _best_across_ranks((-1, float("inf"), float("inf"), False))
    with open(os.path.join(out_dir, "best_%d.txt" % rank), "w") as f:
        f.write(repr((res1, res2, res3)))
    dist.destroy_process_group()


def test_best_candidate_across_ranks_world_size_2(tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 200
    mp.spawn(_best_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    outs = [eval(open(os.path.join(str(tmp_path), "best_%d.txt" % r)).read(), {"inf": float("inf")}) for r in range(2)]
    assert outs[0] == outs[1]
    res1, res2, res3 = outs[0]
    assert res1 == (8197, 7.0, -0.1, True)
    assert res2 == (10, 2.0, -0.5, True)
    assert res3[0] == -1 and res3[3] is False


class _FakeModel(object):
    """Stand-in for BatchedGPSSM on CPU tensors: the factor arena, optionally a float64 operand, and the
    mark_factorized contract of the library (refuses while a needed second buffer is missing)."""

    def __init__(self, rank, need_fp64):
        self.device = torch.device("cpu")
        self.rank = rank
        rng = np.random.RandomState(5)
        self.truth = [rng.randint(0, 255, 4096).astype(np.uint8), rng.randint(0, 255, 1024).astype(np.uint8)]
        self.arena = torch.from_numpy(self.truth[0].copy()) if rank == 0 else torch.zeros(4096, dtype=torch.uint8)
        self.need = need_fp64                # on non-root ranks this is "known" only after the arena has arrived
        self.wt = torch.from_numpy(self.truth[1].copy()) if (rank == 0 and need_fp64) else None
        self.marked = 0

    def factor_views(self):
        return [self.arena] + ([self.wt] if self.wt is not None else [])

    def alloc_fp64_operand(self):
        if self.wt is None:
            self.wt = torch.zeros(1024, dtype=torch.uint8)

    def get_option(self, name):
        assert name == "fp64_operand_needed"
        return 1 if self.need else 0

    def mark_factorized(self):
        assert np.array_equal(self.arena.numpy(), self.truth[0])        # only ever called after buffer 0 arrived
        if self.need and (self.wt is None or not np.array_equal(self.wt.numpy(), self.truth[1])):
            raise ValueError("float64 operand missing")
        self.marked += 1


def _factor_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sd.init_from_env(backend="gloo")
    ok = True
    for need in (False, True):
        gp = _FakeModel(rank, need)
        nbytes = sd.broadcast_factor(gp, src=0)
        ok = ok and nbytes == (4096 + 1024 if need else 4096)          # ONE buffer unless float64 is in play
        ok = ok and np.array_equal(gp.arena.numpy(), gp.truth[0])
        ok = ok and (gp.marked == (1 if rank != 0 else 0))
        if need:
            ok = ok and np.array_equal(gp.wt.numpy(), gp.truth[1])
    dist.barrier()
    with open(os.path.join(out_dir, "factor_ok%d" % rank), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_factor_broadcast_is_one_collective_unless_float64_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_factor_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(os.path.join(str(tmp_path), "factor_ok%d" % r)).read() == "1"
