"""Host-side multi-GPU logic: partition of the test grid and the packed max-loc all-reduce,
covered on CPU with a world_size-2 gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gplasdi_b200 import sharding


def test_shard_bounds_cover_and_balance():
    for P, W in [(441, 8), (4096, 8), (5, 8), (441, 1), (7, 2)]:
        b = sharding.shard_bounds(P, W)
        assert b[0] == 0 and b[-1] == P and len(b) == W + 1
        sizes = np.diff(b)
        assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1


def test_pack_orders_by_value_then_smallest_index():
    k = sharding.pack_maxloc_host
    assert k(0.5, 10) > k(0.4, 0)
    assert k(0.5, 3) > k(0.5, 4)            # tie -> smaller global index wins
    assert k(0.0, 3) == 0 and k(1.0, -1) == 0 and k(float("nan"), 2) == 0
    v, i = sharding.unpack_maxloc(k(0.7, 237))
    assert i == 237 and v == np.float32(0.7)
    assert sharding.unpack_maxloc(0) == (0.0, -1)
    assert k(3.4e38, 2 ** 31) < 2 ** 63     # always a non-negative int64


def _first_strict_max(vals):
    best, idx = 0.0, -1
    for m, v in enumerate(vals):
        if v > best:
            best, idx = v, m
    return idx


def _worker(rank, world, port, vals, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = sharding.shard_bounds(len(vals), world)
    local = vals[b[rank]:b[rank + 1]]
    li = _first_strict_max(local)
    key = torch.tensor([sharding.pack_maxloc_host(local[li] if li >= 0 else 0.0, b[rank] + li if li >= 0 else -1)],
                       dtype=torch.int64)
    sharding.allreduce_maxloc(key)
    q.put((rank, int(key.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["unique", "tie_across_ranks", "all_zero"])
def test_gloo_world2_selects_reference_index(case):
    rs = np.random.RandomState(0)
    vals = rs.uniform(0.1, 1.0, size=11).astype(np.float32)
    if case == "tie_across_ranks":
        vals[2] = vals[9] = 2.0              # same maximum on both ranks -> index 2 must win
    if case == "all_zero":
        vals[:] = 0.0
    expect = _first_strict_max(list(vals))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, list(map(float, vals)), q)) for r in range(2)]
    for p in procs:
        p.start()
    keys = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert keys[0][1] == keys[1][1]
    _, idx = sharding.unpack_maxloc(keys[0][1])
    assert idx == expect


def _worker_fail(rank, world, port, fail_rank, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    err = ValueError("numeric flag on this rank") if rank == fail_rank else None
    try:
        key = sharding.reduce_keys(sharding.pack_maxloc_host(0.25 + rank, 3 + rank), err)
        q.put((rank, "ok", int(key.item())))
    except Exception as e:                                     # noqa: BLE001
        q.put((rank, type(e).__name__, str(e)))
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 0, 1])
def test_gloo_world2_failed_rank_reaches_every_rank(fail_rank):
    """A rank whose step raised must not leave the others waiting in the all-reduce: the error rides in the same
    collective (sentinel key) and EVERY rank raises afterwards -- the failed one its own exception."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_fail, args=(r, 2, port, fail_rank, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (kind, msg)) for r, kind, msg in [q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if fail_rank < 0:
        assert res[0] == res[1] and res[0][0] == "ok"
        assert sharding.unpack_maxloc(res[0][1]) == (1.25, 4)
    else:
        assert res[fail_rank][0] == "ValueError"
        assert res[1 - fail_rank][0] == "RuntimeError" and "another rank failed" in res[1 - fail_rank][1]


def test_more_ranks_than_parameters_gives_empty_blocks():
    b = sharding.shard_bounds(3, 8)
    assert b == [0, 1, 2, 3, 3, 3, 3, 3, 3]
    with pytest.raises(RuntimeError):
        sharding.unpack_maxloc(sharding.ERROR_KEY)
    assert sharding.ERROR_KEY > sharding.pack_maxloc_host(3.4e38, 0)
