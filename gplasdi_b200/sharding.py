"""Multi-GPU sharding of the test-parameter grid for the greedy step.

One process per GPU.  The P test parameters are split into contiguous blocks, rank r owns
[bounds[r], bounds[r+1]); rollouts, decode and variance are independent per parameter, so the
only exchange is ONE max-loc all-reduce of a packed 64-bit key per rank:

    key = (float_bits(max_std) << 32) | (0xFFFFFFFF - global_index)

Non-negative floats order like their bit patterns, so an integer MAX picks the largest value and,
on ties, the SMALLEST global index -- the reference's strict-`>` first-wins rule
(reference src/lasdi/gplasdi.py:62-72).  A rank with no positive std (or an empty block: more ranks than
test parameters) contributes key 0.  A rank whose step FAILED contributes the largest int64 instead, so
the same collective tells every rank and nobody is left waiting in it.
Works with NCCL (CUDA tensors, NVLink) and with gloo (CPU tensors; used by the CPU tests).
"""
from __future__ import annotations

import struct

import torch
import torch.distributed as dist

ERROR_KEY = 0x7FFFFFFFFFFFFFFF        # > every packed key (float bits of a finite positive fp32 < 0x7F800000)


def shard_bounds(P: int, world: int):
    """Balanced contiguous blocks: the first P % world ranks get one extra parameter."""
    base, rem = divmod(P, world)
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


def pack_maxloc_host(val: float, global_idx: int) -> int:
    """Host mirror of glasdi_pack_maxloc (csrc/reduce.cu) for CPU-side logic and tests."""
    if global_idx < 0 or not (val > 0.0):
        return 0
    bits = struct.unpack("<I", struct.pack("<f", float(val)))[0]
    return (bits << 32) | (0xFFFFFFFF - (global_idx & 0xFFFFFFFF))


def unpack_maxloc(key: int):
    """-> (value, global_index) or (0.0, -1) for the empty key."""
    key = int(key)
    if key == 0:
        return 0.0, -1
    if key == ERROR_KEY:
        raise RuntimeError("max-loc key carries the error sentinel: the greedy step failed on at least one rank")
    bits = (key >> 32) & 0xFFFFFFFF
    idx = 0xFFFFFFFF - (key & 0xFFFFFFFF)
    return struct.unpack("<f", struct.pack("<I", bits))[0], int(idx)


def allreduce_maxloc(key: torch.Tensor, group=None) -> torch.Tensor:
    """In-place MAX all-reduce of the int64 key tensor (one element per rank)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)
    return key


def reduce_keys(local_key, error, group=None, device=None):
    """The collective of `sharded_greedy_step` on its own (also what the gloo CPU tests drive): `local_key` is this
    rank's packed key (int64 tensor [1] or int), `error` the exception its step raised (or None).  Every rank
    returns the reduced key tensor, or every rank raises."""
    if not torch.is_tensor(local_key):
        local_key = torch.tensor([int(local_key)], dtype=torch.int64, device=device)
    if error is not None:
        local_key.fill_(ERROR_KEY)
    allreduce_maxloc(local_key, group)
    if error is not None:
        raise error
    if int(local_key.item()) == ERROR_KEY:
        raise RuntimeError("sharded greedy step: another rank failed (its own exception names the cause)")
    return local_key


def sharded_greedy_step(engine, coefs_local, z0_local, T: int, dt: float, index_offset: int, group=None):
    """Runs the fused greedy step on this rank's block and reduces over ranks.
    Returns (max_std_local CUDA fp32 [P_local], key CUDA int64 [1] after the all-reduce)."""
    # numeric flags are checked (and a flagged screened step repeated in direct mode) BEFORE the exchange, so that
    # every rank contributes a trusted key
    err = None
    if int(coefs_local.shape[0]) == 0:                                   # more ranks than test parameters
        ms = torch.empty(0, dtype=torch.float32, device=engine.tdev)
        key = torch.zeros(1, dtype=torch.int64, device=engine.tdev)
    else:
        try:
            ms, am, mv = engine.checked(lambda: engine.greedy_step(coefs_local, z0_local, T, dt))
            key = engine.pack_maxloc(mv, am, index_offset)
        except Exception as e:                                           # noqa: BLE001 -- re-raised after the collective
            err = e
            ms = None
            key = torch.zeros(1, dtype=torch.int64, device=engine.tdev)
    key = reduce_keys(key, err, group)
    return ms, key
