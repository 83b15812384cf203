"""Two-rank NCCL check of the sharded greedy step (run under torchrun by tests/test_gpu_named.py, or by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/nccl_worker.py).
Every rank asserts; rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from gplasdi_b200 import synth, sharding, _lib
from gplasdi_b200.engine import Engine


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    eng = Engine(local)
    wl = synth.WORKLOADS["small_s200"]
    Ws, bs = synth.make_decoder(wl.widths, wl.seed)
    eng.set_decoder(Ws, bs, wl.act)
    inp = synth.make_inputs(wl)
    # 13 parameters (6 + 6 + 1 of the fixture's) -> uneven blocks for any world > 1
    sel = list(range(6)) + list(range(6)) + [3]
    coefs = np.ascontiguousarray(inp["coefs"][sel]); z0 = np.ascontiguousarray(inp["z0"][sel])
    coefs[6:12] *= 1.0 + 1e-3                                     # second copy: different values, same shapes
    P = len(sel)
    report = {"world": world, "cases": []}

    def run_case(name, coefs, z0):
        P = coefs.shape[0]
        full_ms, full_am, full_mv = eng.checked(lambda: eng.greedy_step(coefs, z0, wl.T, wl.dt))     # single GPU, whole grid
        b = sharding.shard_bounds(P, world)
        lo, hi = b[rank], b[rank + 1]
        ms, key = sharding.sharded_greedy_step(eng, torch.from_numpy(coefs[lo:hi]).cuda(), torch.from_numpy(z0[lo:hi]).cuda(),
                                               wl.T, wl.dt, lo)
        val, idx = sharding.unpack_maxloc(int(key.item()))
        assert idx == int(full_am), (name, idx, int(full_am))
        assert np.float32(val) == np.float32(float(full_mv)), (name, val, float(full_mv))
        assert torch.equal(ms, full_ms[lo:hi]), name                                 # bits, not a tolerance
        report["cases"].append({"case": name, "P": P, "bounds": b, "selected": idx, "value": float(val)})

    run_case("uneven blocks", coefs, z0)
    # the same maximum on both ranks: parameters 2 and 9 carry identical draws -> the smaller global index wins
    tie = coefs.copy(); tz = z0.copy()
    src = int(np.argmax(eng.greedy_step(coefs, z0, wl.T, wl.dt)[0].cpu().numpy()))
    tie[2] = coefs[src]; tz[2] = z0[src]; tie[9] = coefs[src]; tz[9] = z0[src]
    run_case("tie across ranks", tie, tz)
    assert report["cases"][-1]["selected"] == min(2, src)
    # more ranks than parameters: the empty block contributes key 0 and skips the engine
    run_case("one parameter", coefs[4:5], z0[4:5])
    # a failing rank (the last one asks for an empty time grid: GLASDI_ERR_INVALID) must raise on EVERY rank after the
    # collective instead of leaving the others waiting in it
    b = sharding.shard_bounds(P, world)
    raised = None
    try:
        sharding.sharded_greedy_step(eng, torch.from_numpy(coefs[b[rank]:b[rank + 1]]).cuda(),
                                     torch.from_numpy(z0[b[rank]:b[rank + 1]]).cuda(), 0 if rank == world - 1 else wl.T,
                                     wl.dt, b[rank])
    except Exception as e:                                         # noqa: BLE001
        raised = type(e).__name__
    assert raised == ("GlasdiError" if rank == world - 1 else "RuntimeError"), raised
    report["cases"].append({"case": "failing rank raises everywhere", "raised_on_rank0": raised})
    dist.barrier()
    if rank == 0:
        report["ok"] = True
        print(json.dumps(report), flush=True)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
