#!/usr/bin/env python
"""Summarises an `ncu --set full` report of one kernel into a small JSON under profiles/ (what bench.py's
`roofline.traffic` reads): duration, DRAM bytes, pipe utilisations, issue rate, plus the sha1 of the kernel's source file so
that a stale capture is never mistaken for a measurement of the current kernel.

    python tools/ncu_summary.py gpurun_out/X.ncu-rep <kernel-name regex> gplasdi_b200/csrc/FILE.cu profiles/OUT.json ["how"]
"""
import csv, hashlib, io, json, os, re, subprocess, sys

rep, kre, src, out = sys.argv[1:5]
how = sys.argv[5] if len(sys.argv) > 5 else "ncu --set full --clock-control none"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, units = rows[0], rows[1]
sel = [r for r in rows[2:] if re.search(kre, r[H.index("Kernel Name")])]
assert sel, "no launch matches " + kre


def val(r, name):
    if name not in H:
        return None
    i = H.index(name)
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return None
    u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u)
    return v * scale if scale and ("bytes" in name or "time_duration" in name) else v


WANT = {
    "duration_s": "gpu__time_duration.sum",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "inst_executed": "smsp__inst_executed.sum",
    "registers_per_thread": "launch__registers_per_thread",
    "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_sectors_read": "lts__t_sectors_op_read.sum",
    "l2_sectors_write": "lts__t_sectors_op_write.sum",
}
launches = []
for r in sel:
    d = {k: val(r, n) for k, n in WANT.items()}
    d["kernel"] = r[H.index("Kernel Name")]
    if d["dram_bytes_read"] is not None and d["dram_bytes_write"] is not None:
        d["dram_bytes"] = d["dram_bytes_read"] + d["dram_bytes_write"]
    launches.append(d)
res = {"report": os.path.basename(rep), "how": how, "source": src,
       "source_sha1": hashlib.sha1(open(src, "rb").read()).hexdigest(), "launches": launches,
       "dram_bytes_per_launch": sum(l.get("dram_bytes") or 0.0 for l in launches) / len(launches),
       "note": "profiler times are cold-cache and serialised: shares and counters count, not the absolute durations"}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res)[:600])
