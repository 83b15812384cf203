#!/usr/bin/env python
"""Per-kernel SASS evidence from the in-tree objects: how many tcgen05 MMA (UTCHMMA / UTCQMMA), TMEM load/store (LDTM /
STTM), bulk-copy (UBLKCP), tensor-map TMA (UTMALDG), MUFU, FFMA2 and DFMA instructions every kernel holds.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

Runs `cuobjdump -sass` on gplasdi_b200/lib/*.o (no GPU needed).  Counts are static instruction counts of the compiled
code (a loop body counts once): they prove which units a kernel is written for, not how often it uses them."""
import glob, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "MUFU", "FFMA2", "HFMA2", "DFMA", "FFMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    rows = []
    for obj in sorted(glob.glob(os.path.join(ROOT, "gplasdi_b200", "lib", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        cur, counts = None, {}
        for line in sass.split("\n"):
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                counts[cur] = dict.fromkeys(MNEMONICS, 0)
                counts[cur]["total"] = 0
                continue
            if cur is None:
                continue
            m = re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            counts[cur]["total"] += 1
            base = op.split(".")[0]
            if base in counts[cur]:
                counts[cur][base] += 1
        names = demangle(list(counts))
        for k, c in counts.items():
            rows.append((os.path.basename(obj), names[k], c))
    print(f"{'object':18s} {'instr':>7s} " + " ".join(f"{m:>7s}" for m in MNEMONICS) + "  kernel")
    for obj, name, c in rows:
        short = re.sub(r"\(.*", "", name)
        short = short.replace("glasdi::", "")
        if len(short) > 110:
            short = short[:107] + "..."
        print(f"{obj:18s} {c['total']:7d} " + " ".join(f"{c[m]:7d}" for m in MNEMONICS) + "  " + short)


if __name__ == "__main__":
    sys.exit(main())
