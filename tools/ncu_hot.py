#!/usr/bin/env python3
"""Top stall locations of one profiled launch: `ncu_hot.py report.ncu-rep <kernel regex> [launch-skip] [top N]`."""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:160])
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[1:]))) if (r["# Samples"] or "0").isdigit()]
for r in rows:
    r["# Samples"] = r["# Samples"] or "0"
tot = sum(int(r["# Samples"] or 0) for r in rows)
tot_inst = sum(int(r["Instructions Executed"] or 0) for r in rows)
print("samples", tot, "warp instructions", tot_inst)
idx = {id(r): i for i, r in enumerate(rows)}
for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:top]:
    print(f'{idx[id(r)]:5d} {int(r["# Samples"]):7d} {100*int(r["# Samples"])/max(tot,1):5.1f}%  ex={int(r["Instructions Executed"] or 0):9d}  {r["Source"][:110]}')
