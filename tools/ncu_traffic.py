#!/usr/bin/env python3
"""Summarise an ncu launch list (profiles/launches_rN.csv) per kernel family and write the attention kernel's DRAM traffic
per launch, which bench.py reports as `roofline.traffic`.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file profiles/launches_r2.csv python bench.py --steps 1 --warmup 3 --no-cuda-graph --skip-extras
    python tools/ncu_traffic.py profiles/launches_r2.csv profiles/ncu_traffic_r2.json [forwards]

`forwards` = how many forwards the capture holds (warm-up + timed + instrumented pass); per-forward numbers are totals
divided by it."""
import collections
import csv
import json
import sys

FAMILIES = [("attention", "attention_tc_kernel"), ("linear", "linear_tc_kernel"), ("linear", "mlp_fused_kernel"),
            ("pool_max", "maxpool_"), ("pool_conv", "pool_tma_kernel"),
            ("pool_conv", "pool_tiled_kernel"), ("pool_max", "pool_generic"), ("pool_max", "pool_kernel"),
            ("layernorm", "layernorm"), ("fold_clip", "fold_clip"), ("mean_head", "mean_head")]


def family(name):
    for fam, key in FAMILIES:
        if key in name:
            return fam
    return "other:" + name.split("(")[0][-40:]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    forwards = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = []
    with open(src) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    per = collections.defaultdict(dict)          # launch id -> metric -> value
    names = {}
    for r in csv.DictReader(lines):
        per[r["ID"]][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * \
            {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3,
             "second": 1e6}.get(r["Metric Unit"], 1.0)
        names[r["ID"]] = r["Kernel Name"]
    fam = collections.defaultdict(lambda: dict(launches=0, us=0.0, dram_bytes=0.0))
    for i, m in per.items():
        f = fam[family(names[i])]
        f["launches"] += 1
        f["us"] += m.get("gpu__time_duration.sum", 0.0)
        f["dram_bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    total_us = sum(f["us"] for f in fam.values())
    out = {"source": src, "forwards": forwards or None, "families": {}}
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        out["families"][k] = dict(launches=f["launches"], ms=f["us"] / 1e3, share=f["us"] / total_us,
                                  dram_GB=f["dram_bytes"] / 1e9)
    a = fam.get("attention")
    if a and a["launches"]:
        out["kernel"] = "attention_tc_kernel"
        out["launches"] = a["launches"]
        out["avg_dram_bytes_per_launch"] = a["dram_bytes"] / a["launches"]
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    for k, v in out["families"].items():
        print(f"{k:28s} {v['launches']:5d} launches {v['ms']:9.3f} ms {100 * v['share']:5.1f} %  {v['dram_GB']:8.3f} GB")


if __name__ == "__main__":
    main()
