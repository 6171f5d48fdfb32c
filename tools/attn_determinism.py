#!/usr/bin/env python3
"""Run-to-run determinism of the fused attention kernel on fixed inputs, per shape and score magnitude.

    python tools/attn_determinism.py [--runs 30]

For every (shape, sigma) the same q/k/v go through `ops.attention` `runs` times; outputs are compared bit for bit with the
first one.  For a mismatching run the script prints where the differences sit (rows, 32-row warps, columns) and how far
each variant is from an fp32 reference on those rows.  sigma scales q and k: sigma >= 3 makes row maxima grow by more than
2^8 after the first key tile, i.e. exercises the lazy-rescale path of the softmax warps.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicity_action_b200 import ops  # noqa: E402


def ref_rows(q, k, v, scale, add_q, b, h, rows):
    qq = q[b, h, rows].float()
    s = (qq @ k[b, h].float().T) * scale
    o = torch.softmax(s, -1) @ v[b, h].float()
    return o + qq if add_q else o


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=30)
    ap.add_argument("--out", default="gpurun_out/attn_determinism.json")
    ap.add_argument("--quick", action="store_true", help="two shapes, sigma 2 and 3, plus the error of EVERY run against fp32")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    scale = 96 ** -0.5
    shapes = [(8, 4, 6272, 1568), (8, 2, 25088, 1568), (8, 8, 1568, 1568), (8, 1, 100352, 1568)]
    results = []
    sigmas = (1.0, 2.0, 3.0, 4.0)
    if args.quick:
        shapes, sigmas = [(8, 4, 6272, 1568), (8, 8, 1568, 1568)], (2.0, 3.0)
    for (B, h, Lq, Lk) in shapes:
        for sigma in sigmas:
            torch.manual_seed(1)
            q = (torch.randn(B, h, Lq, 96, device=dev) * sigma).bfloat16()
            k = (torch.randn(B, h, Lk, 96, device=dev) * sigma).bfloat16()
            v = torch.randn(B, h, Lk, 96, device=dev).bfloat16()
            first = ops.attention(q, k, v, scale, True).clone()
            bad, detail = 0, None
            nref = min(Lq, 1024)
            full_ref = torch.stack([ref_rows(q, k, v, scale, True, 0, hh, torch.arange(nref, device=dev)) for hh in range(h)], 1)
            worst_err = float((first.view(B, Lq, h, 96)[0, :nref].float() - full_ref).abs().max())
            for _ in range(args.runs):
                o = ops.attention(q, k, v, scale, True)
                worst_err = max(worst_err, float((o.view(B, Lq, h, 96)[0, :nref].float() - full_ref).abs().max()))
                if not torch.equal(o, first):
                    bad += 1
                    if detail is None:
                        d = (o.view(B, Lq, h, 96) != first.view(B, Lq, h, 96))
                        rows = d.any(-1).nonzero()                        # [n, 3] = (b, row, head)
                        b0, r0, h0 = (int(x) for x in rows[0])
                        warps = {(int(b), int(hh), int(r) // 32) for b, r, hh in rows[:4096].tolist()}
                        rsel = torch.tensor([int(r) for b, r, hh in rows.tolist() if b == b0 and hh == h0][:64], device=dev)
                        ref = ref_rows(q, k, v, scale, True, b0, h0, rsel)
                        ea = (first.view(B, Lq, h, 96)[b0, rsel, h0].float() - ref).abs().max()
                        eb = (o.view(B, Lq, h, 96)[b0, rsel, h0].float() - ref).abs().max()
                        detail = {"differing_rows": int(rows.shape[0]), "differing_elements": int(d.sum()),
                                  "warps_touched(<=4096 rows)": len(warps),
                                  "rows_per_warp": rows.shape[0] / max(1, len(warps)),
                                  "first": [b0, h0, r0], "row_in_tile": r0 % 128, "q_tile": r0 // 128,
                                  "cols_first_row": d[b0, r0, h0].nonzero().flatten().tolist()[:100],
                                  "max_abs_diff": float((o.float() - first.float()).abs().max()),
                                  "err_first_vs_ref": float(ea), "err_other_vs_ref": float(eb),
                                  "ref_absmax": float(ref.abs().max())}
            rec = {"shape": [B, h, Lq, Lk], "sigma": sigma, "runs": args.runs, "mismatching": bad, "worst_err_vs_fp32(b=0, first 1024 rows)": worst_err,
                   "detail": None if args.quick else detail}
            print(json.dumps(rec), flush=True)
            results.append(rec)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    tag = "_".join(f"{k}{os.environ[k]}" for k in ("MVIT_ATTN_NS", "MVIT_ATTN_POLY") if k in os.environ)
    with open(args.out.replace(".json", f"{('_' + tag) if tag else ''}.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
