"""Run one of the reference's entry scripts, unmodified, on the B200 path.

    python -m aicity_action_b200.launch [--compute auto|bf16|fp32] [--reference-root DIR] [--no-patch] \\
        tools/run_net.py --cfg configs/Aicity/MVITV2_FULL_B_16x4_CONV_448.yaml TRAIN.ENABLE True ...
    python -m aicity_action_b200.launch scripts/run_action_classification_temporal_inf.py videos.lst video_dir \\
        model.pyth out_dir --model_dataset aicity --frame_size 448 --pyslowfast_cfg configs/Aicity/..._448.yaml

What it does, in order, before handing control to the script with `runpy` (`__name__ == "__main__"`, `sys.argv` rebuilt):

1. puts the reference checkout (`--reference-root`, `$AICITY_REF`, or the nearest ancestor of the script that contains
   `slowfast/`) and the script's own directory on `sys.path`, as `python script.py` run from the checkout would;
2. `depshims.install()`: stand-ins for third-party packages the reference imports but this environment lacks (fvcore,
   iopath, fairscale, decord, ...) and for `slowfast.visualization`, which the fork imports (tools/train_net.py:22) but
   does not ship — only for roots that are genuinely not importable;
3. `patch.install(compute_dtype=...)`: `MODEL_REGISTRY["MViT"]` and `slowfast.models.attention.*` are rebound to the
   drop-in modules, so `build_model(cfg)` inside the script constructs the sm_100a path.

`--compute` (or `$MVIT_B200_COMPUTE`) is the arithmetic the drop-in uses for fp32 input tensors: the reference's scripts
feed fp32 clips with no autocast region (scripts/module_wrapper.py:606-608), which `auto` serves with the fp32 kernels
(1e-4 parity); `bf16` stores activations in bf16 and runs the tcgen05 kernels (2e-2 parity, the throughput path).
"""
from __future__ import annotations

import argparse
import os
import runpy
import sys


def find_reference_root(script: str, explicit: str | None = None) -> str | None:
    cands = [explicit, os.environ.get("AICITY_REF")]
    d = os.path.dirname(os.path.abspath(script))
    while d and d != os.path.dirname(d):
        cands.append(d)
        d = os.path.dirname(d)
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "slowfast", "models")):
            return os.path.abspath(c)
    return None


def prepare(script: str, reference_root: str | None = None, compute: str | None = None, do_patch: bool = True) -> dict:
    """Steps 1-3 above without running the script (tests call this, then import the script's modules themselves)."""
    from . import depshims, patch

    root = find_reference_root(script, reference_root)
    if root is None:
        raise FileNotFoundError(f"no reference checkout (a directory containing slowfast/models) found for {script}; "
                                "pass --reference-root or set AICITY_REF")
    for p in (root, os.path.dirname(os.path.abspath(script))):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    stubbed = depshims.install()
    done = patch.install(compute_dtype=compute) if do_patch else {"registry": False, "attention": False}
    return {"reference_root": root, "stubbed": stubbed, "patched": done}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m aicity_action_b200.launch", description=__doc__.split("\n\n")[0])
    ap.add_argument("--compute", choices=["auto", "bf16", "fp32"], default=None,
                    help="arithmetic for fp32 input tensors (default: $MVIT_B200_COMPUTE or auto)")
    ap.add_argument("--reference-root", default=None)
    ap.add_argument("--no-patch", action="store_true", help="only install the dependency stand-ins (stock reference model)")
    ap.add_argument("script")
    ap.add_argument("args", nargs=argparse.REMAINDER)
    ns = ap.parse_args(argv)
    info = prepare(ns.script, ns.reference_root, ns.compute, not ns.no_patch)
    print(f"[aicity_action_b200.launch] reference {info['reference_root']}; stubbed deps {info['stubbed']}; "
          f"patched {info['patched']}", file=sys.stderr)
    if not ns.no_patch:
        import atexit

        from . import ops
        atexit.register(lambda: print(f"[aicity_action_b200.launch] libmvit_b200.so kernel launches issued: "
                                      f"{ops.launch_count}", file=sys.stderr))
    sys.argv = [ns.script] + list(ns.args)
    runpy.run_path(ns.script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
