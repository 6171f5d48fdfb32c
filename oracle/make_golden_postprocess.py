#!/usr/bin/env python3
"""Run the UNMODIFIED reference `scripts/aicity_inf.py` on the synthetic 3-view scores of
tests/golden/postprocess_case.py and store its submission file as tests/golden/aicity_inf_expected.txt.

    python oracle/make_golden_postprocess.py        (needs /root/reference or the vendored oracle/_ref)

TEST INFRASTRUCTURE — not product code.  The reference script runs through `python -m aicity_action_b200.launch --no-patch`,
i.e. with only its missing third-party imports (matplotlib) stubbed."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
from tests.golden.postprocess_case import write_inputs  # noqa: E402


def run_reference(workdir, extra=()):
    pkl, thr, csv = write_inputs(workdir)
    out = os.path.join(workdir, "submission.txt")
    script = os.path.join(ref_shims.REFERENCE_ROOT, "scripts", "aicity_inf.py")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "aicity_action_b200.launch", "--no-patch", "--reference-root",
                        ref_shims.REFERENCE_ROOT, script, pkl, thr, csv, out, *extra], cwd=ROOT, env=env,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-3000:])
    return open(out).read()


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as d:
        text = run_reference(d)
    with tempfile.TemporaryDirectory() as d:
        text_max = run_reference(d, ["--agg_method", "max", "--chunk_sort_base_single_vid", "length",
                                     "--chunk_sort_base_multi_vid", "score", "--use_num_chunk", "2"])
    for name, t in (("aicity_inf_expected.txt", text), ("aicity_inf_expected_max_len_score_k2.txt", text_max)):
        with open(os.path.join(ROOT, "tests", "golden", name), "w") as f:
            f.write(t)
        print(name, len(t.splitlines()), "segments")
