"""bench.py's algorithmic-FLOP accounting reproduces SURVEY.md section 8(d)'s table (the figure `roofline.achieved` and
`step_achieved` are computed from), and the reference arm prints the contract's JSON keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("dims,N,M,S,step_gflop", [
    ([8, 8, 1], 1000, 100, 20, 3.159),                 # config 2
    ([8, 8, 8, 8, 8, 1], 1000, 100, 20, 36.647),       # config 3 (north-star)
    ([9, 9, 9, 1], 4096, 512, 32, 2577.4),             # config 4
    ([784, 30, 10], 1000, 100, 10, 9.682),             # config 5
])
def test_algorithmic_flops_match_survey_table(dims, N, M, S, step_gflop):
    fwd, fixed, step = bench.algorithmic_flops(dims, N, M, S)
    assert abs(step / 1e9 - step_gflop) <= 5e-4 * step_gflop + 5e-4
    assert len(fwd) == len(dims) - 1 and abs(3.0 * (sum(fwd) + fixed) - step) < 1.0


def test_north_star_per_row_figure():
    # f(l) = 2 M D_in + 2 M^2 + 2 M D_out + 2 D_out M^2 + 2 D_out M = 184 800 flops per row of an inner layer
    fwd, _, _ = bench.algorithmic_flops([8, 8, 8, 8, 8, 1], 1000, 100, 20)
    assert fwd[1] == 20000 * 184800


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0 and "workload" in line["config"]
