"""bench.py's algorithmic-work accounting against the worked totals of SURVEY.md section 8(d) / BASELINE.md section 3
(the figures `roofline.achieved` is built from), and the shape of the reference-arm JSON line."""
import json
import subprocess
import sys

import bench


def test_algorithmic_flops_match_the_survey_totals():
    # DCV5T def2-tzvp: N = 1249, Naux = 3177, homo = 143 -> m = q = 431, B = 41 328
    f = bench.algorithmic_flops(1249, 3177, 143, {"gw_iterations": 1, "bse_operator_columns": 20})
    assert abs(f["fill"] / 1e12 - 8.5) < 0.1
    assert abs(f["mul_right"] / 3 / 1e12 - 10.9) < 0.1           # per call; G0W0 makes three (L, phi, U)
    assert abs(f["epsilon"] / 3 / 1e12 - 1.6) < 0.05             # per frequency
    assert abs(f["sigma_x"] / 1e12 - 0.085) < 0.002
    assert abs(f["bse_matvec"] / 1e12 - 2.3) < 0.05              # one factorised singlet-TDA block product, k = 20
    # C60 def2-tzvp: N = 1860, Naux = 4560, homo = 179
    g = bench.algorithmic_flops(1860, 4560, 179, {"gw_iterations": 1, "bse_operator_columns": 20})
    assert abs(g["fill"] / 1e12 - 34.0) < 0.2 and abs(g["mul_right"] / 3 / 1e12 - 41.7) < 0.2
    assert abs(g["epsilon"] / 3 / 1e12 - 6.3) < 0.1 and abs(g["bse_matvec"] / 1e12 - 6.4) < 0.1


def test_reference_arm_line_on_a_small_workload():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--workload", "tiny", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=bench.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is False and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert set(cb["factorised"]["stages"]) == {"fill_3c", "bse_hd", "bse_hx"} and cb["factorised"]["value"] > 0
