"""bench.py --impl reference (the unmodified reference on the host cores) runs without a GPU: check the
JSON contract of that arm here.  Needs oracle/_ref (built by __graft_entry__.build() when the reference
tree is present) and an AVX-512 host; otherwise the arm times the oracle port and says so."""
import json, os, subprocess, sys
from conftest import ROOT


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, check=True, timeout=900).stdout.strip().splitlines()[-1]
    line = json.loads(out)
    assert line["impl"] == "reference" and line["metric"] == "stage1_curves_per_sec_B1_1e6_415bit"
    assert line["unit"] == "curves/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0 and line["steps"] == 1
