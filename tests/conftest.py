import glob, json, os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def golden_cases():
    out = {}
    for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.json"))):
        if os.path.basename(f) == "composites.json":
            continue
        g = json.load(open(f))
        out[g["name"]] = g
    return out


def composites():
    c = json.load(open(os.path.join(ROOT, "tests", "golden", "composites.json")))
    return {k: int(v) for k, v in c.items()}


GOLDEN = golden_cases()

# Full-size known answers (test.csh / test_t35.csh lines and the README example at the reference's own B1/B2): tens of
# seconds each for 8 curves on one warp, so the GPU suite runs them all at once (tests/test_gpu_known_answers.py)
# instead of one by one through the per-case tests; "slow_" cases (huge B1 / huge B2 / 1165-bit input) only run with
# ECM_B200_SLOW=1, on the CPU oracle as well as on the GPU.
SLOW = os.environ.get("ECM_B200_SLOW", "") not in ("", "0")


def is_known_answer(name):
    # ... plus the two wide synthetic cases at a B1 that is not a toy (8 curves x 2048 bits: ~20 s per stage on one warp)
    return name.startswith(("readme508_b1_1e6", "t35_full_", "csh_line", "slow_csh_line", "syn1024_b1_1e5", "syn2048_b1_5e4"))


# test.csh lines whose 8 curves keep one warp busy for 1.5-4 minutes (B1 = 3e6, B2 up to 1e9, 24-limb inputs): with the
# "slow_" cases they run under ECM_B200_SLOW=1 (log of such a run on the GPU: profiles/r2_final_known_answers_all.log)
HEAVY = ("csh_line02", "csh_line09", "csh_line11", "csh_line14", "csh_line22", "csh_line25")


def is_slow(name):
    return name.startswith("slow_") or name in HEAVY


def golden_factor(g, sigma, stage):
    for f in g["factors"]:
        if int(f["sigma"]) == sigma and f["stage"] == stage:
            return int(f["factor"])
    return 0


def golden_base(g):
    """Base number 2^k-1 / 2^k+1 / 2^k-c of a special-form golden case (None for generic inputs)."""
    return int(g["base"]) if g.get("base") else None
