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


def golden_factor(g, sigma, stage):
    for f in g["factors"]:
        if int(f["sigma"]) == sigma and f["stage"] == stage:
            return int(f["factor"])
    return 0


def golden_base(g):
    """Base number 2^k-1 / 2^k+1 / 2^k-c of a special-form golden case (None for generic inputs)."""
    return int(g["base"]) if g.get("base") else None
