#!/usr/bin/env python3
"""Tiny end-to-end run for compute-sanitizer: stage 1 + stage 2 on two limb sizes, checked against the oracle."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import avx_ecm_b200 as E, oracle_lib as O
comp = {k: int(v) for k, v in json.load(open(os.path.join(ROOT, "tests/golden/composites.json"))).items()}
for name, curves, b1, b2 in (("small96", 40, 50, 5000), ("syn415", 500, 300, 20000), ("syn1024", 70, 200, 6000)):
    N = comp[name]
    r = E.vececm(N, curves, b1, b2, sigma=11)
    for i in (0, curves - 1):
        o = O.ecm_curve(N, b1, b2, 11 + i)
        assert r["x"][i] == o["x"] and r["z"][i] == o["z"] and r["acc"][i] == o["acc"], (name, i)
    print(name, "ok", flush=True)
