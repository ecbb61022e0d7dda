#!/usr/bin/env python3
"""Quick stage-1 throughput probe: python tools/perf_probe.py <composite> <curves> <B1>"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avx_ecm_b200 as E

name, curves, b1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
comp = json.load(open(os.path.join(ROOT, "tests/golden/composites.json")))
N = int(comp[name]) if name in comp else int(json.load(open(os.path.join(ROOT, "tests/golden", name + ".json")))["n"])
ctx = E.EcmContext(N, curves)
nl = ctx.nl
t = time.time(); ops, adds, dups = E.plan_stage1(b1); tplan = time.time() - t
modmul = 6 * adds + 5 * dups
W = 2 * nl * nl + nl
t = time.time(); ctx.build_curves(list(range(7, 7 + curves))); tb = time.time() - t
for rep in range(2):
    ctx.build_curves(list(range(7, 7 + curves)))
    t = time.time(); ctx.stage1(b1); wall = time.time() - t
    ms, launches = ctx.last_timing()
    cps = curves / (ms / 1e3)
    print("%s nl=%d curves=%d B1=%d: %.1f ms (%d launches, wall %.2fs, plan %.2fs, build %.2fs) -> %.1f curves/s at this B1; "
          "%.3f Tprod/s (W=%d, modmul/curve=%d); equiv B1=1e6: %.1f curves/s" %
          (name, nl, curves, b1, ms, launches, wall, tplan, tb, cps, cps * modmul * W / 1e12, W, modmul, cps * modmul / 12974547), flush=True)
ctx.close()
