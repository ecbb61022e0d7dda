#!/usr/bin/env python3
"""Stage-1 + stage-2 throughput probe: python tools/perf_probe2.py <composite> <curves> <B1> <B2>"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avx_ecm_b200 as E

name, curves, b1, b2 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
comp = json.load(open(os.path.join(ROOT, "tests/golden/composites.json")))
N = int(comp[name]) if name in comp else int(json.load(open(os.path.join(ROOT, "tests/golden", name + ".json")))["n"])
ctx = E.EcmContext(N, curves)
nl = ctx.nl
W = 2 * nl * nl + nl
ctx.build_curves(list(range(7, 7 + curves)))
t = time.time(); ctx.stage1(b1); w1 = time.time() - t
ms1, l1 = ctx.last_timing()
t = time.time(); ctx.stage2(b1, b2); w2 = time.time() - t
ms2, l2 = ctx.last_timing()
c = ctx.stage2_counters()
# modmuls of stage 2: 6 per point add, 1 per pair, 4 per inverted element (approx: table sizes), inversions extra
mm2 = 6 * c["s2_ptadds"] + c["s2_paired"]
print("%s nl=%d curves=%d B1=%d B2=%d: stage1 %.0f ms (%d launches) | stage2 device %.0f ms (%d launches), wall %.2f s (host planning %.2f s) | "
      "stage2 counters %s | stage-2 rate %.1f curves/s, >= %.3f Tprod/s (adds+pairs only)" %
      (name, nl, curves, b1, b2, ms1, l1, ms2, l2, w2, w2 - ms2 / 1e3, c, curves / (ms2 / 1e3), curves / (ms2 / 1e3) * mm2 * W / 1e12), flush=True)
ctx.close()
