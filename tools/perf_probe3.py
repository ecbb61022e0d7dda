#!/usr/bin/env python3
"""Stage-2 probe at full-size geometry without paying for a full stage 1: stage 1 runs at a small B1 (any point is a
valid Q), stage 2 with the programs of (B1, B2).   python tools/perf_probe3.py <composite> <curves> <B1> <B2> [small_b1]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avx_ecm_b200 as E
sys.argv += [""] * 6
name, curves, b1, b2 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
sb1 = int(sys.argv[5] or 2000)
comp = json.load(open(os.path.join(ROOT, "tests/golden/composites.json")))
N = int(comp[name]) if name in comp else int(json.load(open(os.path.join(ROOT, "tests/golden", name + ".json")))["n"])
ctx = E.EcmContext(N, curves)
nl = ctx.nl
W = 2 * nl * nl + nl
ctx.build_curves(list(range(7, 7 + curves)))
ctx.stage1(sb1)
for rep in range(int(os.environ.get("REPS", "1"))):
    t = time.time(); ctx.stage2(b1, b2); w2 = time.time() - t
    ms2, l2 = ctx.last_timing()
    c = ctx.stage2_counters()
    mm = 0
    def count(words):
        n = 0
        for w in words:
            op = w & 0xff
            n += 1 if op in (0, 1, 10) else 2 if op == 12 else 0
        return n
    mm = count(E.stage2_program(b1, b2, -1)[0])
    which = 0
    while b1 + which * 100000000 < b2:
        mm += count(E.stage2_program(b1, b2, which)[0]); which += 1
    print("%s nl=%d curves=%d stage 2 (B1=%d, B2=%d): device %.0f ms (%d launches), wall %.2f s | %s | modmul/curve %d | %.3f Tprod/s all products | env %s" %
          (name, nl, curves, b1, b2, ms2, l2, w2, c, mm, curves / (ms2 / 1e3) * mm * W / 1e12,
           {k: v for k, v in os.environ.items() if k.startswith("ECM_B200")}), flush=True)
ctx.close()
