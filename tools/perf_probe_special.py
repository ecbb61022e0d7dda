#!/usr/bin/env python3
"""Stage-1 (+ optional stage-2) throughput on special-form bases, shift-and-fold kernels vs the Montgomery
kernels on the same base:   python tools/perf_probe_special.py <k> <kind> <c> <curves> <B1> [B2]
(kind 1: 2^k-c, kind -1: 2^k+1).  Prints curves/s for both and the ratio."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avx_ecm_b200 as E

k, kind, c, curves, b1 = (int(x) for x in sys.argv[1:6])
b2 = int(sys.argv[6]) if len(sys.argv) > 6 else 0
M = (1 << k) - c if kind > 0 else (1 << k) + 1
ops, adds, dups = E.plan_stage1(b1)
modmul = 6 * adds + 5 * dups
res = {}
for mode in ("fold", "montgomery"):
    if mode == "montgomery":
        os.environ["ECM_B200_NO_FOLD"] = "1"
    ctx = E.EcmContext(M, curves, base=M)
    assert ctx.uses_fold == (mode == "fold")
    best = None
    for rep in range(2):
        ctx.build_curves(list(range(1000003, 1000003 + curves)))
        ctx.stage1(b1)
        ms, launches = ctx.last_timing()
        best = ms if best is None else min(best, ms)
    x, z, _ = ctx.read_stage1()
    line = "%s: 2^%d%s%d nl=%d curves=%d B1=%d: %.1f ms -> %.1f curves/s (%.3f T modmul-products/s counted as 2n^2+n)" % (
        mode, k, "-" if kind > 0 else "+", c, ctx.nl, curves, b1, best, curves / best * 1e3,
        curves / best * 1e3 * modmul * (2 * ctx.nl ** 2 + ctx.nl) / 1e12)
    if b2 > b1:
        t = time.time(); ctx.stage2(b1, b2); acc, _, _ = ctx.read_stage2(); t2 = time.time() - t
        line += "; stage 2 to %d: %.2f s" % (b2, t2)
        res[mode + "_acc"] = acc
    print(line, flush=True)
    res[mode] = (best, x, z)
    ctx.close()
assert res["fold"][1:] == res["montgomery"][1:], "fold and Montgomery kernels disagree"
if b2 > b1:
    assert res["fold_acc"] == res["montgomery_acc"]
print("speed-up of the fold kernels: %.2fx (results identical)" % (res["montgomery"][0] / res["fold"][0]))
