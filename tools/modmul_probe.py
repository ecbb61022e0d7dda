#!/usr/bin/env python3
"""Raw modmul throughput vs resident warps: python tools/modmul_probe.py <composite> [repeat]"""
import json, os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avx_ecm_b200 as E
name = sys.argv[1]; repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
N = int(json.load(open(os.path.join(ROOT, "tests/golden/composites.json")))[name])
rng = random.Random(1)
for wps in (4, 8, 12, 16, 24, 32, 48):
    count = 148 * 32 * wps
    ctx = E.EcmContext(N, count)
    nl = ctx.nl; W = 2 * nl * nl + nl
    a = [rng.randrange(N) for _ in range(64)] * (count // 64); b = a[1:] + a[:1]
    for op in (0, 1):
        ctx.fieldop(op, a, b, repeat=100)
        ctx.fieldop(op, a, b, repeat=repeat)
        ms, _ = ctx.last_timing()
        rate = count * (repeat + 3) / (ms / 1e3)
        print("%s nl=%d op=%d warps/SM=%d: %.1f ms, %.3f Tprod/s, %.1f prod/clk/SM @1.965GHz" % (name, nl, op, wps, ms, rate * W / 1e12, rate * W / 148 / 1.965e9), flush=True)
    ctx.close()
