"""Known-answer regressions of the reference at full size, on the GPU through the C ABI.

The cases are lines of the reference's own regression scripts -- test.csh:1-41 and test_t35.csh:1-49 -- and the README
example (BASELINE.json configs[0]: sigma 1007 finds 272602401466814027129 in stage 1 at B1 = 1e6), each run by the
compiled reference with the script's own B1, B2 and sigma (tools/gen_golden.py -> tests/golden/{readme508_b1_1e6,
t35_full_*,csh_line*}.json).  Compared per curve: save_b1.txt lines byte for byte, stage-1 Z, the stage-2 accumulator,
every reported factor (stage, sigma) and the op counters.

Eight curves occupy one warp, so a case takes tens of seconds however fast the kernels are; all cases run at once,
each on its own context (own stream, own tables) from its own host thread, which also exercises what the CLI does
with one thread per GPU -- here several contexts share one GPU."""
import threading
import pytest
from conftest import GOLDEN, golden_factor, golden_base, is_known_answer, is_slow, SLOW
import avx_ecm_b200 as E

pytestmark = pytest.mark.gpu


def run_case(g):
    N, b1, b2, s0 = int(g["n"]), g["b1"], g["b2"], int(g["sigma0"])
    lanes = len(g["save_lines"])
    ctx = E.EcmContext(N, lanes, base=golden_base(g))
    try:
        r = E.vececm(N, lanes, b1, b2 if b2 > b1 else b1, sigma=s0, ctx=ctx)
        cnt = ctx.stage2_counters() if b2 > b1 else {}
    finally:
        ctx.close()
    errs = []
    if r["save_lines"] != g["save_lines"]:
        errs.append("save_b1.txt lines differ")
    if r["z"] != [int(z, 16) for z in g["z1_true_hex"]]:
        errs.append("stage-1 Z differs")
    ref = g["counts"]
    if cnt and cnt != {k: ref[k] for k in cnt}:
        errs.append("stage-2 counters %r != %r" % (cnt, {k: ref[k] for k in cnt}))
    exp_acc = [int(a, 16) for a in g["acc_true_hex"]]
    for i in range(lanes):
        sigma = s0 + i
        got = {(s, st): f for s, st, f in r["factors"] if s == sigma}
        for st in (1, 2):
            if got.get((sigma, st), 0) != golden_factor(g, sigma, st):
                errs.append("sigma %d stage %d: factor %d, reference %d" % (sigma, st, got.get((sigma, st), 0), golden_factor(g, sigma, st)))
        if b2 > b1 and (not r["inv_fail"][i] or i == 0) and r["acc"][i] != exp_acc[i]:
            errs.append("lane %d: stage-2 accumulator differs" % i)
    return errs


# "huge B1" (1.1e8 / 1.34e10) and "huge B2" (7e6 / 1.6e10), test.csh:33-37: 1.4e9 resp. 4.7e8 sequential modular products per
# curve are 0.5-2 HOURS on a GPU whatever the batch size (a warp retires one product per ~5 us); their golden vectors pin
# the oracle on the CPU (tests/test_oracle_vs_golden.py with ECM_B200_SLOW=1) and the planner's op streams for such
# bounds are compared with the oracle's in tests/test_stage1_ranges_cpu.py and tests/test_cabi_cpu.py.
CPU_ONLY = ("slow_csh_line26", "slow_csh_line27")


def test_known_answers_of_the_reference_at_full_size():
    names = sorted(k for k in GOLDEN if is_known_answer(k) and (SLOW or not is_slow(k)) and k not in CPU_ONLY
                   and int(GOLDEN[k]["n"]).bit_length() <= 2048)
    assert len(names) >= (15 if SLOW else 8)
    results = {}

    def work(name):
        try:
            results[name] = run_case(GOLDEN[name])
        except Exception as e:          # noqa: a thread must report, not die silently
            results[name] = ["exception: %r" % (e,)]
    threads = [threading.Thread(target=work, args=(n,)) for n in names]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    bad = {k: v for k, v in results.items() if v}
    assert not bad, bad
    # the answers the scripts are there for
    found = {(k, f["factor"]) for k in names for f in GOLDEN[k]["factors"]}
    assert ("readme508_b1_1e6", "272602401466814027129") in found
    assert ("t35_full_sigma_11919771003873180376", "1147161816393958657432308670357") in found
    assert ("t35_full_sigma_10019108749973911965", "1147161816393958657432308670357") in found
