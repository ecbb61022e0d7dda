"""Pins the CPU oracle (oracle/ecm_oracle.c) against vectors produced by the compiled,
unmodified reference (tools/gen_golden.py -> tests/golden/*.json): save_b1.txt lines byte
for byte, stage-1 Z, stage-2 accumulator, reported factors and the reference's op counters.
"""
import pytest
from conftest import GOLDEN, golden_factor, golden_base, is_known_answer, is_slow, SLOW
import oracle_lib as O


def lanes_for(g):
    n = len(g["save_lines"])
    if is_known_answer(g["name"]) and g["factors"]:      # the lane(s) that found the factor
        return sorted({int(f["sigma"]) - int(g["sigma0"]) for f in g["factors"]})
    if is_known_answer(g["name"]):
        return [0, n - 1]
    if g["b1"] >= 1000000:            # ~5-8 s per curve on the CPU: keep lane 0 (+ the factor lane)
        keep = {0}
        for f in g["factors"]:
            keep.add(int(f["sigma"]) - int(g["sigma0"]))
        return sorted(keep)
    if "two_ranges" in g["name"]:
        return [0]
    return list(range(n))


# Full-size known answers cost 10-60 s per curve here; the default suite keeps three of them (the README example with
# its stage-1 find, one t35 line, one test.csh line with a stage-2 find) and ECM_B200_SLOW=1 runs them all.
CPU_DEFAULT_KNOWN = {"readme508_b1_1e6", "t35_full_sigma_10019108749973911965", "csh_line19", "syn1024_b1_1e5", "syn2048_b1_5e4"}


def selected(name, g):
    if is_known_answer(name):
        return SLOW or name in CPU_DEFAULT_KNOWN
    return g["b1"] <= 100000000      # B1 > 1e8: minutes per curve, tests/test_stage1_ranges_cpu.py (ECM_B200_SLOW=1) covers it


@pytest.mark.parametrize("name", sorted(k for k, g in GOLDEN.items() if selected(k, g)))
def test_oracle_matches_reference(name):
    g = GOLDEN[name]
    N, b1, b2 = int(g["n"]), g["b1"], g["b2"]
    c = g["counts"]
    for i in lanes_for(g):
        sigma = int(g["sigma0"]) + i
        r = O.ecm_curve(N, b1, b2, sigma, M=golden_base(g))
        assert r["save_line"] == g["save_lines"][i]
        assert r["z"] == int(g["z1_true_hex"][i], 16)
        assert r["f1"] == golden_factor(g, sigma, 1)
        assert r["counters"][:2] == [c["s1_ptadds"], c["s1_ptdups"]]
        if b2 > b1:
            assert r["counters"][2:6] == [c["s2_ptadds"], c["s2_numinv"], c["s2_paired"], c["pairmap_steps"]]
            assert r["f2"] == golden_factor(g, sigma, 2)
            if not r["counters"][6] or i == 0:
                # lanes that met a non-invertible element are only bit-comparable in vector
                # lane 0 (see batch_invert() in oracle/ecm_oracle.c)
                assert r["acc"] == int(g["acc_true_hex"][i], 16)


def test_known_op_counts():
    # exact, N-independent stage-1 op counts printed by the reference (ecm.c:1849)
    tr = O.stage1_trace(100000)
    adds = sum(tr.count(c) for c in b"3459F")
    dups = sum(tr.count(c) for c in b"ID459")
    g = GOLDEN["syn415_b1_1e5"]["counts"]
    assert (adds, dups) == (g["s1_ptadds"], g["s1_ptdups"])


def test_stage2_map_sizes():
    m = O.stage2_map(1000000)
    assert max(m) == 7682 and len(m) == 16 * 2311 + 3     # 7683 stored points
    assert O.lib().oracle_stage2_D(4096) == 1155 and O.lib().oracle_stage2_D(4097) == 2310
