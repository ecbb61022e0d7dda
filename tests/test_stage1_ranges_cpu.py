"""Stage 1 over several prime ranges (B1 > 1e8 in the reference: vececm's loop ecm.c:1207-1311 calls ecm_stage1
once per range of 1e8, repeating the doublings and skipping each later range's first prime, and writes
checkpoint.txt in between).  The range width is a test hook on both sides (ECM_B200_S1_RANGE for the planner,
oracle_set_prime_range for the oracle) so that the whole mechanism runs at small B1; the real width is pinned by
the golden vector syn206_b1_1.1e8_two_stage1_ranges (slow tests at the bottom, ECM_B200_SLOW=1)."""
import os
import pytest
from conftest import GOLDEN, composites
import oracle_lib as O
import avx_ecm_b200 as E
from test_programs_cpu import run_stage1_stream

TYPE_CH = "DI3459FN"
SLOW = pytest.mark.skipif(not os.environ.get("ECM_B200_SLOW"), reason="minutes of CPU; set ECM_B200_SLOW=1")


@pytest.fixture
def prime_range(monkeypatch):
    def set_range(width):
        monkeypatch.setenv("ECM_B200_S1_RANGE", str(width))
        O.set_prime_range(width)
    yield set_range
    O.set_prime_range(0)
    O.set_checkpoint(0)


@pytest.mark.parametrize("width,b1", [(20000, 50000), (1000, 3500), (4096, 4097), (16, 200), (30000, 30000), (30000, 30001)])
def test_plan_with_ranges_matches_oracle_trace(prime_range, width, b1):
    prime_range(width)
    ops, adds, dups = E.plan_stage1(b1)
    mine = "".join(TYPE_CH[b & 7] for b in ops).replace("N", "")
    ref = O.stage1_trace(b1).decode().replace("S", "")
    assert mine == ref
    nranges = (b1 + width - 1) // width
    assert E.stage1_ranges(b1) == nranges
    # the quirks are really there: the doublings once per range
    ndbl = len([1 for q in range(1, 64) if 2 ** q < b1])
    assert mine.count("D") == ndbl * nranges


def test_ranges_change_the_residues_like_the_oracle_says(prime_range):
    N = composites()["syn415"]
    x, s = O.build_curve(N, 77)
    one = O.ecm_curve(N, 9000, 9000, 77)
    prime_range(2500)
    ops, _, _ = E.plan_stage1(9000)
    X, Z = run_stage1_stream(N, x, s, ops)
    four = O.ecm_curve(N, 9000, 9000, 77)
    assert (X, Z) == (four["x"], four["z"])
    assert (four["x"], four["z"]) != (one["x"], one["z"])


def test_oracle_checkpoints_are_prefixes(prime_range):
    # stopping after r ranges == running the stream up to the end of range r
    N = composites()["t35"]
    prime_range(3000)
    x, s = O.build_curve(N, 123)
    ops, _, _ = E.plan_stage1(10000)
    # range ends in the op stream: each range starts with the block of doublings
    ndbl = len([1 for q in range(1, 64) if 2 ** q < 10000])
    starts = [i for i in range(len(ops)) if all((ops[i + j] & 7) == 0 for j in range(ndbl)) and (i == 0 or (ops[i - 1] & 7) != 0)]
    assert len(starts) == 4
    for r in range(1, 4):
        O.set_checkpoint(r)
        o = O.ecm_curve(N, 10000, 10000, 123)
        assert run_stage1_stream(N, x, s, ops[:starts[r]]) == (o["x"], o["z"])
    O.set_checkpoint(0)


@SLOW
@pytest.mark.timeout(3600)
def test_plan_matches_oracle_trace_beyond_1e8():
    b1 = 110000000
    ops, adds, dups = E.plan_stage1(b1)
    ref = O.stage1_trace(b1)
    mine = bytes(TYPE_CH[b & 7].encode()[0] for b in ops if (b & 7) != 7)
    assert mine == ref.replace(b"S", b"")
    g = GOLDEN["syn206_b1_1.1e8_two_stage1_ranges"]["counts"]
    assert (adds, dups) == (g["s1_ptadds"], g["s1_ptdups"])


@SLOW
@pytest.mark.timeout(7200)
def test_oracle_matches_reference_beyond_1e8():
    g = GOLDEN["syn206_b1_1.1e8_two_stage1_ranges"]
    N, b1, s0 = int(g["n"]), g["b1"], int(g["sigma0"])
    for lane in (0, 5):
        r = O.ecm_curve(N, b1, b1, s0 + lane)
        assert r["save_line"] == g["save_lines"][lane]
        assert r["counters"][:2] == [g["counts"]["s1_ptadds"], g["counts"]["s1_ptdups"]]
    O.set_checkpoint(1)
    try:
        r = O.ecm_curve(N, b1, b1, s0)
        line = g["checkpoint_lines"][0]
        assert "B1=99999989;" in line                                  # the last prime of the first range
        assert ("X=0x%x; Z=0x%x;" % (r["x"], r["z"])) in line
    finally:
        O.set_checkpoint(0)
