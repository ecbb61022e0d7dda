"""Operator-level pin of the oracle's special-form claim against the UNMODIFIED reference operators
(oracle/_ref/fieldop-ref, built by oracle/build_ref.sh from the reference sources where they lie):

  * for random operands vecmulmod52_mersenne / vecsqrmod52_mersenne / vecaddmod52_mersenne /
    vecsubmod52_mersenne / vec_simul_addsub52_mersenne return the canonical residue mod 2^k-1, 2^k+1, 2^k-c
    -- the arithmetic the oracle (and the GPU engine) use for special-form inputs;
  * the one case where they do not: the carry helpers of vecarith52.c:102-116 raise a carry-out whenever a
    word is all-ones (borrow: zero) even without a carry-in, so a lane whose residue contains such words is
    corrupted when ANOTHER lane keeps the carry loop running.  Reproduced here with the very operands of
    sigma = 10 on (2^523+1)/3; this is why the special golden vectors use random-looking sigmas.
Skipped when the reference binary is absent (it cannot be built without AVX-512)."""
import os, random, subprocess
import pytest
from conftest import ROOT
import oracle_lib as O

HARNESS = os.path.join(ROOT, "oracle", "_ref", "fieldop-ref")
pytestmark = pytest.mark.skipif(not os.path.exists(HARNESS), reason="oracle/_ref/fieldop-ref not built")


def run(k, c, script):
    out = subprocess.run([HARNESS, str(k), str(c)], input=script, capture_output=True, text=True, check=True).stdout
    return [[int(v, 16) for v in line.split()] for line in out.splitlines()]


@pytest.mark.parametrize("k,c", [(277, 1), (523, -1), (127, -1), (220, 69), (220, 57), (607, 1), (1024, -1), (416, 3)])
def test_reference_special_ops_are_canonical_residues(k, c):
    M = (1 << k) - c if c > 0 else (1 << k) + 1
    rng = random.Random(k * 1000 + c)
    ops, exp = [], []
    for _ in range(60):
        a, b = rng.randrange(M), rng.randrange(M)
        if a >> k or b >> k:                      # operands are k-bit words in the reference
            continue
        ops += ["mul %x %x" % (a, b), "sqr %x" % a, "add %x %x" % (a, b), "sub %x %x" % (a, b), "addsub %x %x" % (a, b)]
        exp += [[a * b % M], [a * a % M], [(a + b) % M], [(a - b) % M], [(a + b) % M, (a - b) % M]]
    got = run(k, c, "\n".join(ops) + "\n")
    assert len(got) == len(exp)
    bad = [(o, g, e) for o, g, e in zip(ops, got, exp) if [v % M for v in g] != e]
    assert not bad, bad[:2]
    # and the representatives themselves are the canonical ones (< M), not merely congruent
    assert all(v < M for g in got for v in g)


def test_reference_carry_helper_quirk_is_lane_coupled():
    M = 2 ** 523 + 1

    def dup(X, Z, s):
        d, sm = (X - Z) % M, (X + Z) % M
        t1, t2 = d * d % M, sm * sm % M
        t3 = (t2 - t1) % M
        return t1 * t2 % M, (t3 * s + t1) * t3 % M

    X, s = O.build_curve(M, 10)
    alone = "dup " + " ".join(["%x 1 %x" % (X, s)] + ["0 0 0"] * 7) + "\n"
    assert run(523, -1, alone)[0] == list(dup(X, 1, s))                 # fine on its own
    lanes = [O.build_curve(M, 10 + i) for i in range(8)]
    together = "dup " + " ".join("%x 1 %x" % xs for xs in lanes) + "\n"
    got = run(523, -1, together)[0]
    assert got[0] == dup(X, 1, s)[0] and got[1] != dup(X, 1, s)[1]      # Z of lane 0 corrupted by its neighbours
