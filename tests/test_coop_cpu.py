"""Host check of the warp-cooperative field arithmetic (avx-ecm_b200/csrc/coop.cuh): the very source the GPU kernels
use, compiled for the CPU (tests/coop_host.cpp: the L lanes of a group run as L lock-stepped threads, shuffles and votes
through a barrier) and compared with Python integers: Montgomery product, modular sum and difference, canonical results,
for every lane split the kernels use and a few odd ones, random and adversarial operands (all-ones / all-zero limb runs
that make carries and borrows cross every lane boundary, operands next to N, moduli that fill their top limb)."""
import os, random, subprocess
import pytest
from conftest import ROOT

EXE = os.path.join(ROOT, "tests", "_build", "coop_host")


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(ROOT, "tests", "coop_host.cpp")
    hdr = os.path.join(ROOT, "avx-ecm_b200", "csrc", "coop.cuh")
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-o", EXE, src], check=True)

    def run(lines):
        out = subprocess.run([EXE], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout
        return [int(x, 16) for x in out.split()]
    return run


def operands(rng, N, bits):
    full = (1 << bits) - 1
    special = [0, 1, N - 1, N - 2, N >> 1, (N >> 1) + 1]
    for w in (32, 64, bits // 4, bits // 2):                         # limb runs of ones / zeros inside the value
        special += [full >> w, (full >> w) << (w // 2), full ^ (full >> w), (1 << (bits - w)) - 1, 1 << (bits - w)]
    special = [v % N for v in special]
    out = [(rng.choice(special), rng.choice(special)) for _ in range(14)]
    out += [(rng.randrange(N), rng.choice(special)) for _ in range(6)]
    out += [(rng.randrange(N), rng.randrange(N)) for _ in range(12)]
    out += [(a, a) for a, _ in out[:4]] + [(N - 1, N - 1), (0, 0), (N - 1, 1), (1, N - 1)]
    return out


def moduli(rng, bits):
    full = (1 << bits) - 1
    return [full,                                                    # all ones: every slice of N is all ones
            full - 2 * rng.getrandbits(20),                          # fills the top limb
            (1 << (bits - 1)) + 1,                                   # 1 0 0 ... 0 1
            rng.getrandbits(bits) | (1 << (bits - 1)) | 1,
            rng.getrandbits(bits - 17) | (1 << (bits - 18)) | 1,     # shorter than the limbs: zero top words
            (1 << (bits - 31)) - 1]


@pytest.mark.parametrize("M,L", [(16, 4), (14, 4), (12, 4), (10, 4), (16, 2), (8, 4), (2, 2), (4, 8), (2, 4), (10, 2), (16, 8)])
def test_coop_field_ops_match_python(harness, M, L):
    rng = random.Random(1000 * M + L)
    bits = 32 * M * L
    R = 1 << bits
    lines, exp = [], []
    for N in moduli(rng, bits):
        Rinv = pow(R, -1, N)
        for a, b in operands(rng, N, bits):
            for op, v in (("mul", a * b * Rinv % N), ("add", (a + b) % N), ("sub", (a - b) % N)):
                lines.append("%s %d %d %x %x %x" % (op, L, M, a, b, N))
                exp.append(v)
    got = harness(lines)
    assert len(got) == len(exp)
    bad = [(l, "%x" % g, "%x" % e) for l, g, e in zip(lines, got, exp) if g != e]
    assert not bad, (len(bad), bad[0])
