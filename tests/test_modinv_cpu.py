"""Host check of the word-batched modular inverse / gcd (avx-ecm_b200/csrc/modinv_fast.hpp, the source the GPU kernels
use) against Python: inverse, gcd, the `invertible` flag, for every compiled limb count; random values, values sharing
factors with N (the inversion-failure path of stage 2), tiny and huge values, y = 0, moduli that fill their top limb or
are far shorter than the limbs, and gcd-only calls whose operand exceeds the modulus (special-form factor checks)."""
import os, random, subprocess
from math import gcd
import pytest
from conftest import ROOT

EXE = os.path.join(ROOT, "tests", "_build", "modinv_host")


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(ROOT, "tests", "modinv_host.cpp")
    hdr = os.path.join(ROOT, "avx-ecm_b200", "csrc", "modinv_fast.hpp")
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", EXE, src], check=True)

    def run(lines):
        out = subprocess.run([EXE], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout
        return [(int(a), int(b, 16), int(c, 16)) for a, b, c in (l.split() for l in out.splitlines())]
    return run


@pytest.mark.parametrize("NL", [1, 2, 3, 6, 10, 13, 16, 20, 24, 32, 48, 64])
def test_fast_inverse_matches_python(harness, NL):
    rng = random.Random(77 + NL)
    bits = 32 * NL
    full = (1 << bits) - 1
    mods = [full, full - 2 * rng.getrandbits(10), (1 << (bits - 1)) + 1, rng.getrandbits(bits) | (1 << (bits - 1)) | 1,
            rng.getrandbits(max(bits - 40, 8)) | 1 | (1 << max(bits - 41, 2)), 3, 5, (1 << min(bits, 61)) - 1,
            1000000007 * 998244353 * 4294967311 if bits >= 96 else 1000003]
    mods = [m | 1 for m in mods if 3 <= (m | 1) <= full]
    cases = []
    for n in mods:
        ys = [0, 1, 2, n - 1, n - 2, n >> 1, (n >> 1) + 1, 1 << (n.bit_length() - 2), (1 << (n.bit_length() - 1)) - 1]
        ys += [rng.randrange(n) for _ in range(25)]
        for p in (3, 5, 7, 1000000007, 998244353):
            if n % p == 0:
                ys += [p, p * rng.randrange(1, max(2, n // p)) % n, (n // p) % n]
        ys += [gcd(rng.getrandbits(bits), n) * rng.randrange(1, 1000) % n for _ in range(4)]
        cases += [(1, y % n, n) for y in ys]
        cases += [(0, rng.getrandbits(bits), n) for _ in range(6)] + [(0, 0, n), (0, n, n), (0, full, n)]
    got = harness(["%d %d %x %x" % (NL, w, y, n) for w, y, n in cases])
    assert len(got) == len(cases)
    for (w, y, n), (ok, inv, g) in zip(cases, got):
        eg = gcd(y, n) if y else n
        assert g == eg, (w, hex(y), hex(n))
        assert ok == (1 if eg == 1 else 0)
        if w and eg == 1:
            assert inv == pow(y, -1, n), (hex(y), hex(n))
