"""csrc/special_fold.hpp (the shift-and-fold reduction of the special-form kernels) compiled for the host and
checked limb for limb against Python integers: every form (2^k-1, 2^k-c, 2^k+1), every word/bit position of k
that a limb count serves, random and extreme operands."""
import ctypes, os, random, subprocess
import pytest
from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "special_fold_host.cpp")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("fold") / "fold.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, SRC], check=True)
    L = ctypes.CDLL(so)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    L.special_fold_host.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, u32p, u32p, u32p]
    return L


def limbs(v, n):
    return (ctypes.c_uint32 * n)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def fold(L, nl, k, kind, c, T):
    M = (1 << k) - c if kind > 0 else (1 << k) + 1
    r = (ctypes.c_uint32 * nl)()
    rc = L.special_fold_host(nl, k, kind, c, limbs(T, 2 * nl), limbs(M, nl), r)
    assert rc == 0, (nl, k, rc)
    return sum(r[i] << (32 * i) for i in range(nl)), M


@pytest.mark.parametrize("nl", [3, 6, 10, 13, 16, 20, 24, 32])
def test_fold_matches_python(lib, nl):
    rng = random.Random(nl)
    low = nl - 9 if nl > 10 else 2
    ks = set()
    for w in range(low, nl):
        ks.update(32 * w + s for s in (0, 1, 5, 16, 31))
    ks.update(rng.randrange(32 * low, 32 * nl) for _ in range(20))
    for k in sorted(ks):
        for kind, c in ((1, 1), (1, 3), (1, 69), (1, (1 << 31) - 1), (-1, 1)):
            M = (1 << k) - c if kind > 0 else (1 << k) + 1
            ops = [(M - 1, M - 1), (0, 0), (1, M - 1), (M - 1, 2), ((1 << (k - 1)), (1 << (k - 1)) + 1)]
            ops += [(rng.randrange(M), rng.randrange(M)) for _ in range(12)]
            ops += [((1 << k) - 1 if kind > 0 and c > 1 else M - 1, rng.randrange(M))]
            for a, b in ops:
                a %= M
                b %= M
                got, _ = fold(lib, nl, k, kind, c, a * b)
                assert got == a * b % M, (nl, k, kind, c, hex(a), hex(b))


def test_unsupported_positions_are_rejected(lib):
    T = (ctypes.c_uint32 * 64)()
    n = (ctypes.c_uint32 * 32)()
    r = (ctypes.c_uint32 * 32)()
    assert lib.special_fold_host(32, 32 * 22 + 3, 1, 1, T, n, r) == -2      # below the word range this limb count serves
    assert lib.special_fold_host(13, 32 * 13, 1, 1, T, n, r) == -2          # bit k outside the limbs
