"""GPU parity of the field operators (the reference's vecmulmod/vecsqrmod/vecaddmod/vecsubmod
operator table, avx_ecm.h:205-209) against Python big integers, through the C ABI."""
import random
import pytest
from conftest import composites
import avx_ecm_b200 as E

pytestmark = pytest.mark.gpu


def moduli():
    c = composites()
    return {
        "small96": c["small96"], "t35_297b": c["t35"], "syn415": c["syn415"], "readme508": c["readme508"],
        "syn1024_fullwidth": c["syn1024"], "syn2048_fullwidth": c["syn2048"], "odd_700b": (1 << 700) - 1 - (1 << 350), "allones_416": (1 << 416) - 1 - 2 * 0,   # odd, every limb 0xffffffff
        "tiny_3limb": (1 << 65) + 13,
    }


@pytest.mark.parametrize("name", sorted(moduli()))
def test_field_ops_match_python(name):
    N = moduli()[name]
    if N % 2 == 0:
        N += 1
    rng = random.Random(1234)
    n = 1000
    a = [rng.randrange(N) for _ in range(n)]
    b = [rng.randrange(N) for _ in range(n)]
    # edge operands: 0, 1, N-1, values around 2^32k
    edge = [0, 1, N - 1, N - 2, (1 << 32) % N, ((1 << 64) - 1) % N, N >> 1]
    for i, e in enumerate(edge):
        a[i] = e
        b[i] = edge[(i * 3 + 1) % len(edge)]
    a[10], b[10] = N - 1, N - 1
    a[11], b[11] = 0, 0
    ctx = E.EcmContext(N, n)
    try:
        assert ctx.fieldop(0, a, b) == [x * y % N for x, y in zip(a, b)]
        assert ctx.fieldop(1, a, b) == [x * x % N for x in a]
        assert ctx.fieldop(2, a, b) == [(x + y) % N for x, y in zip(a, b)]
        assert ctx.fieldop(3, a, b) == [(x - y) % N for x, y in zip(a, b)]
        # chained: a <- a*b, 5 times
        assert ctx.fieldop(0, a, b, repeat=5) == [x * pow(y, 5, N) % N for x, y in zip(a, b)]
    finally:
        ctx.close()


def test_rejects_even_modulus_and_oversize():
    with pytest.raises(E.EcmError):
        E.EcmContext(1 << 200, 8)
    with pytest.raises(E.EcmError):
        E.EcmContext((1 << 2100) + 1, 8)      # larger than the widest compiled kernel (64 limbs)


SPECIAL_BASES = [  # (k, kind, c): 2^k - c (kind 1) or 2^k + 1 (kind -1)
    (64, 1, 59), (89, 1, 1), (95, -1, 1), (127, 1, 1), (160, 1, 47), (191, -1, 1), (220, 1, 69), (277, 1, 1),
    (288, -1, 1), (319, 1, 1), (320, 1, (1 << 31) - 1), (415, 1, 1), (416, -1, 1), (511, 1, 1), (523, -1, 1), (607, 1, 1),
    (640, 1, 3), (767, -1, 1), (768, 1, 1), (800, -1, 1), (1000, 1, 1), (1023, 1, 1), (1023, -1, 1)]


@pytest.mark.parametrize("k,kind,c", SPECIAL_BASES)
def test_special_base_field_ops_match_python(k, kind, c, monkeypatch):
    """Shift-and-fold kernels (create_special with a base of the reference's special shapes) against Python
    integers, and against the Montgomery kernels on the same base (ECM_B200_NO_FOLD)."""
    M = (1 << k) - c if kind > 0 else (1 << k) + 1
    rng = random.Random(k)
    n = 600
    a = [rng.randrange(M) for _ in range(n)]
    b = [rng.randrange(M) for _ in range(n)]
    edge = [0, 1, M - 1, M - 2, (1 << (k - 1)), (1 << (k - 1)) + 1, (1 << 32) % M, M >> 1, (1 << k) % M, ((1 << k) - 1) % M]
    for i, e in enumerate(edge):
        a[i] = e
        b[i] = edge[(i * 3 + 1) % len(edge)]
    a[20], b[20] = M - 1, M - 1
    ctx = E.EcmContext(M, n, base=M)
    try:
        assert ctx.uses_fold
        assert ctx.fieldop(0, a, b) == [x * y % M for x, y in zip(a, b)]
        assert ctx.fieldop(1, a, b) == [x * x % M for x in a]
        assert ctx.fieldop(2, a, b) == [(x + y) % M for x, y in zip(a, b)]
        assert ctx.fieldop(3, a, b) == [(x - y) % M for x, y in zip(a, b)]
        assert ctx.fieldop(0, a, b, repeat=5) == [x * pow(y, 5, M) % M for x, y in zip(a, b)]
    finally:
        ctx.close()


def test_fold_and_montgomery_kernels_agree_end_to_end(monkeypatch):
    base = (1 << 523) + 1
    N = base // 3
    fold = E.vececm(N, 40, 2000, 120000, sigma=555000111, base=base)
    monkeypatch.setenv("ECM_B200_NO_FOLD", "1")
    ctx = E.EcmContext(N, 40, base=base)
    try:
        assert not ctx.uses_fold
        mont = E.vececm(N, 40, 2000, 120000, sigma=555000111, ctx=ctx)
    finally:
        ctx.close()
    for key in ("save_lines", "factors", "acc", "inv_fail"):
        assert fold[key] == mont[key]


def test_base_outside_the_fold_range_uses_montgomery():
    # 2^1100-1 needs 35 limbs: beyond the 1024-bit fold kernels; and a base that is not of a special shape
    for base in ((1 << 1100) - 1, 11 * ((1 << 200) + 235)):
        ctx = E.EcmContext(base, 8, base=base)
        try:
            assert not ctx.uses_fold
            assert ctx.fieldop(0, [base - 1, 5], [base - 1, 7]) == [1, 35]
        finally:
            ctx.close()
