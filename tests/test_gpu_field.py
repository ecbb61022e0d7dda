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
