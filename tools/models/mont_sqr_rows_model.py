#!/usr/bin/env python3
"""Word-level model of the row-interleaved Montgomery squaring (mont_sqr_rows in csrc/mp.cuh).

Same CIOS row structure as mont_mul, but row i only multiplies a_i by the limbs j >= i of the
pre-doubled operand:   a^2 = sum_i 2^(32 i) * a_i * c^(i),
   c^(i)_i = a_i,  c^(i)_(i+1) = (a_(i+1) << 1) mod 2^32,  c^(i)_j = (2a)_j for j >= i+2 (incl. j = n)
so the products per squaring drop from n^2 to n(n+1)/2 + n.  The carry of the "E' = O + e1" add is not
rippled through the accumulator: it waits in a small pending counter p (weight: limb 1) and joins the same
add of the next row, and the final merge.
Run: python tools/models/mont_sqr_rows_model.py"""
import random
from mont_sqr_model import Machine, M32, W_of


def mont_sqr_rows(a, N, m0inv, n):
    m = Machine()
    W = W_of(n)
    X, Y = [0] * W, [0] * W
    a2 = [((a[j] << 1) | (a[j - 1] >> 31 if j else 0)) & M32 for j in range(n)] + [a[n - 1] >> 31]

    def c(i, j):
        if j == i:
            return a[i]
        if j == i + 1:
            return (a[j] << 1) & M32 if j < n else a2[n] * 0 + (0)  # j == n only when i == n-1: (a_n << 1) = 0
        return a2[j]

    def mad_row_from(acc, vec, y, par, j0, jmax):
        """acc += vec(j)*y for j in [j0, jmax] with j % 2 == par; words above the last pair catch the carry"""
        top = n + 1 if par == 0 else n
        first = True
        k = None
        for j in range(par, jmax + 1, 2):
            if j < j0:
                continue
            k = j - par
            acc[k] = m.mad_lo_cc(acc[k], vec(j), y) if first else m.madc_lo_cc(acc[k], vec(j), y)
            if k + 1 == top and False:
                pass
            acc[k + 1] = m.madc_hi_cc(acc[k + 1], vec(j), y)
            first = False
        if k is None:
            return
        if k + 1 >= top:
            assert m.cc == 0, "carry out of the top word"
            return
        for t in range(k + 2, top + 1):
            acc[t] = m.addc(acc[t], 0) if t == top else m.addc_cc(acc[t], 0)

    E, O = X, Y
    p = 0
    for i in range(n):
        # shift: (E,O) <- (O + e1 + p, E >> 64); carries of the 3-way add wait in p
        e1 = E[1]
        for k in range(W - 2):
            E[k] = E[k + 2]
        E[W - 2] = 0; E[W - 1] = 0
        E, O = O, E
        t = m.add_cc(E[0], e1); c1 = m.addc(0, 0)
        t = m.add_cc(t, p); c2 = m.addc(0, 0)
        E[0] = t; p = c1 + c2
        ai = a[i]
        mad_row_from(O, lambda j: c(i, j), ai, 1, i, n)
        mad_row_from(E, lambda j: c(i, j), ai, 0, i, n)
        mm = (E[0] * m0inv) & M32
        mad_row_from(O, lambda j: N[j], mm, 1, 0, n - 1)
        mad_row_from(E, lambda j: N[j], mm, 0, 0, n - 1)
        assert E[0] == 0
    t = [0] * (n + 1)
    t[0] = m.add_cc(E[1], O[0]); cA = m.addc(0, 0)
    t[0] = m.add_cc(t[0], p); cB = m.addc(0, 0)
    carry = cA + cB
    t[1] = m.add_cc(E[2], O[1]); c1 = m.addc(0, 0)
    t[1] = m.add_cc(t[1], carry); c2 = m.addc(0, 0)
    m.cc = 0
    carry = c1 + c2
    assert carry <= 1
    m.cc = carry
    for k in range(2, n):
        t[k] = m.addc_cc(E[k + 1], O[k])
    t[n] = m.addc(E[n + 1], O[n])
    val = sum(t[k] << (32 * k) for k in range(n + 1))
    Nv = sum(N[k] << (32 * k) for k in range(n))
    assert val < 2 * Nv
    return val - Nv if val >= Nv else val


def test(n, trials, rng):
    for tr in range(trials):
        kind = tr % 6
        if kind == 0:
            Nv = (1 << (32 * n)) - 1 - 2 * rng.randrange(1000)
        elif kind == 1:
            Nv = (1 << (32 * n - 1)) + 1 + 2 * rng.randrange(1 << 20)
        else:
            Nv = rng.getrandbits(32 * n - rng.randrange(0, 40)) | 1
        if Nv < 3:
            Nv = 3
        av = Nv - 1 - rng.randrange(3) if kind in (0, 2) else (rng.randrange(3) if kind == 3 else rng.randrange(Nv))
        N = [(Nv >> (32 * k)) & M32 for k in range(n)]
        a = [(av >> (32 * k)) & M32 for k in range(n)]
        m0inv = (-pow(Nv, -1, 1 << 32)) & M32
        got = mont_sqr_rows(a, N, m0inv, n)
        exp = av * av * pow(1 << (32 * n), -1, Nv) % Nv
        assert got == exp, (n, hex(Nv), hex(av))


if __name__ == "__main__":
    rng = random.Random(5)
    for n in (2, 3, 4, 5, 6, 10, 13, 16, 20, 32):
        test(n, 600 if n <= 16 else 150, rng)
        print("n=%d ok" % n)
