#!/usr/bin/env python3
"""Word-level model of mont_sqr<NL> (csrc/mp.cuh): every PTX carry-chain instruction is modelled on
32-bit words with an explicit CC flag, so the CUDA code can be transcribed from a checked algorithm.
Run: python tools/models/mont_sqr_model.py"""
import random

M32 = 0xFFFFFFFF


class Machine:
    def __init__(self):
        self.cc = 0

    def add_cc(self, a, b):
        s = a + b; self.cc = s >> 32; return s & M32

    def addc_cc(self, a, b):
        s = a + b + self.cc; self.cc = s >> 32; return s & M32

    def addc(self, a, b):
        s = a + b + self.cc
        assert s >> 32 == 0, "lost carry in addc"
        return s & M32

    def addc_wrap(self, a, b):          # addc whose carry-out is known to be irrelevant is NOT allowed here
        return self.addc(a, b)

    def mad_lo_cc(self, d, a, b):
        s = ((a * b) & M32) + d; self.cc = s >> 32; return s & M32

    def madc_lo_cc(self, d, a, b):
        s = ((a * b) & M32) + d + self.cc; self.cc = s >> 32; return s & M32

    def madc_hi_cc(self, d, a, b):
        s = ((a * b) >> 32) + d + self.cc; self.cc = s >> 32; return s & M32

    def madc_hi(self, d, a, b):
        s = ((a * b) >> 32) + d + self.cc
        assert s >> 32 == 0, "lost carry in madc.hi"
        return s & M32


def W_of(n):
    return n + 2 if n % 2 == 0 else n + 3


def mont_sqr(a, N, m0inv, n):
    m = Machine()
    W2 = 2 * n + 2
    SE = [0] * W2
    SO = [0] * W2          # word k has weight limb k+1
    # 1. off-diagonal products a_i*a_j, i<j
    for i in range(n - 1):
        for par in (0, 1):
            js = [j for j in range(i + 1, n) if (i + j) % 2 == par]
            if not js:
                continue
            acc = SE if par == 0 else SO
            first = True
            for j in js:
                k = i + j - par
                acc[k] = m.mad_lo_cc(acc[k], a[j], a[i]) if first else m.madc_lo_cc(acc[k], a[j], a[i])
                acc[k + 1] = m.madc_hi_cc(acc[k + 1], a[j], a[i])
                first = False
            acc[k + 2] = m.addc(acc[k + 2], 0)
    # 2. T = SE + SO<<32   (2n words)
    T = [0] * (2 * n)
    T[0] = SE[0]
    T[1] = m.add_cc(SE[1], SO[0])
    for k in range(2, 2 * n):
        T[k] = m.addc_cc(SE[k], SO[k - 1]) if k < 2 * n - 1 else m.addc(SE[k], SO[k - 1])
    assert SE[2 * n] == 0 and SE[2 * n + 1] == 0 and SO[2 * n - 1] == 0
    # 3. double
    assert T[2 * n - 1] >> 31 == 0
    for k in range(2 * n - 1, 0, -1):
        T[k] = ((T[k] << 1) | (T[k - 1] >> 31)) & M32
    T[0] = (T[0] << 1) & M32
    # 4. diagonal
    for i in range(n):
        T[2 * i] = m.mad_lo_cc(T[2 * i], a[i], a[i]) if i == 0 else m.madc_lo_cc(T[2 * i], a[i], a[i])
        T[2 * i + 1] = m.madc_hi_cc(T[2 * i + 1], a[i], a[i]) if i < n - 1 else m.madc_hi(T[2 * i + 1], a[i], a[i])
    assert sum(T[k] << (32 * k) for k in range(2 * n)) == sum(a[k] << (32 * k) for k in range(n)) ** 2
    # 5. windowed reduction, E/O roles as in mont_mul
    W = W_of(n)
    X = T[:n + 2] + [0] * (W - n - 2)
    Y = [0] * W
    pend = 0
    nxt = n + 2

    def mad_row(acc, x, y, par, first_carry):
        top = n + 1 if par == 0 else n
        k = None
        first = True
        for j in range(par, n, 2):
            k = j - par
            if first and not first_carry:
                acc[k] = m.mad_lo_cc(acc[k], x[j], y)
            else:
                acc[k] = m.madc_lo_cc(acc[k], x[j], y)
            acc[k + 1] = m.madc_hi_cc(acc[k + 1], x[j], y)
            first = False
        npair = (n - par + 1) // 2
        t0 = 2 * npair
        carry_out = 0
        for t in range(t0, top + 1):
            if t == top:
                if par == 0:
                    acc[t] = m.addc_cc(acc[t], 0); carry_out = m.addc(0, 0)
                else:
                    acc[t] = m.addc(acc[t], 0)
            else:
                acc[t] = m.addc_cc(acc[t], 0)
        return carry_out

    E, O = X, Y
    for i in range(n):
        if i > 0:
            # shift: (E,O) <- (O + E[1], E >> 64), then the next limb of the square enters at the top of E
            e1 = E[1]
            for k in range(W - 2):
                E[k] = E[k + 2]
            E[W - 2] = 0; E[W - 1] = 0
            E, O = O, E
            tin = T[nxt] if nxt < 2 * n else 0
            nxt += 1
            assert E[n + 1] == 0
            E[n + 1] = m.add_cc(tin, pend); pend = m.addc(0, 0)
            E[0] = m.add_cc(E[0], e1)
            mm = (E[0] * m0inv) & M32
            mad_row(O, N, mm, 1, True)
        else:
            mm = (E[0] * m0inv) & M32
            mad_row(O, N, mm, 1, False)
        pend += mad_row(E, N, mm, 0, False)
        assert E[0] == 0
    # final: T' = (E>>32) + O, n+1 words (+pend above)
    t = [0] * (n + 1)
    t[0] = m.add_cc(E[1], O[0])
    for k in range(1, n):
        t[k] = m.addc_cc(E[k + 1], O[k])
    t[n] = m.addc(E[n + 1], O[n])
    assert pend == 0, "pending carry left"
    val = sum(t[k] << (32 * k) for k in range(n + 1))
    Nv = sum(N[k] << (32 * k) for k in range(n))
    assert val < 2 * Nv
    if val >= Nv:
        val -= Nv
    return val


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
        if kind in (0, 2):
            av = Nv - 1 - rng.randrange(3)
        elif kind == 3:
            av = rng.randrange(3)
        else:
            av = rng.randrange(Nv)
        N = [(Nv >> (32 * k)) & M32 for k in range(n)]
        a = [(av >> (32 * k)) & M32 for k in range(n)]
        m0inv = (-pow(Nv, -1, 1 << 32)) & M32
        got = mont_sqr(a, N, m0inv, n)
        exp = av * av * pow(1 << (32 * n), -1, Nv) % Nv
        assert got == exp, (n, hex(Nv), hex(av))


if __name__ == "__main__":
    rng = random.Random(7)
    for n in (2, 3, 4, 5, 6, 10, 13, 16, 32):
        test(n, 600 if n <= 16 else 120, rng)
        print("n=%d ok" % n)
