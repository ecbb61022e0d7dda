#!/usr/bin/env python3
"""Word-level model of the two-thread cooperative Montgomery multiply (wide moduli, NL = 2H limbs).

Thread A owns the low H limbs of a and N, thread B the high H limbs.  Each keeps a private running sum in
the same E/O form as mont_mul<H> (V = E + O*2^32, H+2..H+3 words), the global sum being
T = V_A + V_B * 2^(32H) in *redundant* form: V_A may exceed 2^(32H), so nothing has to flow from A to B
inside the loop except the quotient digit m.  The division by 2^32 after each row moves B's lowest word
(E_B[0]) to A at limb H-1.  Only at the very end A's overflow words are handed to B and the conditional
subtraction runs across the two threads.  Per row: two warp shuffles (m, E_B[0]).
Run: python tools/models/mont_coop_model.py"""
import random
from mont_sqr_model import Machine, M32, W_of


class Half:
    """one thread: private E/O accumulators of W words over H limbs"""

    def __init__(self, H):
        self.H, self.W = H, W_of(H)
        self.X, self.Y = [0] * self.W, [0] * self.W
        self.E, self.O = self.X, self.Y
        self.m = Machine()

    def mad_row(self, acc, x, y, par, first_carry):
        m, H = self.m, self.H
        top = H + 1 if par == 0 else H
        first = True
        for j in range(par, H, 2):
            k = j - par
            acc[k] = m.madc_lo_cc(acc[k], x[j], y) if (first_carry or not first) else m.mad_lo_cc(acc[k], x[j], y)
            acc[k + 1] = m.madc_hi_cc(acc[k + 1], x[j], y)
            first = False
        t0 = 2 * ((H - par + 1) // 2)
        for t in range(t0, top + 1):
            acc[t] = m.addc(acc[t], 0) if t == top else m.addc_cc(acc[t], 0)

    def shift(self):
        """(E,O) <- (O + E[1], E>>64); returns the dropped low word E[0]; leaves CC = carry of the e1 add"""
        E, O, W = self.E, self.O, self.W
        low, e1 = E[0], E[1]
        for k in range(W - 2):
            E[k] = E[k + 2]
        E[W - 2] = 0; E[W - 1] = 0
        self.E, self.O = O, E
        self.E[0] = self.m.add_cc(self.E[0], e1)
        return low

    def add_word(self, limb, w):
        """add a 32-bit word at local limb `limb` (into the E role array), carries ripple to the top"""
        m, E = self.m, self.E
        E[limb] = m.add_cc(E[limb], w)
        for t in range(limb + 1, self.H + 2):
            E[t] = m.addc(E[t], 0) if t == self.H + 1 else m.addc_cc(E[t], 0)


def mont_mul_coop(a, b, N, m0inv, n):
    H = n // 2
    A, B = Half(H), Half(H)
    a_lo, a_hi, N_lo, N_hi = a[:H], a[H:], N[:H], N[H:]
    for i in range(n):
        if i > 0:
            # divide by 2^32: B's lowest word moves to A's limb H-1
            # (B first so that its value is available; order inside a thread matters for the CC chain only)
            wB = B.shift()
            B.mad_row(B.O, a_hi, b[i], 1, True)
            B.mad_row(B.E, a_hi, b[i], 0, False)
            low = A.shift()
            assert low == 0
            A.mad_row(A.O, a_lo, b[i], 1, True)
            A.add_word(H - 1, wB)
            A.mad_row(A.E, a_lo, b[i], 0, False)
        else:
            A.mad_row(A.O, a_lo, b[i], 1, False); A.mad_row(A.E, a_lo, b[i], 0, False)
            B.mad_row(B.O, a_hi, b[i], 1, False); B.mad_row(B.E, a_hi, b[i], 0, False)
        mm = (A.E[0] * m0inv) & M32                      # shuffled A -> B
        A.mad_row(A.O, N_lo, mm, 1, False); A.mad_row(A.E, N_lo, mm, 0, False)
        B.mad_row(B.O, N_hi, mm, 1, False); B.mad_row(B.E, N_hi, mm, 0, False)
        assert A.E[0] == 0
    # pending last division by 2^32 and merge of each thread's E/O into plain words
    def merge(T):
        m = T.m
        t = [0] * (H + 2)
        t[0] = m.add_cc(T.E[1], T.O[0])
        for k in range(1, H + 1):
            t[k] = m.addc_cc(T.E[k + 1], T.O[k]) if k < H + 1 else 0
        t[H + 1] = m.addc(0, 0) if False else 0
        return t
    # do it explicitly with python ints to keep the model simple, but assert the word ranges
    wB = B.E[0]
    vB = ((sum(B.E[k] << (32 * k) for k in range(B.W)) - wB) >> 32) + sum(B.O[k] << (32 * k) for k in range(B.W))
    vA = (sum(A.E[k] << (32 * k) for k in range(A.W)) >> 32) + sum(A.O[k] << (32 * k) for k in range(A.W)) + (wB << (32 * (H - 1)))
    assert vA < 1 << (32 * (H + 2)) and vB < 1 << (32 * (H + 2))
    # A hands its overflow words (limbs >= H) to B
    over = vA >> (32 * H)
    assert over < 1 << 64
    vA &= (1 << (32 * H)) - 1
    vB += over
    T = vA + (vB << (32 * H))
    Nv = sum(N[k] << (32 * k) for k in range(n))
    assert T < 2 * Nv
    return T - Nv if T >= Nv else T


def test(n, trials, rng):
    for tr in range(trials):
        kind = tr % 5
        if kind == 0:
            Nv = (1 << (32 * n)) - 1 - 2 * rng.randrange(1000)
        elif kind == 1:
            Nv = (1 << (32 * n - 1)) + 1 + 2 * rng.randrange(1 << 20)
        else:
            Nv = rng.getrandbits(32 * n - rng.randrange(0, 40)) | 1
        av = Nv - 1 - rng.randrange(3) if kind in (0, 2) else rng.randrange(Nv)
        bv = Nv - 1 - rng.randrange(3) if kind in (0, 3) else rng.randrange(Nv)
        N = [(Nv >> (32 * k)) & M32 for k in range(n)]
        a = [(av >> (32 * k)) & M32 for k in range(n)]
        b = [(bv >> (32 * k)) & M32 for k in range(n)]
        m0inv = (-pow(Nv, -1, 1 << 32)) & M32
        got = mont_mul_coop(a, b, N, m0inv, n)
        exp = av * bv * pow(1 << (32 * n), -1, Nv) % Nv
        assert got == exp, (n, hex(Nv), hex(av), hex(bv))


if __name__ == "__main__":
    rng = random.Random(11)
    for n in (4, 6, 8, 10, 12, 48, 64):
        test(n, 400 if n <= 12 else 60, rng)
        print("n=%d ok" % n)
