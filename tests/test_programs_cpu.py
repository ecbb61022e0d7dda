"""CPU-side check of the host compilers: the stage-1 macro-op stream and the stage-2 instruction stream
are interpreted here with Python integers (plain arithmetic mod N, no Montgomery form) and must reproduce
the oracle's residues.  This validates plan.cpp / plan2.cpp and the micro-programs' data flow without a GPU;
the device kernels execute exactly these streams."""
import pytest
from conftest import composites
import oracle_lib as O
import avx_ecm_b200 as E

PERM = [0xE4, 0xB4, 0xD8, 0x78, 0x9C, 0x6C, 0xE1, 0xB1, 0xC9, 0x39, 0x8D, 0x2D,
        0xD2, 0x72, 0xC6, 0x36, 0x4E, 0x1E, 0x93, 0x63, 0x87, 0x27, 0x4B, 0x1B]


def run_stage1_stream(N, x, s, ops):
    """Interpret vm.cuh's macro-ops.  Point slot p holds (X,Z); P starts in slot 0."""
    pts = {0: (x, 1), 1: None, 2: None, 3: None}

    def add(s1, d1, s2, d2, pin):          # vec_add, ecm.c:407-443
        t1, t2 = d1 * s2 % N, s1 * d2 % N
        return (pow(t1 + t2, 2, N) * pin[1] % N, pow(t1 - t2, 2, N) * pin[0] % N)

    def dup(sm, df):                       # vec_duplicate, ecm.c:445-457
        t1, t2 = df * df % N, sm * sm % N
        t3 = (t2 - t1) % N
        return (t1 * t2 % N, (t3 * s + t1) % N * t3 % N)

    sums = lambda p: ((p[0] + p[1]) % N, (p[0] - p[1]) % N)
    last_T = 0
    for b in ops:
        t, p = b & 7, PERM[b >> 3]
        A, B, C, T = p & 3, (p >> 2) & 3, (p >> 4) & 3, (p >> 6) & 3
        if t == 0:                                         # M_DBL
            pts[T] = dup(*sums(pts[T])); last_T = T
        elif t == 1:                                       # M_INIT: C = B = P, A = 2P
            pts[C] = pts[B]; pts[A] = dup(*sums(pts[B]))
        elif t == 2:                                       # M_C3: T = B + A (C)
            (s1, d1), (s2, d2) = sums(pts[B]), sums(pts[A]); pts[T] = add(s1, d1, s2, d2, pts[C])
        elif t == 3:                                       # M_C4: B = B + A (C); A = 2A
            (s1, d1), (s2, d2) = sums(pts[B]), sums(pts[A]); pts[B] = add(s1, d1, s2, d2, pts[C]); pts[A] = dup(s2, d2)
        elif t == 4:                                       # M_C5: C = C + A (B); A = 2A
            (s1, d1), (s2, d2) = sums(pts[C]), sums(pts[A]); pts[C] = add(s1, d1, s2, d2, pts[B]); pts[A] = dup(s2, d2)
        elif t == 5:                                       # M_C9: C = C + B (A); B = 2B
            (s1, d1), (s2, d2) = sums(pts[C]), sums(pts[B]); pts[C] = add(s1, d1, s2, d2, pts[A]); pts[B] = dup(s2, d2)
        elif t == 6:                                       # M_FINAL: T = A + B (C)
            (s1, d1), (s2, d2) = sums(pts[A]), sums(pts[B]); pts[T] = add(s1, d1, s2, d2, pts[C]); last_T = T
    return pts[last_T]


@pytest.mark.parametrize("name,b1,sigma", [("syn415", 3000, 7), ("t35", 1200, 2 ** 63 + 5), ("syn1024", 500, 12)])
def test_stage1_stream_reproduces_oracle_residues(name, b1, sigma):
    N = composites()[name]
    x, s = O.build_curve(N, sigma)
    ops, _, _ = E.plan_stage1(b1)
    X, Z = run_stage1_stream(N, x, s, ops)
    o = O.ecm_curve(N, b1, b1, sigma)
    assert (X, Z) == (o["x"], o["z"])


def run_stage1_phases(N, x, s, ops):
    """Interpret the PHASE programs of the register-resident stage-1 kernel (rv.cuh) exactly as k_stage1_rv does: four
    operand registers A0, A1, B0, B1, a two-value park, point slots addressed through the macro-op's permutation."""
    progs = [E.rv_program(t) for t in range(8)]
    pts = {0: [x, 1], 1: [0, 0], 2: [0, 0], 3: [0, 0]}
    A0 = A1 = B0 = B1 = 0
    park = [0, 0]
    last_T = 0
    for b in ops:
        t, p = b & 7, PERM[b >> 3]
        phys = lambda sym: (p >> (2 * sym)) & 3
        for u in progs[t]:
            kind, px, py, flag = u & 15, phys((u >> 4) & 3), phys((u >> 8) & 3), (u >> 16) & 1
            if kind == 0:                                      # A1
                X, Z = pts[px]; A0, A1 = (X - Z) % N, (X + Z) % N
                X, Z = pts[py]; B0, B1 = (X + Z) % N, (X - Z) % N
                if flag:
                    park = [B0, B1]
            elif kind == 1:                                    # A2
                A0, A1 = (A0 + A1) % N, (A0 - A1) % N
                B0, B1 = A0, A1
            elif kind == 2:                                    # A3
                B0, B1 = pts[px][1], pts[px][0]
            elif kind in (3, 4):                               # D1L / D1P
                if kind == 3:
                    X, Z = pts[px]; A0, A1 = (X + Z) % N, (X - Z) % N
                else:
                    A0, A1 = park
                B0, B1 = A0, A1
            elif kind == 5:                                    # D2
                B0, B1, A1 = A1, (A0 - A1) % N, s
            elif kind == 6:                                    # D3
                pts[px][0] = A0
                A0 = (A1 + B0) % N
            elif kind == 7:                                    # COPY
                pts[py] = list(pts[px])
                continue
            if kind == 6:
                A0 = A0 * B1 % N
                pts[px][1] = A0
            else:
                A0, A1 = A0 * B0 % N, A1 * B1 % N
                if kind == 2:
                    pts[py] = [A0, A1]
        if t in (0, 6):
            last_T = (p >> 6) & 3
    return tuple(pts[last_T])


@pytest.mark.parametrize("name,b1,sigma", [("syn415", 3000, 7), ("t35", 1200, 2 ** 63 + 5), ("syn2048", 300, 12)])
def test_rv_phase_programs_reproduce_oracle_residues(name, b1, sigma):
    N = composites()[name]
    x, s = O.build_curve(N, sigma)
    ops, _, _ = E.plan_stage1(b1)
    X, Z = run_stage1_phases(N, x, s, ops)
    o = O.ecm_curve(N, b1, b1, sigma)
    assert (X, Z) == (o["x"], o["z"])


def run_stage2_program(N, code, slots, tab):
    UX, ACC, S1, SP, T1 = 0, 6, 7, 11, 12
    for ins in code:
        lo, imm = ins & 0xFFFFFFFF, ins >> 32
        op, d, x, y = lo & 0xFF, (lo >> 8) & 0xFF, (lo >> 16) & 0xFF, lo >> 24
        if op in (0, 1):
            slots[d] = slots[x] * slots[y] % N
        elif op == 2:
            slots[d] = (slots[x] + slots[y]) % N
        elif op == 3:
            slots[d] = (slots[x] - slots[y]) % N
        elif op == 4:
            a, b = slots[x], slots[y]
            slots[d], slots[imm] = (a + b) % N, (a - b) % N
        elif op == 5:
            slots[d] = slots[x]
        elif op == 6:
            slots[d] = tab[imm]
        elif op == 7:
            tab[imm] = slots[x]
        elif op == 8:
            slots[d] = pow(slots[x], -1, N)
        elif op == 9:
            slots[d] = 1
        elif op == 10:
            slots[ACC] = slots[ACC] * (tab[imm & 0xFFFF] - tab[imm >> 16]) % N
        elif op == 12:
            d4, x4, y4, e4, u4, v4 = [(lo >> s) & 15 for s in (8, 12, 16, 20, 24, 28)]
            r0, r1 = slots[x4] * slots[y4] % N, slots[u4] * slots[v4] % N
            slots[d4], slots[e4] = r0, r1
        elif op == 11:
            pass
        else:
            raise AssertionError("unknown op %d" % op)


@pytest.mark.parametrize("name,b1,b2,sigma", [("syn415", 300, 20000, 7), ("t35", 5000, 120000, 99), ("syn415", 50, 3000, 11),
                                              # every D the planner selects (30 ... 2310), several window shifts each
                                              ("syn206", 100, 10000, 13), ("syn206", 200, 20000, 14), ("syn206", 400, 40000, 15),
                                              ("syn206", 1500, 150000, 16), ("syn206", 3000, 300000, 17), ("syn206", 20000, 600000, 18)])
def test_stage2_program_reproduces_oracle_accumulator(name, b1, b2, sigma):
    N = composites()[name]
    o = O.ecm_curve(N, b1, b2, sigma)
    assert not o["counters"][6]                  # no inversion failure in these cases
    init, lay = E.stage2_program(b1, b2, -1)
    slots = [0] * 14
    _, s = O.build_curve(N, sigma)
    slots[11] = s                                # SP
    tab = {lay["qx"]: o["x"], lay["qz"]: o["z"]} # Q = stage-1 result
    run_stage2_program(N, init, slots, tab)
    r = 0
    while True:
        code, _ = E.stage2_program(b1, b2, r)
        if not code:
            break
        run_stage2_program(N, code, slots, tab)
        r += 1
    assert slots[6] == o["acc"]
