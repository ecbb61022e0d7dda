"""CPU-side checks of the product's host logic: the C-ABI library loads and exports every symbol
declared in include/ecm_b200.h, and the host planners (PRAC op stream, PAIR, stage-2 geometry)
agree with the oracle / the reference's printed counters.  No GPU compute is called."""
import ctypes, os, re
import pytest
from conftest import GOLDEN, ROOT
import oracle_lib as O
import avx_ecm_b200 as E

TYPE_CH = {0: "D", 1: "I", 2: "3", 3: "4", 4: "5", 5: "9", 6: "F"}


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ecm_b200.h")).read()
    declared = set(re.findall(r"\b(ecm_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(E.EXPORTS)
    L = ctypes.CDLL(E.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.EcmError, match="no usable CUDA device"):
        E.EcmContext((1 << 127) - 1, 8)


@pytest.mark.parametrize("b1", [50, 400, 3000, 20000, 100000])
def test_stage1_plan_matches_oracle_trace(b1):
    ops, adds, dups = E.plan_stage1(b1)
    mine = "".join(TYPE_CH[b & 7] for b in ops)
    ref = O.stage1_trace(b1).decode().replace("S", "")
    assert mine == ref
    assert adds == sum(ref.count(c) for c in "3459F")
    assert dups == sum(ref.count(c) for c in "DI459")


def test_stage1_plan_counts_match_reference_printout():
    # ecm.c:1849 prints these for B1=1e6; golden files hold them too
    _, adds, dups = E.plan_stage1(1000000)
    assert (adds, dups) == (1980817, 217929)
    g = GOLDEN["csh250k_stage1_factor"]["counts"]
    _, adds, dups = E.plan_stage1(250000)
    assert (adds, dups) == (g["s1_ptadds"], g["s1_ptdups"])


def test_stage1_plan_slot_permutations_are_consistent():
    # replay the permutation bookkeeping: every op must read slots that hold live points
    ops, _, _ = E.plan_stage1(5000)
    perm_tab = [0xE4, 0xB4, 0xD8, 0x78, 0x9C, 0x6C, 0xE1, 0xB1, 0xC9, 0x39, 0x8D, 0x2D,
                0xD2, 0x72, 0xC6, 0x36, 0x4E, 0x1E, 0x93, 0x63, 0x87, 0x27, 0x4B, 0x1B]
    live = {0}            # physical slots holding defined points; P starts in slot 0
    for b in ops:
        t, p = b & 7, perm_tab[b >> 3]
        A, B, C, T = p & 3, (p >> 2) & 3, (p >> 4) & 3, (p >> 6) & 3
        assert len({A, B, C, T}) == 4
        if t == 0:
            assert T in live
        elif t == 1:
            assert B in live; live = {A, B, C}
        elif t == 2:
            assert {A, B, C} <= live; live = {A, B, C, T}
        elif t in (3, 4, 5):
            assert {A, B, C} <= live
        elif t == 6:
            assert {A, B, C} <= live; live = {T}


@pytest.mark.parametrize("lo,hi,b1", [(50000, 5000000, 50000), (3000, 300000, 3000), (1500, 150000, 1500),
                                      (400, 40000, 400), (200, 20000, 200), (100, 10000, 100), (50, 5000, 50),
                                      (100002000, 100100000, 2000)])
def test_pair_matches_oracle(lo, hi, b1):
    D, U, L, R = E.stage2_params(b1)
    assert D == O.lib().oracle_stage2_D(b1) and (U, L) == (16, 32)
    assert E.pair(lo, hi, D) == O.pair(lo, hi, D)


def test_pair_counts_match_reference_printout():
    g = GOLDEN["readme508_b1_5e4"]["counts"]
    v, u, amin, npairs = E.pair(50000, 5000000, g["D"])
    assert (len(v), npairs) == (g["pairmap_steps"], g["s2_paired"])
    D, U, L, R = E.stage2_params(50000)
    assert (D, U, L, R - 3) == (g["D"], g["U"], g["L"], g["R"])


@pytest.mark.parametrize("name", ["readme508_b1_5e4", "syn415_b1_1e5", "small96_D1155", "small96_D385", "small96_D210",
                                  "small96_D120", "small96_D60", "small96_D30", "syn2048_b1_5e3", "syn415_two_ranges"])
def test_stage2_program_counters_match_reference_printout(name):
    # "performed %u pt-adds, %u inversions, and %u pair-muls in stage 2" (ecm.c:1482) + pairmap steps
    g = GOLDEN[name]
    c = E.plan_stage2(g["b1"], g["b2"])
    ref = g["counts"]
    assert {k: c[k] for k in ("s2_ptadds", "s2_numinv", "s2_paired", "pairmap_steps")} == \
           {k: ref[k] for k in ("s2_ptadds", "s2_numinv", "s2_paired", "pairmap_steps")}


CLI = os.path.join(ROOT, "avx-ecm_b200", "avx-ecm-b200")


def _fib(n):
    a, b = 0, 1
    for _ in range(n):
        a, b = b, a + b
    return a


def _fact(n):
    r = 1
    for i in range(2, n + 1):
        r *= i
    return r


@pytest.mark.parametrize("expr,value", [
    ("fib(791)/13/677/216416017", _fib(791) // 13 // 677 // 216416017),
    ("(2+110!)/446", (2 + _fact(110)) // 446),
    ("(26*10^238-17)/(9*3*31*17914895525348997871953180891109)", (26 * 10 ** 238 - 17) // (9 * 3 * 31 * 17914895525348997871953180891109)),
    ("((5801^61-1)/((5801-1)*4027*5763040637*48081214823351791*195790721913324330907*1811340220375495245599))",
     (5801 ** 61 - 1) // ((5801 - 1) * 4027 * 5763040637 * 48081214823351791 * 195790721913324330907 * 1811340220375495245599)),
    ("2^3^2", 512), ("7 - 2 - 1", 4), ("-3 + 10 % 4", -1), ("13# + gcd(12, 18) + 1<<4", (30030 + 6 + 1) << 4),
    ("modexp(3, 100, 1000007) + sqrt(1000000) + lg2(1024)", pow(3, 100, 1000007) + 1000 + 11),
    ("0x10 * 2", 32),
])
def test_cli_expression_evaluator_matches_python(expr, value):
    # the inputs of the reference's test.csh use this expression language (calc.c)
    import subprocess
    out = subprocess.run([CLI, "--eval", expr], capture_output=True, text=True, check=True).stdout.strip()
    assert int(out) == value


def test_header_is_plain_c_and_links(tmp_path):
    """include/ecm_b200.h compiles as C99 and a C client links against the library (examples/abi_check.c);
    without a GPU the client must see the engine refuse to run (no CPU fallback)."""
    import subprocess
    exe = str(tmp_path / "abi_check")
    libdir = os.path.join(ROOT, "avx-ecm_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "abi_check.c"), "-o", exe, "-L", libdir, "-lecm_b200",
                    "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "1980817 point-adds" in r.stdout and "D=2310 U=16 L=32 R=963" in r.stdout


def test_planners_match_oracle_on_random_bounds():
    """Seeded random B1 / prime ranges: the PRAC plan and PAIR output must equal the oracle's."""
    import random
    rng = random.Random(20261018)
    for _ in range(6):
        b1 = rng.randrange(7, 60000)
        ops, adds, dups = E.plan_stage1(b1)
        ref = O.stage1_trace(b1).decode().replace("S", "")
        assert "".join(TYPE_CH[b & 7] for b in ops) == ref, b1
    for _ in range(6):
        b1 = rng.choice([rng.randrange(40, 5000), rng.randrange(5000, 400000)])
        lo = b1 + rng.randrange(0, 3) * 1000
        hi = lo + rng.randrange(1, 2000000)
        D, U, L, R = E.stage2_params(b1)
        assert E.pair(lo, hi, D) == O.pair(lo, hi, D), (lo, hi, D)


@pytest.mark.parametrize("b1,b2", [(3000, 300000), (50000, 5000000), (2000, 100100000)])
def test_caller_pairmap_compiles_to_the_range_program(b1, b2):
    """ecm_b200_stage2_range takes the reference's own arguments (steps, pm_v, pm_u with work->amin, ecm.c:2342-2351).
    Fed with the ORACLE's pair() output -- the arrays a maintainer who keeps the reference's pair() would pass -- it must
    compile to exactly the program ecm_b200_stage2 runs for the same prime range."""
    D, U, L, R = E.stage2_params(b1)
    lo = b1
    which = 0
    while lo < b2:
        hi = min(lo + 100000000, b2)
        v, u, amin_final, npairs = O.pair(lo, hi, D)
        amin = (lo + D) // (2 * D)
        prog = E.stage2_pairmap_program(b1, amin, v, u)
        ref, _ = E.stage2_program(b1, b2, which)
        assert prog == ref and len(prog) > 0
        lo, which = hi, which + 1


def test_bad_pairmaps_are_rejected():
    b1 = 3000
    D, U, L, R = E.stage2_params(b1)
    v, u, _, _ = O.pair(b1, 300000, D)
    amin = (b1 + D) // (2 * D)
    k = next(i for i in range(len(v)) if v[i] or u[i])
    assert E.stage2_pairmap_program(b1, amin, v, u)
    assert E.stage2_pairmap_program(b1, 0, v, u) == []                              # A - w < 0
    bad = list(v); bad[k] = amin + 2 * L                                            # beyond the giant-step window
    assert E.stage2_pairmap_program(b1, amin, bad, u) == []
    bad = list(v); bad[k] = amin - 1 if amin else 0
    assert E.stage2_pairmap_program(b1, amin + 1, bad, u) == []                     # before the window
    bad = list(u); bad[k] = U * (D + 1) + 3                                         # outside the baby-step map
    assert E.stage2_pairmap_program(b1, amin, v, bad) == []
    bad = list(u); bad[k] = 5                                                       # 5 | D for every D: baby step 5 is not stored
    assert E.stage2_pairmap_program(b1, amin, v, bad) == []
