#!/usr/bin/env python3
"""Generate tests/golden/*.json from the COMPILED REFERENCE (oracle/_ref/avx-ecm-ref).

TEST INFRASTRUCTURE.  Runs only in the dev container (needs /root/reference to have
been compiled by oracle/build_ref.sh).  For every case the unmodified reference binary is
run with threads=1 and an explicit sigma in a scratch directory; we keep

  * the save_b1.txt lines byte for byte                       (ecm.c:1372-1380)
  * stage-1 Z and stage-2 accumulator of every lane, captured by the mpz_gcd LD_PRELOAD
    tap (oracle/shim/gcd_tap.c) and taken out of Montgomery form (x * R^-1 mod N,
    R = 2^MAXBITS as printed by main.c:529-533), i.e. the R-independent true residues
  * the "found ... factor" reports                            (ecm.c:1356-1358,1510-1513)
  * the op counters the reference prints                      (ecm.c:1849,1482, pairmap steps)

Usage: python tools/gen_golden.py [case ...]
"""
import json, os, re, subprocess, sys, tempfile, random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "avx-ecm-ref")
TAP = os.path.join(ROOT, "oracle", "_ref", "gcd_tap.so")
OUT = os.path.join(ROOT, "tests", "golden")


def is_prime(n, rounds=24):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    rng = random.Random(n & 0xFFFFFFFF)
    for _ in range(rounds):
        a = rng.randrange(2, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def next_prime(n):
    n |= 1
    while not is_prime(n):
        n += 2
    return n


def synthetic(bits_p, bits_q, seed):
    """N = p*q, p,q = next primes after seeded random odd numbers with the top bit set."""
    rng = random.Random(seed)
    p = next_prime(rng.getrandbits(bits_p) | (1 << (bits_p - 1)) | 1)
    q = next_prime(rng.getrandbits(bits_q) | (1 << (bits_q - 1)) | 1)
    return p * q


def fib(n):
    a, b = 0, 1
    for _ in range(n):
        a, b = b, a + b
    return a


def composites():
    return {
        "syn415": synthetic(207, 208, 12345),     # 415 bits -> NWORDS=8 class
        "syn1024": synthetic(512, 512, 12346),
        "syn2048": synthetic(1024, 1024, 12347),
        "readme508": fib(791) // 13 // 677 // 216416017,
        "t35": 142946323174762557214361604817789197531833590620956958433836799929503392464892596183803921,
        "csh250k": 171527316193270871507108435893460246746982712299171622350010323023149618461701108180621787596877308885636902619030669,
        "csh1m": 7908926676514675413083853032827063880118980193445471625562601469958414706043143581401715516956542424923236530406833110566233,
        "small96": 1000000007 * 998244353 * 4294967311,   # tiny composite: exercises factor/inversion-failure paths
        "csh150m": 19223719229397103735869895564468606263251785680561653388554202432164204897138631706690937388406707574740021324772129,
        "syn206": synthetic(103, 103, 12348),     # 206 bits: the reference's smallest word count, for the B1 > 1e8 case
        # sizes that land in the in-between kernel sets (28 limbs one thread per curve; 40 and 56 limbs four lanes per curve)
        "syn880": synthetic(440, 440, 12349),
        "syn1250": synthetic(625, 625, 12350),
        "syn1750": synthetic(875, 875, 12351),
    }


def csh_line(k):
    """Line k (1-based) of the reference's own regression script test.csh: (N, curves, B1, B2, sigma).  Expressions
    are evaluated here with Python integers (the calc.c grammar of these lines is + - * / ^ ! and parentheses)."""
    from math import factorial  # noqa: F401  (used by eval below)
    lines = [l for l in open("/root/reference/test.csh") if l.startswith("./avx-ecm")]
    m = re.match(r"./avx-ecm\s+(\"[^\"]+\"|'[^']+'|\S+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)", lines[k - 1])
    e = m.group(1).strip("\"'").replace("^", "**").replace("/", "//")
    e = re.sub(r"(\d+)!", r"factorial(\1)", e)
    return (eval(e), int(m.group(2)), int(m.group(3)), int(m.group(5)), int(m.group(6)))


def cases():
    c = composites()
    extra = [
        # BASELINE config [0]: the README example at full size; sigma 1007 (lane 7) finds 272602401466814027129 in stage 1
        ("readme508_b1_1e6", c["readme508"], 8, 1000000, 100000000, 1000),
        # test_t35.csh: the two other survey-verified stage-2 hits, full size
        ("t35_full_sigma_11919771003873180376", c["t35"], 8, 1000000, 100000000, 11919771003873180376),
        ("t35_full_sigma_10019108749973911965", c["t35"], 8, 1000000, 100000000, 10019108749973911965),
        # wider operands at a B1 that is not a toy
        ("syn1024_b1_1e5", c["syn1024"], 8, 100000, 10000000, 7),
        ("syn2048_b1_5e4", c["syn2048"], 8, 50000, 5000000, 7),
    ]
    # more lines of test.csh (default GPU suite: B1 <= 3e6, B2 <= 1.05e9)
    for k in (2, 9, 11, 12, 13, 14, 19, 22, 25):
        extra.append(("csh_line%02d" % k,) + csh_line(k))
    # slow ones (ECM_B200_SLOW=1 on the GPU): 1165-bit input, "huge B1", "huge B2"  (test.csh:7, 33-37)
    for k in (7, 26, 27):
        extra.append(("slow_csh_line%02d" % k,) + csh_line(k))
    return extra + [
        # name, N, curves, B1, B2, sigma0
        ("readme508_b1_5e4", c["readme508"], 8, 50000, 5000000, 1000),
        ("syn415_b1_1e5", c["syn415"], 8, 100000, 10000000, 7),
        ("syn415_b1_3e4_s1only", c["syn415"], 16, 30000, 30000, 100),
        ("syn1024_b1_2e4", c["syn1024"], 8, 20000, 2000000, 7),
        ("syn2048_b1_5e3", c["syn2048"], 8, 5000, 500000, 7),
        ("syn880_b1_2e4", c["syn880"], 8, 20000, 2000000, 7),
        ("syn1250_b1_1e4", c["syn1250"], 8, 10000, 1000000, 7),
        ("syn1750_b1_5e3", c["syn1750"], 8, 5000, 500000, 7),
        ("csh250k_stage1_factor", c["csh250k"], 8, 250000, 250000, 3462348953),
        ("csh1m_stage1_factor", c["csh1m"], 8, 1000000, 1000000, 7372562557),
        ("t35_stage2_factor", c["t35"], 8, 1000000, 100000000, 416265588),
        ("t35_b1_2e4_sigma64", c["t35"], 8, 20000, 2000000, 11919771003873180376),
        ("small96_D1155", c["small96"], 8, 3000, 300000, 11),
        ("small96_D385", c["small96"], 8, 1500, 150000, 11),
        ("small96_D210", c["small96"], 8, 400, 40000, 11),
        ("small96_D120", c["small96"], 8, 200, 20000, 11),
        ("small96_D60", c["small96"], 8, 100, 10000, 11),
        ("small96_D30", c["small96"], 8, 50, 5000, 11),
        ("syn415_two_ranges", c["syn415"], 8, 2000, 100100000, 7),   # B2 spans two 1e8 prime ranges
        ("csh150m_b1_1e6_two_ranges", c["csh150m"], 8, 1000000, 150000000, 3018506502),   # test.csh line 6
        # stage 1 over two 1e8 prime ranges (ecm.c:1207-1311): repeated doublings, skipped first prime, checkpoint.txt
        ("syn206_b1_1.1e8_two_stage1_ranges", c["syn206"], 8, 110000000, 110000000, 1000003),
    ]


def run_case(name, N, curves, B1, B2, sigma0):
    with tempfile.TemporaryDirectory() as d:
        tap = os.path.join(d, "tap.log")
        env = dict(os.environ, GCD_TAP_FILE=tap, LD_PRELOAD=TAP)
        out = subprocess.run([REF, str(N), str(curves), str(B1), "1", str(B2), str(sigma0)],
                             cwd=d, env=env, capture_output=True, text=True, check=True).stdout
        save = open(os.path.join(d, "save_b1.txt")).read().splitlines(keepends=True)
        taps = [l.split()[1:] for l in open(tap).read().splitlines()]
        ckpt = os.path.join(d, "checkpoint.txt")
        checkpoint = open(ckpt).read().splitlines(keepends=True) if os.path.exists(ckpt) else []
    maxbits = int(re.search(r"Choosing MAXBITS = (\d+)", out).group(1))
    Rinv = pow(1 << maxbits, -1, N)
    do2 = B2 > B1
    nb = len(save) // 8                      # batches actually run (stops after a find, ecm.c:1531)
    # gcd calls per batch: 8 x stage-1 Z, [inversion failures...], 8 x stage-2 acc (all with b == N)
    z1, acc, inv_fail_operands = [], [], []
    calls = [(int(a, 16), int(b, 16)) for a, b in taps if int(b, 16) == N]
    if checkpoint:
        calls = calls[len(checkpoint):]          # the checkpoint writer runs check_factor too (ecm.c:1262)
    per = []
    # split calls into batches: the reference processes batches sequentially
    idx = 0
    for b in range(nb):
        z1 += [a * Rinv % N for a, _ in calls[idx:idx + 8]]
        idx += 8
        if do2:
            # everything up to the last 8 calls of this batch are inversion-failure gcds
            nxt = idx
            # count failure gcds: total calls minus 16 per batch when only one batch; for
            # several batches we rely on the run stopping at the first find.
            remaining = len(calls) - idx - 8 - (nb - b - 1) * 16
            inv_fail_operands += [hex(a) for a, _ in calls[idx:idx + remaining]]
            idx += remaining
            acc += [a * Rinv % N for a, _ in calls[idx:idx + 8]]
            idx += 8
    factors = []
    for m in re.finditer(r"found \S+ factor (\d+) in stage (\d) \(B\d = \d+\): thread 0, vec (\d+), sigma (\d+)", out):
        factors.append({"factor": m.group(1), "stage": int(m.group(2)), "lane": int(m.group(3)), "sigma": m.group(4)})
    cnt = {}
    m = re.findall(r"with (\d+) point-adds and (\d+) point-doubles", out)[-1]     # the counters run on over the prime ranges
    cnt["s1_ptadds"], cnt["s1_ptdups"] = int(m[0]), int(m[1])
    if do2:
        m = re.search(r"performed (\d+) pt-adds, (\d+) inversions, and (\d+) pair-muls", out)
        cnt["s2_ptadds"], cnt["s2_numinv"], cnt["s2_paired"] = map(int, m.groups())
        cnt["pairmap_steps"] = sum(int(x) for x in re.findall(r"pairmap step 0 of (\d+)", out))
        m = re.search(r"w = (\d+), R = (\d+), L = (\d+), U = (\d+)", out)
        cnt["D"], cnt["R"], cnt["L"], cnt["U"] = map(int, m.groups())
    return {
        "name": name, "n": str(N), "curves_requested": curves, "b1": B1, "b2": B2, "sigma0": str(sigma0),
        "maxbits_ref": maxbits, "save_lines": save,
        "z1_true_hex": [hex(x)[2:] for x in z1], "acc_true_hex": [hex(x)[2:] for x in acc],
        "inv_fail_gcd_calls": len(inv_fail_operands),
        "factors": factors, "counts": cnt, "checkpoint_lines": checkpoint,
    }


def special_cases():
    """Special-form inputs (main.c:405-457): given to the reference as expressions, like a user would.
    sigma0 is 'random-looking' on purpose: with tiny sigmas the first residues mod 2^k+-1 contain words that are
    exactly 0 or 2^52-1, which trips the reference's carry helpers (vecarith52.c:102-116 report a carry-out
    whenever the word is all-ones / zero, even without a carry-in) in a lane-coupled way no per-curve engine can
    reproduce -- see DESIGN.md."""
    return [
        # name, expression, curves, B1, B2, sigma0
        ("special_m277", "2^277-1", 8, 2000, 200000, 1000003),
        ("special_p523", "(2^523+1)/3", 8, 2000, 200000, 1000003),          # base divisible by 3: inversions fail
        ("special_pm220_69", "2^220-69", 8, 2000, 200000, 1000003),
        ("special_pm220_57", "2^220-57", 8, 3000, 300000, 424242421),
        ("special_m127x", "(2^254-1)/(2^127-1)/3", 8, 1000, 100000, 77777771),   # 2^127+1 over 3: found as 2^127+1
        ("special_redc", "36667531*129175771*58052548129*83207209", 8, 2000, 200000, 1000003),  # | 2^523+1, REDC wins
        ("special_pm128_169", "2^64+13", 8, 1500, 150000, 31415927),            # found as 2^128-169: N far shorter than its base
        ("special_p128", "(2^128+1)/59649589127497217", 8, 1500, 150000, 27182819),  # k a multiple of 32
    ]


def run_special_case(name, expr, curves, B1, B2, sigma0):
    with tempfile.TemporaryDirectory() as d:
        tap = os.path.join(d, "tap.log")
        env = dict(os.environ, GCD_TAP_FILE=tap, LD_PRELOAD=TAP)
        out = subprocess.run([REF, expr, str(curves), str(B1), "1", str(B2), str(sigma0)],
                             cwd=d, env=env, capture_output=True, text=True, check=True).stdout
        save = open(os.path.join(d, "save_b1.txt")).read().splitlines(keepends=True)
        taps = [l.split()[1:] for l in open(tap).read().splitlines()]
    N = int(re.search(r"commencing parallel ecm on (\d+)", out).group(1))
    m = re.search(r"Using special (?:pseudo-)?Mersenne mod for factor of: 2\^(\d+)([-+])(\d+)", out)
    kind, k, base = 0, N.bit_length(), None
    if m:
        k, c = int(m.group(1)), int(m.group(3))
        kind = c if m.group(2) == "-" else -1
        base = (1 << k) - c if kind > 0 else (1 << k) + 1
    maxbits = int(re.search(r"Choosing MAXBITS = (\d+)", out).group(1))
    Rinv = 1 if base else pow(1 << maxbits, -1, N)
    mod = base or N
    if abs(kind) == 1:
        taps = taps[1:]                     # main.c:456 gcd(input, primitive part)
    calls = [(int(a, 16), int(b, 16)) for a, b in taps if len(b) and int(b, 16) == N]
    fails = [(int(a, 16), int(b, 16)) for a, b in taps if base and int(b, 16) == base]
    z1 = [a * Rinv % mod for a, _ in calls[:8]]
    acc = [a * Rinv % mod for a, _ in calls[-8:]] if B2 > B1 else []
    factors = []
    for mm in re.finditer(r"found \S+ factor (\d+) in stage (\d) \(B\d = \d+\): thread 0, vec (\d+), sigma (\d+)", out):
        factors.append({"factor": mm.group(1), "stage": int(mm.group(2)), "lane": int(mm.group(3)), "sigma": mm.group(4)})
    cnt = {}
    mm = re.search(r"with (\d+) point-adds and (\d+) point-doubles", out)
    cnt["s1_ptadds"], cnt["s1_ptdups"] = int(mm.group(1)), int(mm.group(2))
    if B2 > B1:
        mm = re.search(r"performed (\d+) pt-adds, (\d+) inversions, and (\d+) pair-muls", out)
        cnt["s2_ptadds"], cnt["s2_numinv"], cnt["s2_paired"] = map(int, mm.groups())
        cnt["pairmap_steps"] = sum(int(x) for x in re.findall(r"pairmap step 0 of (\d+)", out))
    return {
        "name": name, "expr": expr, "n": str(N), "kind": kind, "k": k, "base": str(base) if base else None,
        "curves_requested": curves, "b1": B1, "b2": B2, "sigma0": str(sigma0), "maxbits_ref": maxbits,
        "redc_forced": "determined to be faster by REDC" in out,
        "save_lines": save[:8], "z1_true_hex": [hex(x)[2:] for x in z1], "acc_true_hex": [hex(x)[2:] for x in acc],
        "inv_fail_gcd_calls": len(fails), "factors": factors, "counts": cnt,
    }


def main():
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:2] == ["special"]:
        for c in special_cases():
            if sys.argv[2:] and c[0] not in sys.argv[2:]:
                continue
            g = run_special_case(*c)
            json.dump(g, open(os.path.join(OUT, c[0] + ".json"), "w"), indent=1)
            print(c[0], "kind", g["kind"], "k", g["k"], "redc", g["redc_forced"], "factors",
                  [(f["sigma"], f["stage"], f["factor"]) for f in g["factors"]], g["counts"], "invfail", g["inv_fail_gcd_calls"], flush=True)
        return
    comp = {k: str(v) for k, v in composites().items()}
    json.dump(comp, open(os.path.join(OUT, "composites.json"), "w"), indent=1)
    want = set(sys.argv[1:])
    for c in cases():
        if want and c[0] not in want:
            continue
        g = run_case(*c)
        json.dump(g, open(os.path.join(OUT, c[0] + ".json"), "w"), indent=1)
        print(c[0], "lanes", len(g["save_lines"]), "factors", [(f["sigma"], f["stage"], f["factor"]) for f in g["factors"]],
              g["counts"], "invfail", g["inv_fail_gcd_calls"], flush=True)


if __name__ == "__main__":
    main()
