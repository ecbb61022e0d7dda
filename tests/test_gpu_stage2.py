"""GPU parity of stage 2 (ecm_stage2_init + ecm_stage2_pair driven by PAIR) against the golden
vectors of the compiled reference: the stage-2 accumulator of every curve (true residue), the
factors reported in stage 1 and stage 2, and the op counters.  Through the C ABI."""
import pytest
from conftest import GOLDEN, golden_factor, golden_base, composites, is_known_answer
import oracle_lib as O
import avx_ecm_b200 as E

pytestmark = pytest.mark.gpu


def usable(g):
    return int(g["n"]).bit_length() <= 2048 and g["b2"] > g["b1"] and not is_known_answer(g["name"])


@pytest.mark.parametrize("name", sorted(k for k, g in GOLDEN.items() if usable(g)))
def test_stage2_matches_reference_golden(name):
    g = GOLDEN[name]
    N, b1, b2, s0 = int(g["n"]), g["b1"], g["b2"], int(g["sigma0"])
    lanes = len(g["save_lines"])
    ctx = E.EcmContext(N, lanes, base=golden_base(g))
    try:
        r = E.vececm(N, lanes, b1, b2, sigma=s0, ctx=ctx)
        cnt = ctx.stage2_counters()
    finally:
        ctx.close()
    assert r["save_lines"] == g["save_lines"]
    ref = g["counts"]
    assert cnt == {k: ref[k] for k in cnt}
    exp_acc = [int(a, 16) for a in g["acc_true_hex"]]
    for i in range(lanes):
        sigma = s0 + i
        # factors: every lane, both stages
        got = {(s, st): f for s, st, f in r["factors"] if s == sigma}
        assert got.get((sigma, 1), 0) == golden_factor(g, sigma, 1)
        assert got.get((sigma, 2), 0) == golden_factor(g, sigma, 2)
        # accumulator: bit-exact wherever the reference's value is R-independent (no failed inversion),
        # and in vector lane 0 even then (see oracle/ecm_oracle.c batch_invert)
        if not r["inv_fail"][i] or i == 0:
            assert r["acc"][i] == exp_acc[i], "lane %d" % i


def test_stage2_matches_oracle_on_fresh_sigmas():
    # sigmas that are in no golden file, 415- and 1024-bit, against the pinned oracle
    c = composites()
    for N, b1, b2, sig in ((c["syn415"], 3000, 200000, [12345, 2 ** 40 + 3, 99]), (c["syn1024"], 5000, 300000, [31, 2 ** 63 + 9])):
        ctx = E.EcmContext(N, len(sig))
        try:
            ctx.build_curves(sig)
            ctx.stage1(b1)
            x, z, f1 = ctx.read_stage1()
            ctx.stage2(b1, b2)
            acc, f2, fail = ctx.read_stage2()
        finally:
            ctx.close()
        for i, s in enumerate(sig):
            o = O.ecm_curve(N, b1, b2, s)
            assert (x[i], z[i], f1[i], acc[i], f2[i]) == (o["x"], o["z"], o["f1"], o["acc"], o["f2"])


def test_stage2_many_curves_spanning_blocks():
    # more curves than one block; all must agree with single-curve oracle runs (spot checks)
    N = composites()["syn415"]
    count, b1, b2 = 700, 2000, 60000
    r = E.vececm(N, count, b1, b2, sigma=1000)
    for i in (0, 287, 288, 699):
        o = O.ecm_curve(N, b1, b2, 1000 + i)
        assert r["acc"][i] == o["acc"] and r["x"][i] == o["x"]


def test_cli_writes_reference_files(tmp_path):
    """The command-line driver (avx-ecm's argv contract) appends the reference's save_b1.txt lines byte
    for byte and reports the same factors in ecm_results.txt."""
    import os, re, subprocess
    from conftest import ROOT
    cli = os.path.join(ROOT, "avx-ecm_b200", "avx-ecm-b200")
    for name in ("readme508_b1_5e4", "small96_D210", "special_m277", "special_p523", "special_pm220_57", "special_redc"):
        g = GOLDEN[name]
        d = tmp_path / name
        d.mkdir()
        expr = "fib(791)/13/677/216416017" if name.startswith("readme") else g.get("expr", g["n"])
        out = subprocess.run([cli, expr, str(len(g["save_lines"])), str(g["b1"]), "1", str(g["b2"]), g["sigma0"]],
                             cwd=d, capture_output=True, text=True, check=True).stdout
        assert open(d / "save_b1.txt").read() == "".join(g["save_lines"])
        got = set()
        res = (d / "ecm_results.txt").read_text() if (d / "ecm_results.txt").exists() else ""
        for m in re.finditer(r"found \S+ factor (\d+) in stage (\d) .*sigma (\d+)", res):
            got.add((m.group(3), int(m.group(2)), m.group(1)))
        assert got == {(f["sigma"], f["stage"], f["factor"]) for f in g["factors"]}


def test_stage2_in_several_waves_equals_one_wave(monkeypatch):
    """When the stage-2 tables of a batch do not fit in HBM the batch is processed in waves (1024-bit,
    65 536 curves needs two).  Force small waves and compare with the single-wave result and the oracle."""
    N = composites()["syn415"]
    count, b1, b2 = 900, 1500, 40000
    one = E.vececm(N, count, b1, b2, sigma=77)
    monkeypatch.setenv("ECM_B200_S2_WAVE", "300")
    ctx = E.EcmContext(N, count)
    try:
        many = E.vececm(N, count, b1, b2, sigma=77, ctx=ctx)
    finally:
        ctx.close()
    assert many["acc"] == one["acc"] and many["factors"] == one["factors"] and many["inv_fail"] == one["inv_fail"]
    for i in (0, 299, 300, 511, 899):
        assert many["acc"][i] == O.ecm_curve(N, b1, b2, 77 + i)["acc"]


def test_plain_c_client_runs_both_stages(tmp_path):
    """examples/abi_check.c (C99, no Python, no torch) drives stage 1 + stage 2 through the C ABI:
    N = (2^89-1)(2^107-1), B1 = 2e4, B2 = 2e6 must expose the factor 2^89-1 for some sigma."""
    import os, subprocess
    from conftest import ROOT
    exe = str(tmp_path / "abi_check")
    libdir = os.path.join(ROOT, "avx-ecm_b200")
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "abi_check.c"),
                    "-o", exe, "-L", libdir, "-lecm_b200", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    N = (2 ** 89 - 1) * (2 ** 107 - 1)
    lines = [l for l in r.stdout.splitlines() if l.startswith("sigma")]
    assert len(lines) == 8
    for i, l in enumerate(lines):
        o = O.ecm_curve(N, 20000, 2000000, 1000 + i)
        got = int(l.split("0x")[1], 16) if "factor 0x" in l else 0
        assert got == o["f2"], l


def test_cli_on_two_gpus_writes_the_single_gpu_files(tmp_path, monkeypatch):
    """The 4th argument (the reference's thread count) is the number of GPUs: one host thread and one context per
    GPU, disjoint sigma slices, files merged in sigma order -- byte-identical to the one-GPU / reference output.
    Also with stage-1 prime ranges (checkpoint.txt) and a special-form input.  Skipped on a one-GPU box."""
    import os, subprocess
    from conftest import ROOT
    try:
        ngpu = len(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.strip().splitlines())
    except OSError:
        ngpu = 0
    if ngpu < 2:
        pytest.skip("needs two GPUs")
    cli = os.path.join(ROOT, "avx-ecm_b200", "avx-ecm-b200")
    for name in ("syn415_b1_3e4_s1only", "small96_D210", "special_p523"):
        g = GOLDEN[name]
        d = tmp_path / name
        d.mkdir()
        subprocess.run([cli, g.get("expr", g["n"]), str(len(g["save_lines"])), str(g["b1"]), "2", str(g["b2"]), g["sigma0"]],
                       cwd=d, capture_output=True, text=True, check=True)
        assert open(d / "save_b1.txt").read() == "".join(g["save_lines"])
    # range-by-range stage 1 on two GPUs == on one
    monkeypatch.setenv("ECM_B200_S1_RANGE", "2000")
    N = composites()["t35"]
    outs = []
    for gpus in ("1", "2"):
        d = tmp_path / ("ranges" + gpus)
        d.mkdir()
        subprocess.run([cli, str(N), "10", "5000", gpus, "5000", "424242"], cwd=d, capture_output=True, text=True, check=True)
        outs.append(((d / "checkpoint.txt").read_text(), (d / "save_b1.txt").read_text()))
    assert outs[0] == outs[1] and outs[0][0].count("\n") == 20


def test_stage2_at_the_reference_granularity_with_oracle_pairmaps():
    """ecm_b200_stage2_init / ecm_b200_stage2_range (the reference's ecm_stage2_init -> foundDuringInv and
    ecm_stage2_pair(steps, pm_v, pm_u), ecm.c:67-72) driven like vececm's loop (ecm.c:1400-1476) with the pairmaps of
    the ORACLE's pair(): two prime ranges, accumulators read after init, after each range, equal to ecm_b200_stage2
    and to the oracle."""
    N = composites()["syn415"]
    b1, b2, sig = 2000, 100100000, [7, 8, 9, 2 ** 63 + 11]       # the golden syn415_two_ranges covers sigma 7..14
    D, U, L, R = E.stage2_params(b1)
    ctx = E.EcmContext(N, len(sig))
    try:
        ctx.build_curves(sig); ctx.stage1(b1)
        assert ctx.stage2_init(b1) is False
        acc0, _, fail0 = ctx.read_stage2()
        assert acc0 == [1] * len(sig) and fail0 == [0] * len(sig)               # acc = one (ecm.c:2318)
        lo = b1
        while lo < b2:
            hi = min(lo + 100000000, b2)
            v, u, _, _ = O.pair(lo, hi, D)
            ctx.stage2_range((lo + D) // (2 * D), v, u)
            lo = hi
        acc, f2, fail = ctx.read_stage2()
        with pytest.raises(E.EcmError, match="pairmap leaves"):
            ctx.stage2_range(0, v, u)
        ctx.build_curves(sig); ctx.stage1(b1)
        with pytest.raises(E.EcmError, match="stage2_init has not been run"):
            ctx.stage2_range((b1 + D) // (2 * D), v, u)
        ctx.stage2(b1, b2)
        assert ctx.read_stage2() == (acc, f2, fail)
    finally:
        ctx.close()
    for i, s in enumerate(sig):
        assert acc[i] == O.ecm_curve(N, b1, b2, s)["acc"]
    g = GOLDEN["syn415_two_ranges"]
    assert acc[:3] == [int(a, 16) for a in g["acc_true_hex"][:3]]


def test_stage2_init_reports_found_during_inv():
    g = GOLDEN["small96_D210"]                                   # tiny composite: inversions fail on some curves
    N, b1, s0 = int(g["n"]), g["b1"], int(g["sigma0"])
    ctx = E.EcmContext(N, 8)
    try:
        ctx.build_curves([s0 + i for i in range(8)]); ctx.stage1(b1)
        found = ctx.stage2_init(b1)
        _, _, fail = ctx.read_stage2()
        assert found == any(fail)
        D = E.stage2_params(b1)[0]
        v, u, _, _ = O.pair(b1, g["b2"], D)
        ctx.stage2_range((b1 + D) // (2 * D), v, u)
        acc, f2, fail = ctx.read_stage2()
    finally:
        ctx.close()
    assert any(fail) and g["inv_fail_gcd_calls"] > 0
    for i in range(8):
        assert f2[i] == golden_factor(g, s0 + i, 2)
        if not fail[i] or i == 0:
            assert acc[i] == int(g["acc_true_hex"][i], 16)


def test_stage2_state_checks_and_resume_from_loaded_points():
    N = composites()["syn415"]
    sig = [7, 8, 9]
    ctx = E.EcmContext(N, 200)
    try:
        ctx.build_curves(list(range(100, 300)))
        ctx.stage1_begin(20000)
        ctx.stage1_step(1)                                       # stage 1 only partly run
        with pytest.raises(E.EcmError, match="still in progress"):
            ctx.stage2(20000, 2000000)
        with pytest.raises(E.EcmError, match="still in progress"):
            ctx.stage1_range(20000, 0)
        with pytest.raises(E.EcmError, match="still in progress"):
            ctx.stage1_begin(20000)
        ctx.build_curves(sig); ctx.stage1(3000)
        with pytest.raises(E.EcmError, match="B1 too small"):
            ctx.stage2(20, 2000)                                 # amin = 0: the reference aborts in next_pt_vec
        x, z, _ = ctx.read_stage1()
        ctx.stage2(3000, 300000)
        acc = ctx.read_stage2()[0]
        # resume: the stage-1 points of a save_b1.txt file loaded as X/Z with the curve parameter, then stage 2 only
        # (GMP-ECM's -resume work flow); stage 2 normalises every table entry, so the accumulator is the same
        xs = [xi * pow(zi, -1, N) % N for xi, zi in zip(x, z)]
        ss = [O.build_curve(N, s)[1] for s in sig]
        ctx.load_curves(xs, ss, sig)
        ctx.stage2(3000, 300000)
        assert ctx.read_stage2()[0] == acc
    finally:
        ctx.close()


def test_cli_random_sigmas_are_valid_resume_lines(tmp_path):
    """sigma omitted / 0: the CLI draws sigmas with the reference's LCG, all >= 6 (ecm.c:1564-1570, main.c:757-763);
    every save_b1.txt line must still be the stage-1 point of the curve it names."""
    import os, re, subprocess
    from conftest import ROOT
    from test_resume_lines_cpu import check_line
    N = composites()["t35"]
    cli = os.path.join(ROOT, "avx-ecm_b200", "avx-ecm-b200")
    for args in ([str(N), "12", "3000", "1", "3000"], [str(N), "12", "3000", "1", "3000", "0"]):
        d = tmp_path / ("r%d" % len(args))
        d.mkdir()
        subprocess.run([cli] + args, cwd=d, capture_output=True, text=True, check=True)
        lines = (d / "save_b1.txt").read_text().splitlines(keepends=True)
        sig = [int(re.search(r"SIGMA=(\d+)", l).group(1)) for l in lines]
        assert len(lines) == 12 and len(set(sig)) == 12 and min(sig) >= 6
        assert [check_line(l, expect_n=N) for l in lines] == ["ok"] * 12


def _next_prime(n):
    n |= 1
    while not all(pow(a, n - 1, n) == 1 for a in (2, 3, 5, 7, 11, 13)):      # Fermat tests: plenty for a test modulus
        n += 2
    return n


@pytest.mark.parametrize("bits", [1100, 1400, 1700, 2000])
def test_cooperative_stage2_equals_one_thread_stage2_and_oracle(bits, monkeypatch):
    """40 / 48 / 56 / 64 limbs: stage 2 on the four-lanes-per-curve layout (coop_s2.cuh: striped tables, pair loop with register
    accumulators, inverse by the group's lane 0) against the one-thread-per-curve kernels and the oracle -- on a composite
    with a 30-bit prime factor, so that inversions FAIL on some curves (foundDuringInv path, lane-0 semantics) and
    factors are found in both stages."""
    N = 1000000007 * _next_prime((1 << bits) + 12345)
    b1, b2, count = 3000, 300000, 40
    sig = [11 + i for i in range(count)]
    res = {}
    for kern in ("coop", "solo"):
        monkeypatch.setenv("ECM_B200_S2_KERNEL", kern)
        ctx = E.EcmContext(N, count)
        try:
            ctx.build_curves(sig); ctx.stage1(b1)
            ctx.stage2(b1, b2)
            res[kern] = ctx.read_stage2()
        finally:
            ctx.close()
    acc, f2, fail = res["coop"]
    sacc, sf2, sfail = res["solo"]
    assert fail == sfail, [i for i in range(count) if fail[i] != sfail[i]]
    assert f2 == sf2, [i for i in range(count) if f2[i] != sf2[i]]
    diff = [i for i in range(count) if acc[i] != sacc[i]]
    assert not diff, (diff, [fail[i] for i in diff])
    assert any(fail) and not all(fail) and any(f2)
    picks = [i for i in range(count) if fail[i]][:3] + [i for i in range(count) if not fail[i]][:3]
    for i in picks:
        o = O.ecm_curve(N, b1, b2, sig[i])
        assert (acc[i], f2[i]) == (o["acc"], o["f2"]), (i, fail[i])
