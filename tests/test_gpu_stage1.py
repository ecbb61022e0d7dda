"""GPU parity of curve construction + stage 1 against the golden vectors of the compiled
reference (byte-identical save_b1.txt lines) and against the oracle, through the C ABI."""
import pytest
from conftest import GOLDEN, golden_factor, golden_base, composites, is_known_answer
import oracle_lib as O
import avx_ecm_b200 as E

pytestmark = pytest.mark.gpu

MAXBITS = 2048


def usable(g):
    # B1 > 1e8: hours for 8 curves on one warp; full-size known answers run concurrently in test_gpu_known_answers.py
    return int(g["n"]).bit_length() <= MAXBITS and g["b1"] <= 100000000 and not is_known_answer(g["name"])


@pytest.mark.parametrize("name", sorted(k for k, g in GOLDEN.items() if usable(g)))
def test_stage1_matches_reference_golden(name):
    g = GOLDEN[name]
    N, b1 = int(g["n"]), g["b1"]
    lanes = len(g["save_lines"])
    r = E.vececm(N, lanes, b1, b2=b1, sigma=int(g["sigma0"]), base=golden_base(g))
    assert r["save_lines"] == g["save_lines"]
    assert r["z"] == [int(z, 16) for z in g["z1_true_hex"]]
    exp = [(int(g["sigma0"]) + i, 1, golden_factor(g, int(g["sigma0"]) + i, 1)) for i in range(lanes)]
    assert r["factors"] == [e for e in exp if e[2]]


def test_curve_construction_with_failed_inversions_matches_oracle():
    """Special-form inputs keep the small algebraic factors of 2^k+-1 in the arithmetic modulus, so the two
    mpz_invert calls of build_one_curve (ecm.c:1745,1759) fail routinely and leave their outputs untouched.
    11 | base: sigma = 15 makes u = sigma^2-5 non-invertible (v still is), sigma = 22 makes v non-invertible."""
    base = 11 * ((1 << 200) + 235)
    sig = [7, 15, 22, 26, 33, 1000003, 2 ** 63 + 11, 37]
    ctx = E.EcmContext(base, len(sig), base=base)
    try:
        ctx.build_curves(sig)
        x, z, _ = ctx.read_stage1()
        ctx.stage1(500)
        x1, z1, f1 = ctx.read_stage1()
    finally:
        ctx.close()
    for i, s in enumerate(sig):
        ox, _ = O.build_curve(base, s)
        assert (x[i], z[i]) == (ox, 1), "sigma %d" % s
        o = O.ecm_curve(base, 500, 500, s, M=base)
        assert (x1[i], z1[i], f1[i]) == (o["x"], o["z"], o["f1"]), "sigma %d" % s


def test_special_base_2k_plus_1_divisible_by_3():
    # (2^523+1)/3: every third sigma has a non-invertible v^3 (3 | 4 sigma); residues are reported mod 2^523+1
    base = 2 ** 523 + 1
    N = base // 3
    sig = [1000003 + i for i in range(12)]
    r = E.vececm(N, len(sig), 300, b2=300, sigma=sig[0], base=base)
    for i, s in enumerate(sig):
        o = O.ecm_curve(N, 300, 300, s, M=base)
        assert r["save_lines"][i] == o["save_line"]
        assert r["z"][i] == o["z"]


def test_curve_construction_matches_oracle():
    c = composites()
    for N in (c["syn415"], c["syn1024"], c["t35"], c["small96"]):
        sig = [6, 7, 1000, 2 ** 32 - 1, 2 ** 32 + 5, 2 ** 63 + 11, 2 ** 64 - 1, 11919771003873180376]
        ctx = E.EcmContext(N, len(sig))
        try:
            ctx.build_curves(sig)
            x, z, _ = ctx.read_stage1()          # before stage 1: the initial point
            for s, xi, zi in zip(sig, x, z):
                ox, _ = O.build_curve(N, s)
                assert (xi, zi) == (ox, 1)
        finally:
            ctx.close()


def test_host_loaded_curves_and_ragged_batches():
    # load_curves path (host-built X, s) and batch sizes that do not fill a block / span blocks
    N = composites()["syn415"]
    for count in (1, 33, 321, 700):
        sig = [100 + i for i in range(count)]
        built = [O.build_curve(N, s) for s in sig]
        ctx = E.EcmContext(N, count)
        try:
            ctx.load_curves([b[0] for b in built], [b[1] for b in built], sig)
            ctx.stage1(2000)
            x1, z1, _ = ctx.read_stage1()
            ctx.build_curves(sig)
            ctx.stage1(2000)
            x2, z2, _ = ctx.read_stage1()
        finally:
            ctx.close()
        assert (x1, z1) == (x2, z2)
        for i in (0, count // 2, count - 1):
            r = O.ecm_curve(N, 2000, 2000, sig[i])
            assert (x1[i], z1[i]) == (r["x"], r["z"])


def test_time_sliced_stage1_equals_one_shot():
    N = composites()["syn415"]
    sig = list(range(7, 7 + 64))
    ctx = E.EcmContext(N, len(sig))
    try:
        ctx.build_curves(sig); ctx.stage1(50000); a = ctx.read_stage1()
        ctx.build_curves(sig); ctx.stage1_begin(50000)
        while not ctx.stage1_step(1):
            pass
        ctx.sync(); b = ctx.read_stage1()
    finally:
        ctx.close()
    assert a == b


def test_error_behaviour_of_the_abi():
    """Misuse returns an error code with a message instead of crashing (the reference exit(1)s)."""
    N = composites()["syn415"]
    ctx = E.EcmContext(N, 64)
    try:
        with pytest.raises(E.EcmError, match="no curves"):
            ctx.stage1(1000)
        with pytest.raises(E.EcmError, match="sigma must be >= 6"):
            ctx.build_curves([5])
        with pytest.raises(E.EcmError, match="exceeds"):
            ctx.build_curves(list(range(10, 10 + 65)))
        ctx.build_curves([7, 8, 9])
        with pytest.raises(E.EcmError, match="B1 out of range"):
            ctx.stage1(2000000001)          # beyond the 2e9 cap (ecm_b200.h)
        with pytest.raises(E.EcmError, match="not been run"):
            ctx.read_stage2()
        ctx.stage1(1000)
        with pytest.raises(E.EcmError, match="already run"):
            ctx.stage1(1000)
        with pytest.raises(E.EcmError, match="B2 must exceed B1"):
            ctx.stage2(1000, 1000)
        # a fresh batch on the same context works after the errors
        ctx.build_curves([7, 8, 9]); ctx.stage1(1000)
        x, z, _ = ctx.read_stage1()
        o = O.ecm_curve(N, 1000, 1000, 8)
        assert (x[1], z[1]) == (o["x"], o["z"])
    finally:
        ctx.close()


def test_results_do_not_depend_on_block_size_or_schedule(monkeypatch):
    """Size-independent property: the same batch run with different curve-group sizes (hence different
    memory layouts, wave counts and launch schedules) must give identical residues for every curve."""
    N = composites()["syn415"]
    sig = list(range(500, 500 + 1500))
    res = []
    for threads in (None, "128", "320"):
        if threads:
            monkeypatch.setenv("ECM_B200_THREADS", threads)
        else:
            monkeypatch.delenv("ECM_B200_THREADS", raising=False)
        ctx = E.EcmContext(N, len(sig))
        try:
            ctx.build_curves(sig); ctx.stage1(20000)
            res.append(ctx.read_stage1())
        finally:
            ctx.close()
    assert res[0] == res[1] == res[2]
    o = O.ecm_curve(N, 20000, 20000, sig[1234])
    assert (res[0][0][1234], res[0][1][1234]) == (o["x"], o["z"])


@pytest.mark.parametrize("b1", [2, 3, 4, 5, 6, 10])
def test_tiny_b1_edge_cases(b1):
    # B1 = 2 has an empty op stream, 3..4 only the power-of-two doublings (ecm.c:1815-1832)
    N = composites()["syn415"]
    ctx = E.EcmContext(N, 4)
    try:
        ctx.build_curves([7, 8, 9, 10]); ctx.stage1(b1)
        x, z, _ = ctx.read_stage1()
    finally:
        ctx.close()
    for i in range(4):
        o = O.ecm_curve(N, b1, b1, 7 + i)
        assert (x[i], z[i]) == (o["x"], o["z"])


def test_stage1_range_by_range_matches_oracle_checkpoints(monkeypatch):
    """B1 beyond one prime range (1e8 in the reference; 3000 here through the test hooks): the reference calls
    ecm_stage1 per range -- repeated doublings, first prime of later ranges skipped -- and saves the point after
    each range to checkpoint.txt (ecm.c:1207-1311).  Range-by-range and one-go runs against the oracle."""
    monkeypatch.setenv("ECM_B200_S1_RANGE", "3000")
    O.set_prime_range(3000)
    try:
        N, b1 = composites()["syn415"], 10000
        sig = [7, 1000003, 2 ** 63 + 11] + list(range(50, 53))
        assert E.stage1_ranges(b1) == 4
        ctx = E.EcmContext(N, len(sig))
        try:
            ctx.build_curves(sig)
            with pytest.raises(E.EcmError):
                ctx.stage1_range(b1, 1)                  # ranges must be taken in order
            lasts = []
            for r in range(4):
                lasts.append(ctx.stage1_range(b1, r))
                x, z, _ = ctx.read_stage1()
                O.set_checkpoint(r + 1)
                for i, s in enumerate(sig):
                    o = O.ecm_curve(N, b1, b1, s)
                    assert (x[i], z[i]) == (o["x"], o["z"]), (r, s)
            assert lasts == [2999, 5987, 8999, 9973]
            with pytest.raises(E.EcmError):
                ctx.stage1(b1)                           # stage 1 is complete
            O.set_checkpoint(0)
            ctx.build_curves(sig)
            ctx.stage1(b1)                               # the same in one call
            assert ctx.read_stage1()[:2] == (x, z)
            ctx.build_curves(sig)
            ctx.stage1_range(b1, 0)
            with pytest.raises(E.EcmError):
                ctx.stage1(b1)                           # not after a range-by-range start
        finally:
            ctx.close()
    finally:
        O.set_prime_range(0)
        O.set_checkpoint(0)


def test_cli_writes_checkpoints(tmp_path, monkeypatch):
    import os, subprocess
    from conftest import ROOT
    monkeypatch.setenv("ECM_B200_S1_RANGE", "2000")
    O.set_prime_range(2000)
    try:
        N, b1, s0 = composites()["t35"], 5000, 424242
        cli = os.path.join(ROOT, "avx-ecm_b200", "avx-ecm-b200")
        subprocess.run([cli, str(N), "8", str(b1), "1", str(b1), str(s0)], cwd=tmp_path, capture_output=True, text=True, check=True)
        ck = (tmp_path / "checkpoint.txt").read_text().splitlines(keepends=True)
        assert len(ck) == 16                            # 3 ranges -> 2 checkpoints of 8 curves
        for r, last in ((1, 1999), (2, 3989)):
            O.set_checkpoint(r)
            for i in range(8):
                o = O.ecm_curve(N, b1, b1, s0 + i)
                assert ck[(r - 1) * 8 + i] == E.save_line(s0 + i, last, N, o["x"], o["z"])
        O.set_checkpoint(0)
        save = (tmp_path / "save_b1.txt").read_text().splitlines(keepends=True)
        assert save == [O.ecm_curve(N, b1, b1, s0 + i)["save_line"] for i in range(8)]
    finally:
        O.set_prime_range(0)
        O.set_checkpoint(0)


def test_gpu_lines_are_valid_gmp_ecm_resume_points():
    """Independent of the oracle: the (X:Z) the engine records must be [k]P0 on the Suyama curve of SIGMA
    (plain Montgomery ladder on Python integers, tests/test_resume_lines_cpu.py) -- what `ecm -resume` continues."""
    import random
    from test_resume_lines_cpu import check_line
    rng = random.Random(99)
    for name, b1 in (("syn415", 20000), ("syn1024", 3000), ("t35", 50000)):
        N = composites()[name]
        r = E.vececm(N, 12, b1, b2=b1, sigma=rng.randrange(6, 2 ** 63))
        assert [check_line(l, expect_n=N) for l in r["save_lines"]] == ["ok"] * 12
    base = (1 << 277) - 1
    r = E.vececm(base, 12, 5000, b2=5000, sigma=rng.randrange(6, 2 ** 63), base=base)
    assert [check_line(l, expect_n=base, base=base) for l in r["save_lines"]] == ["ok"] * 12


@pytest.mark.parametrize("name,curves,b1", [("small96", 70, 3000), ("syn206", 70, 3000), ("t35", 200, 3000), ("syn415", 500, 5000),
                                            ("readme508", 200, 3000), ("csh_line19", 100, 3000), ("csh_line02", 100, 3000),
                                            ("slow_csh_line07", 100, 1500), ("syn2048", 150, 1000), ("syn880", 100, 2000),
                                            ("syn1250", 100, 1500), ("syn1750", 100, 1000)])
def test_register_machine_and_slot_machine_kernels_agree(name, curves, b1, monkeypatch):
    """Stage 1 has two kernel generations: the slot-file machine (vm.cuh, one thread per curve) and the register-resident
    macro-op machine (rv.cuh) -- one thread per curve up to 16 limbs, FOUR LANES per curve (warp-cooperative limbs,
    coop.cuh) at 48 and 64 limbs.  Same op stream, different state layouts and arithmetic routines: every residue of every
    curve must be identical, and equal to the oracle's (3, 7, 10, 13, 16, 20, 24, 28, 40, 48, 56 and 64 limbs)."""
    N = composites()[name] if name in composites() else int(GOLDEN[name]["n"])
    sig = [1000 + 3 * i for i in range(curves)]
    res = {}
    for kern in ("vm", "rv"):
        monkeypatch.setenv("ECM_B200_S1_KERNEL", kern)
        ctx = E.EcmContext(N, curves)
        try:
            ctx.build_curves(sig); ctx.stage1(b1)
            res[kern] = ctx.read_stage1()
        finally:
            ctx.close()
    assert res["vm"] == res["rv"]
    for i in (0, curves // 3, curves - 1):
        o = O.ecm_curve(N, b1, b1, sig[i])
        assert (res["rv"][0][i], res["rv"][1][i]) == (o["x"], o["z"])


def test_cooperative_kernel_with_ragged_batches_and_stage2(monkeypatch):
    """2048-bit curves on the four-lanes-per-curve kernel: batch sizes that do not fill a warp's eight curves or a block,
    time-sliced launches, and stage 2 (one thread per curve) picking the points up from the cooperative layout."""
    N = composites()["syn2048"]
    for count in (1, 7, 9, 97, 200):
        sig = [50 + i for i in range(count)]
        ctx = E.EcmContext(N, count)
        try:
            ctx.build_curves(sig); ctx.stage1_begin(700)
            while not ctx.stage1_step(1):
                pass
            ctx.sync()
            x, z, _ = ctx.read_stage1()
            acc = None
            if count == 9:
                ctx.stage2(700, 30000)
                acc = ctx.read_stage2()[0]
        finally:
            ctx.close()
        for i in {0, count // 2, count - 1}:
            o = O.ecm_curve(N, 700, 30000 if acc else 700, sig[i])
            assert (x[i], z[i]) == (o["x"], o["z"]), (count, i)
            if acc:
                assert acc[i] == o["acc"]
