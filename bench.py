#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 ECM engine (BASELINE.json: stage-1 curves/sec).

Workload (BASELINE.json configs[1]): synthetic 415-bit composite (13 x 32-bit limbs, the
reference's NWORDS=8 class), B1 = 1e6, 65 536 curves per GPU, stage 1 only, sigma = 7 + i.
Curves are independent, so N GPUs run N disjoint sigma ranges with no data-path collective
(weak scaling: 65 536 curves per GPU).

A *step* is one kernel launch of the stage-1 schedule: one pass of the field-op machine over
(up to) one resident wave of curve groups x one chunk of the PRAC op stream.  A full job is
`steps_per_job` such launches; with no --steps the timed region is exactly one full job, so the
headline number is "65 536 curves through the whole of stage 1".  With an explicit --steps K the
timed region is K launches (wrapping into a fresh batch if K exceeds one job) and the rate is
curves x (fraction of the job's field operations executed) / time.

  value     device-resident rate: CUDA events around the timed launches on the engine's stream
  e2e       the same job through the public C ABI with HOST buffers: sigmas in pinned/host memory
            -> curve construction -> stage 1 -> X, Z, factor flags back on the host (wall clock)
  roofline  algorithmic 32x32->64 products/s (modmuls x (2n^2+n)) against the IMAD.WIDE issue peak
            measured live on the same GPU (ecm_b200_measure_imad_peak)
  cpu_baseline / --impl reference   the unmodified reference (oracle/_ref/avx-ecm-ref, SKYLAKEX build)
            on all host threads, at the same B1 = 1e6 whenever the samples fit the time budget

`also` carries bounded samples of the other BASELINE configs, each with an oracle check and a roofline that counts
EVERY modular product of the compiled programs:
  config2_1024bit   [2] 1024-bit, B1 = 3e6: launches of the stage-1 schedule at 65 536 curves; stage 2 (B2 = 3e8
                    geometry) = ecm_stage2_init + the first pairmap steps of the first prime range on one wave
  config4_2048bit   [4] 2048-bit, B1 = 1.1e7: the same on the four-lanes-per-curve kernels
  sweep             [4]'s partitioning: ONE sigma range split over the ranks (strong scaling, avx_ecm_b200.dist.run_sharded),
                    merged on rank 0 in sigma order; `merged_sha256` must be the same at every N
Full-size runs of the other configs:  --config syn1024_s12 | syn2048_sweep  (minutes to hours; results under profiles/).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

B1 = 1000000
CURVES_PER_GPU = 65536
SIGMA0 = 7
COMPOSITE = "syn415"
S1_ADDS, S1_DUPS = 1980817, 217929                  # ecm.c:1849 printout for B1=1e6
MODMUL_PER_CURVE = 6 * S1_ADDS + 5 * S1_DUPS        # 12 974 547
METRIC = "stage1_curves_per_sec_B1_1e6_415bit"


PROFILE_SUMMARY = "profiles/r2_final_stage1_ncu_full_summary.txt"


def profile_traffic():
    """DRAM bytes per launch of the headline kernel, read from the committed `ncu --set full` summary (ncu prints the unit
    it likes per metric); None when the summary is absent."""
    try:
        txt = open(os.path.join(ROOT, PROFILE_SUMMARY)).read()
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            m = re.search(re.escape(name) + r" \[(\w+)\] = ([0-9.]+)", txt)
            total += float(m.group(2)) * scale[m.group(1)]
        return int(total)
    except Exception:
        return None


def composite():
    return int(json.load(open(os.path.join(ROOT, "tests", "golden", "composites.json")))[COMPOSITE])


# ------------------------------------------------------------------------------------------------
# clocks / throttle reasons during the timed region (NVML)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:          # noqa
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified avx-ecm on the host cores
# ------------------------------------------------------------------------------------------------
def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run_reference_sample(b1_sample, threads):
    """One bounded sample: 8 curves per thread on `threads` threads, stage 1 only at b1_sample.
    Returns (curves/s scaled to B1=1e6 by the exact modmul ratio, description)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "avx-ecm-ref")
    N = composite()
    if os.path.exists(ref) and "avx512f" in open("/proc/cpuinfo").read():
        with tempfile.TemporaryDirectory() as d:
            out = subprocess.run([ref, str(N), str(8 * threads), str(b1_sample), str(threads), str(b1_sample), str(SIGMA0)],
                                 cwd=d, capture_output=True, text=True, check=True).stdout
        t = float(re.search(r"Stage 1 took ([0-9.]+) seconds", out).group(1))
        m = re.search(r"with (\d+) point-adds and (\d+) point-doubles", out)
        modmul = 6 * int(m.group(1)) + 5 * int(m.group(2))
        rate = 8 * threads / t * modmul / MODMUL_PER_CURVE
        return rate, "reference", threads, ("avx-ecm-ref (SKYLAKEX build, AVX-512) %d threads x 8 curves, stage 1 at B1=%d "
                                            "(%.2f s)%s" % (threads, b1_sample, t, "" if b1_sample == B1 else
                                                            ", scaled to B1=1e6 by the exact modmul count"))
    # no AVX-512 host or no prebuilt reference: time the oracle port on one core
    import oracle_lib as O
    t0 = time.time()
    r = O.ecm_curve(N, b1_sample, b1_sample, SIGMA0)
    t = time.time() - t0
    modmul = 6 * r["counters"][0] + 5 * r["counters"][1]
    return (1.0 / t) * modmul / MODMUL_PER_CURVE, "port", 1, "oracle/ecm_oracle.c (GMP) 1 curve at B1=%d (%.2f s), scaled" % (b1_sample, t)


def reference_arm(args, rank):
    if rank != 0:
        return
    threads = host_threads()
    steps = args.steps if args.steps else 3
    warm = args.warmup if args.warmup is not None else 1
    # same config as the GPU arm (B1 = 1e6) whenever steps + warm-up samples fit ~4 minutes; the first sample is
    # timed to decide (it is a warm-up sample if any were asked for, else an extra one)
    t0 = time.time()
    first = run_reference_sample(B1, threads)
    t_full = time.time() - t0
    b1s = B1 if (steps + max(warm - 1, 0)) * t_full <= 240 else 100000
    for _ in range(max(warm - 1, 0)):
        run_reference_sample(b1s, threads)
    t0 = time.time()
    rates = []
    for _ in range(steps):
        rate, kind, cores, desc = run_reference_sample(b1s, threads)
        rates.append(rate)
    wall = time.time() - t0
    v = sum(rates) / len(rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "curves/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": wall / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u52-in-f64/u64 (AVX-512)",
        "data": "synthetic",
        "config": {"workload": "synthetic 415-bit composite, B1=1e6, stage 1 only, sigma=7.. (each step = one bounded sample)",
                   "composite": COMPOSITE, "b1": B1, "sample_b1": b1s, "same_config": b1s == B1,
                   "first_full_b1_sample": {"curves_per_sec": first[0], "seconds": t_full}},
        "cpu_baseline": {"value": v, "unit": "curves/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": v, "unit": "curves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------
# bounded samples of the other BASELINE configs (`also`) and the strong-scaling sweep sample
# ------------------------------------------------------------------------------------------------
def named_composite(name):
    return int(json.load(open(os.path.join(ROOT, "tests", "golden", "composites.json")))[name])


def program_modmuls(words):
    """Modular products (and inversions) in a compiled stage-2 program: V_MUL/V_SQR/V_PAIR = 1, V_MUL2 = 2 (plan2.hpp)."""
    mm = inv = pairs = 0
    for w in words:
        op = w & 0xff
        if op in (0, 1):
            mm += 1
        elif op == 10:
            mm += 1
            pairs += 1
        elif op == 12:
            mm += 2
        elif op == 8:
            inv += 1
    return mm, inv, pairs


def oracle_check(E, ctx, N, sig, b1, b2, picks):
    """Coherent small run on the SAME context (same kernels, batch geometry and layouts as the timed sample): stage 1 to
    b1, stage 2 to b2, curves `picks` compared with the oracle (test infrastructure; outside every timed region)."""
    import oracle_lib as O
    ctx.build_curves(sig)
    ctx.stage1(b1)
    x, z, _ = ctx.read_stage1()
    acc = None
    if b2 > b1:
        ctx.stage2(b1, b2)
        acc = ctx.read_stage2()[0]
    ok = True
    for i in picks:
        o = O.ecm_curve(N, b1, b2, sig[i])
        ok = ok and (x[i], z[i]) == (o["x"], o["z"]) and (acc is None or acc[i] == o["acc"])
    return bool(ok)


def config_sample(E, name, b1, curves_s1, curves_s2, launches, span, peak_prod, device, chk):
    """BASELINE config [2] / [4] as a bounded sample on one GPU.  Stage 1: `launches` launches of the schedule for b1
    on curves_s1 curves.  Stage 2 with the geometry of b1 (B2 = 100 b1): ecm_b200_stage2_init, then ecm_b200_stage2_range
    over the primes of [b1, b1 + span) -- the head of the first prime range -- on curves_s2 curves (one wave).
    chk = (b1, b2) of the coherent small run that is compared with the oracle on the same contexts."""
    N = named_composite(name)
    out = {"composite": name, "b1": b1, "b2": 100 * b1}
    _, adds, dups = E.plan_stage1(b1)
    modmul1 = 6 * adds + 5 * dups
    sig = [SIGMA0 + i for i in range(curves_s1)]
    ctx = E.EcmContext(N, curves_s1, device=device)
    try:
        nl = ctx.nl
        W = 2 * nl * nl + nl
        ctx.build_curves(sig); ctx.stage1_begin(b1)
        total, _ = ctx.stage1_launches()
        ctx.stage1_step(1); ctx.sync()
        ctx.build_curves(sig); ctx.stage1_begin(b1)
        ctx.timer_start(); ctx.stage1_step(launches); frac = ctx.stage1_progress(); ctx.timer_stop(); ctx.sync()
        ms = ctx.timer_ms()
        rate = curves_s1 * frac / (ms / 1e3)
        out["stage1"] = {"curves": curves_s1, "limbs": nl, "launches_timed": launches, "launches_per_full_job": total,
                         "ms_per_launch": ms / launches, "curves_per_sec": rate, "point_adds": adds, "point_doubles": dups,
                         "modmul_per_curve": modmul1, "products_per_sec": rate * modmul1 * W,
                         "frac_of_imad_peak": rate * modmul1 * W / peak_prod,
                         "residue_check_vs_oracle": oracle_check(E, ctx, N, sig, chk[0], chk[0], (0, curves_s1 - 1)),
                         "residue_check": "full stage 1 at B1=%d on this context, first and last curve" % chk[0]}
    finally:
        ctx.close()
    # stage 2
    D, U, L, R = E.stage2_params(b1)
    sig = [SIGMA0 + i for i in range(curves_s2)]
    ctx = E.EcmContext(N, curves_s2, device=device)
    try:
        ok = oracle_check(E, ctx, N, sig, chk[0], chk[1], (0, curves_s2 - 1))
        init_words, lay = E.stage2_program(b1, 100 * b1, -1)
        mm_init, inv_init, _ = program_modmuls(init_words)
        pm_v, pm_u, _, npairs = E.pair(b1, b1 + span, D)
        amin = (b1 + D) // (2 * D)
        mm_rng, inv_rng, pairs = program_modmuls(E.stage2_pairmap_program(b1, amin, pm_v, pm_u))
        ctx.build_curves(sig); ctx.stage1(chk[0])           # any point serves as Q
        t0 = time.time()
        ctx.stage2_init(b1)
        ms_init, l_init = ctx.last_timing()
        ctx.stage2_range(amin, pm_v, pm_u)
        ms_rng, l_rng = ctx.last_timing()
        wall = time.time() - t0
        acc = ctx.read_stage2()[0]
        out["stage2"] = {
            "curves": curves_s2, "D": D, "table_entries_per_curve": lay["entries"], "table_bytes": lay["entries"] * 4 * nl * curves_s2,
            "init": {"device_s": ms_init / 1e3, "launches": l_init, "modmul_per_curve": mm_init, "inversions_per_curve": inv_init,
                     "products_per_sec": curves_s2 * mm_init * W / (ms_init / 1e3), "frac_of_imad_peak": curves_s2 * mm_init * W / (ms_init / 1e3) / peak_prod},
            "range_head": {"primes": [b1, b1 + span], "pairmap_steps": len(pm_v), "pairs": npairs, "device_s": ms_rng / 1e3, "launches": l_rng,
                           "modmul_per_curve": mm_rng, "pair_products_per_curve": pairs, "inversions_per_curve": inv_rng,
                           "products_per_sec": curves_s2 * mm_rng * W / (ms_rng / 1e3),
                           "frac_of_imad_peak": curves_s2 * mm_rng * W / (ms_rng / 1e3) / peak_prod,
                           "table_read_GBs_algorithmic": 2 * pairs * 4 * nl * curves_s2 / (ms_rng / 1e3) / 1e9, "hbm_peak_GBs": 6540.8},
            "wall_s": wall, "accumulators_nontrivial": bool(len(set(acc[:64])) > 1),
            "note": "every modular product of the compiled programs is counted (products x (2n^2+n)); inversions are extra work, not counted",
            "check_vs_oracle": ok, "check": "stage 1 to %d + stage 2 to %d on this context, first and last curve" % chk}
    finally:
        ctx.close()
    return out


def sweep_sample(E, torch, dist, rank, world, device, total_curves, b1):
    """BASELINE config [4]'s partitioning as a bounded sample: ONE sigma range of total_curves 2048-bit curves split over the
    ranks (strong scaling), stage 1 at b1, save lines gathered on rank 0 in sigma order.  Returns the dict on rank 0."""
    import hashlib
    from avx_ecm_b200 import dist as D
    N = named_composite("syn2048")
    first, count = D.shard_range(total_curves, rank, world)
    ctx = E.EcmContext(N, max(count, 1), device=device)
    dev_ms = [0.0]

    def compute(first_sigma, cnt):
        r = E.vececm(N, cnt, b1, b2=b1, sigma=first_sigma, ctx=ctx)
        dev_ms[0] = ctx.last_timing()[0]
        return r
    try:
        E.plan_stage1(b1)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        merged = D.run_sharded(total_curves, SIGMA0, compute)
        if dist:
            dist.barrier()
        wall = time.time() - t0
    finally:
        ctx.close()
    wall = D.all_max(wall) if dist else wall
    dev = (D.all_max(dev_ms[0]) if dist else dev_ms[0]) / 1e3
    if rank != 0:
        return None
    import oracle_lib as O
    ok = all(merged["save_lines"][i] == O.ecm_curve(N, b1, b1, merged["sigmas"][i])["save_line"] for i in (0, total_curves - 1))
    _, adds, dups = E.plan_stage1(b1)
    return {"workload": "2048-bit, %d curves in ONE sigma range split over %d GPU(s), stage 1 at B1=%d, lines merged in sigma order" % (total_curves, world, b1),
            "scaling": "strong", "curves_total": total_curves, "n_gpus": world, "b1": b1,
            "stage1_device_s_max_over_ranks": dev, "wall_s_with_gather": wall,
            "curves_per_sec": total_curves / dev, "curves_per_sec_wall": total_curves / wall,
            "products_per_sec": total_curves / dev * (6 * adds + 5 * dups) * (2 * 64 * 64 + 64),
            "merged_lines": len(merged["save_lines"]), "merged_sha256": hashlib.sha256("".join(merged["save_lines"]).encode()).hexdigest(),
            "first_and_last_line_vs_oracle": bool(ok)}


def cli_multi_gpu_check(world):
    """The product's own multi-GPU path: avx-ecm-b200 with <world> GPUs (one host thread per GPU) on a golden case must
    write the reference's save_b1.txt byte for byte."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "syn415_b1_3e4_s1only.json")))
    cli = os.path.join(ROOT, "avx-ecm_b200", "avx-ecm-b200")
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([cli, g["n"], str(len(g["save_lines"])), str(g["b1"]), str(world), str(g["b2"]), g["sigma0"]],
                           cwd=d, capture_output=True, text=True)
        if r.returncode != 0:
            return {"error": (r.stdout + r.stderr)[-300:]}
        return {"case": "syn415_b1_3e4_s1only", "gpus": world, "identical": open(os.path.join(d, "save_b1.txt")).read() == "".join(g["save_lines"])}


def s2_program_modmuls(E, b1, b2):
    """(modular products, inversions, pair products) per curve of the whole compiled stage-2 program for (b1, b2)."""
    mm, inv, pairs = program_modmuls(E.stage2_program(b1, b2, -1)[0])
    which = 0
    while b1 + which * 100000000 < b2:
        a, b, c = program_modmuls(E.stage2_program(b1, b2, which)[0])
        mm, inv, pairs, which = mm + a, inv + b, pairs + c, which + 1
    return mm, inv, pairs


def full_config_run(args, E, torch, dist, rank, world, device):
    """--config syn1024_s12 / syn2048_sweep: a BASELINE config at full size (or with --b1 / --curves overrides), both stages,
    one sigma range split over the ranks, merged on rank 0.  One JSON line."""
    import hashlib
    from avx_ecm_b200 import dist as D
    if args.config == "syn1024_s12":
        name, b1, total = "syn1024", args.b1 or 3000000, args.curves * world
    else:
        name, b1, total = "syn2048", args.b1 or 11000000, (args.curves if args.curves != CURVES_PER_GPU else 8 * 14208)
    b2 = 100 * b1
    N = named_composite(name)
    first, count = D.shard_range(total, rank, world)
    peak_prod, peak_clk = E.measure_imad_peak(device)
    ctx = E.EcmContext(N, count, device=device)
    nl = ctx.nl
    W = 2 * nl * nl + nl
    tm = {}
    sampler = ClockSampler(device)
    L0 = E.lib().ecm_b200_launch_count()

    def compute(first_sigma, cnt):
        sig = [first_sigma + i for i in range(cnt)]
        t0 = time.time()
        ctx.build_curves(sig)
        ctx.stage1(b1)
        tm["s1_dev"] = ctx.last_timing()[0] / 1e3
        x, z, f1 = ctx.read_stage1()
        tm["s1_wall"] = time.time() - t0
        t0 = time.time()
        ctx.stage2(b1, b2)
        tm["s2_dev"] = ctx.last_timing()[0] / 1e3
        acc, f2, fail = ctx.read_stage2()
        tm["s2_wall"] = time.time() - t0
        tm["acc"] = (acc[0], acc[-1])
        return {"save_lines": [E.save_line(sg, b1, N, xi, zi) for sg, xi, zi in zip(sig, x, z)],
                "factors": [(sg, 1, f) for sg, f in zip(sig, f1) if f] + [(sg, 2, f) for sg, f in zip(sig, f2) if f]}
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.time()
    merged = D.run_sharded(total, SIGMA0, compute)
    if dist:
        dist.barrier()
    wall = time.time() - t0
    clocks = sampler.stop()
    launches = E.lib().ecm_b200_launch_count() - L0
    s1 = D.all_max(tm["s1_dev"]) if dist else tm["s1_dev"]
    s2 = D.all_max(tm["s2_dev"]) if dist else tm["s2_dev"]
    wall = D.all_max(wall) if dist else wall
    launches = int(D.all_sum(launches)) if dist else launches
    peak_all = D.all_sum(peak_prod) if dist else peak_prod
    ctx.close()
    if rank != 0:
        return
    import oracle_lib as O
    o = O.ecm_curve(N, b1, b2, SIGMA0)
    check = bool(merged["save_lines"][0] == o["save_line"] and tm["acc"][0] == o["acc"])
    _, adds, dups = E.plan_stage1(b1)
    mm1 = 6 * adds + 5 * dups
    mm2, inv2, pairs2 = s2_program_modmuls(E, b1, b2)
    line = {
        "metric": "curves_per_sec_stage1_plus_stage2_B1_%g_%dbit" % (b1, N.bit_length()), "value": total / (s1 + s2), "unit": "curves/s",
        "n_gpus": world, "steps": 1, "warmup": 0, "ms_per_step": (s1 + s2) * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.config == "syn2048_sweep" else "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "%s: %d-bit composite, B1=%d, B2=%d, %d curves in one sigma range over %d GPU(s), stage 1 + stage 2" %
                               (args.config, N.bit_length(), b1, b2, total, world), "composite": name, "b1": b1, "b2": b2, "curves_total": total,
                   "limbs": nl, "first_curve_vs_oracle": check},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": total / wall, "unit": "curves/s", "what": "sigmas on host -> both stages -> save lines, factors and accumulators on host, merged on rank 0 (wall clock)",
                "h2d_bytes_per_step": 32 * total, "d2h_bytes_per_step": (3 * nl * 4 + 2) * total},
        "stage1": {"device_s": s1, "curves_per_sec": total / s1, "point_adds": adds, "point_doubles": dups, "modmul_per_curve": mm1,
                   "products_per_sec": total / s1 * mm1 * W, "frac_of_imad_peak": total / s1 * mm1 * W / peak_all},
        "stage2": {"device_s": s2, "curves_per_sec": total / s2, "modmul_per_curve": mm2, "inversions_per_curve": inv2, "pair_products_per_curve": pairs2,
                   "products_per_sec": total / s2 * mm2 * W, "frac_of_imad_peak": total / s2 * mm2 * W / peak_all,
                   "table_read_GBs_algorithmic": 2 * pairs2 * 4 * nl * total / s2 / 1e9 / world, "hbm_peak_GBs": 6540.8,
                   "note": "every modular product of the compiled programs counted; inversions are extra work"},
        "roofline": {"bound": "imad", "achieved": total / (s1 + s2) * (mm1 + mm2) * W / 1e9, "peak": peak_all / 1e9, "unit": "Gprod/s",
                     "frac": total / (s1 + s2) * (mm1 + mm2) * W / peak_all, "traffic": None},
        "merged_lines": len(merged["save_lines"]), "merged_sha256": hashlib.sha256("".join(merged["save_lines"]).encode()).hexdigest(),
        "factors_found": len(merged["factors"]),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--curves", type=int, default=CURVES_PER_GPU, help="curves per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--config", default="syn415_s1", choices=["syn415_s1", "syn1024_s12", "syn2048_sweep"],
                    help="syn415_s1: the headline line (default); the others run a BASELINE config at full size")
    ap.add_argument("--b1", type=int, default=None, help="override B1 of a full-size config run")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import avx_ecm_b200 as E
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if args.config != "syn415_s1":
        full_config_run(args, E, torch, dist, rank, world, local_rank)
        if dist:
            dist.destroy_process_group()
        return

    N = composite()
    curves = args.curves
    sig = [SIGMA0 + rank * curves + i for i in range(curves)]
    ctx = E.EcmContext(N, curves, device=local_rank)
    nl = ctx.nl
    W = 2 * nl * nl + nl

    # job geometry
    ctx.build_curves(sig)
    ctx.stage1_begin(B1)
    steps_per_job, _ = ctx.stage1_launches()
    K = args.steps if args.steps else steps_per_job
    Wm = args.warmup if args.warmup is not None else 3
    Wm = max(Wm, 3)

    # warm-up launches (untimed), then a fresh batch
    ctx.stage1_step(Wm)
    ctx.sync()
    peak_prod, peak_clk = E.measure_imad_peak(local_rank)

    # ---- timed region: host buffers -> build -> K launches -> (read back if the job completed) ----
    sampler = ClockSampler(local_rank)
    L0 = E.lib().ecm_b200_launch_count()
    barrier()
    sampler.start()
    t_wall0 = time.time()
    ctx.build_curves(sig)                 # H2D of the curve seeds + construction kernel
    ctx.stage1_begin(B1)
    ctx.timer_start()
    done_frac, issued, full_jobs = 0.0, 0, 0
    while issued < K:
        ctx.flush_l2()
        fin = ctx.stage1_step(1)
        issued += 1
        if fin and issued < K:            # wrap into a fresh batch
            done_frac += 1.0
            full_jobs += 1
            ctx.build_curves(sig)
            ctx.stage1_begin(B1)
    done_frac += ctx.stage1_progress() if not fin else 1.0
    ctx.timer_stop()
    ctx.sync()
    dev_ms = ctx.timer_ms()
    complete = fin and full_jobs == 0
    x = z = f = None
    if complete:
        x, z, f = ctx.read_stage1()       # D2H of X, Z, factor flags
    t_wall = time.time() - t_wall0
    barrier()
    clocks = sampler.stop()
    launches = E.lib().ecm_b200_launch_count() - L0

    dev_s = allmax(dev_ms / 1e3)
    wall_s = allmax(t_wall)
    total_curve_equiv = allsum(curves * done_frac)
    value = total_curve_equiv / dev_s
    prod_rate = value * MODMUL_PER_CURVE * W          # algorithmic products / s over all GPUs
    peak_all = allsum(peak_prod)
    total_launches = int(allsum(launches))            # collective: every rank must take part

    # ---- e2e: the complete job through the C ABI with host buffers --------------------------------
    e2e = None
    h2d = 8 * 4 * curves                                # u,v seeds: 8 limbs per curve
    d2h = (2 * nl * 4 + 1) * curves + nl * 4 * curves   # X, Z, flags, gcd words
    if complete:
        e2e_rate = allsum(curves) / wall_s
    elif not args.no_e2e:
        barrier()
        t0 = time.time()
        ctx.build_curves(sig)
        ctx.stage1(B1)
        x, z, f = ctx.read_stage1()
        e2e_wall = allmax(time.time() - t0)
        e2e_rate = allsum(curves) / e2e_wall
    else:
        e2e_rate = None
    if e2e_rate is not None:
        e2e = {"value": e2e_rate, "unit": "curves/s", "h2d_bytes_per_step": h2d / steps_per_job, "d2h_bytes_per_step": d2h / steps_per_job,
               "h2d_bytes_per_job": h2d, "d2h_bytes_per_job": d2h,
               "what": "sigmas on host -> ecm_b200_build_curves -> ecm_b200_stage1 -> ecm_b200_read_stage1 (X, Z, gcd flags on host), wall clock"}

    # sanity: the job's output must be the reference's residues (checked against the oracle on one curve)
    check = None
    if x is not None and rank == 0:
        import oracle_lib as O
        check = True
        for i in (0, curves - 1):                      # first and last curve of the full-size job
            o = O.ecm_curve(N, B1, B1, sig[i])
            check = check and bool(o["x"] == x[i] and o["z"] == z[i])
        if not check:
            raise SystemExit("bench.py: stage-1 residues differ from the oracle")

    # ---- the other BASELINE configs as bounded samples (rank 0) ----------------------------------------------------
    also = None
    if rank == 0 and not args.no_e2e:
        also = {}
        try:    # [2] 1024-bit, B1 = 3e6, 65 536 curves (stage 2: one wave of 32 768, tables 2.98 MB/curve)
            also["config2_1024bit"] = config_sample(E, "syn1024", 3000000, curves, 32768, 3, 2000000, peak_prod, local_rank, (5000, 300000))
        except Exception as e:          # a side measurement must never take the headline line down
            also["config2_1024bit"] = {"error": repr(e)}
        try:    # [4] 2048-bit, B1 = 1.1e7: one resident wave of the four-lanes-per-curve kernels (148 blocks x 96 curves), both stages
            also["config4_2048bit"] = config_sample(E, "syn2048", 11000000, 14208, 14208, 3, 1000000, peak_prod, local_rank, (2000, 60000))
        except Exception as e:
            also["config4_2048bit"] = {"error": repr(e)}
        # stage-2 sample on the bench composite: B1=1e5 -> B2=1e7 (D=2310, U=16), all curves of this GPU
        sb1, sb2 = 100000, 10000000
        ctx.build_curves(sig)
        ctx.stage1(sb1)
        t0 = time.time()
        ctx.stage2(sb1, sb2)
        s2_wall = time.time() - t0
        s2_ms, s2_launches = ctx.last_timing()
        cnt = ctx.stage2_counters()
        mm2, inv2, _ = s2_program_modmuls(E, sb1, sb2)
        table_bytes = (2 * cnt["s2_paired"]) * 4 * nl * curves          # Pa_inv + Pb operand of every pair step
        also["stage2_sample"] = {"b1": sb1, "b2": sb2, "curves": curves, "device_s": s2_ms / 1e3, "wall_s": s2_wall,
                                 "curves_per_sec": curves / (s2_ms / 1e3), "kernel_launches": s2_launches,
                                 "pair_steps": cnt["s2_paired"], "point_adds": cnt["s2_ptadds"], "inversions": cnt["s2_numinv"],
                                 "modmul_per_curve": mm2, "inversions_per_curve": inv2,
                                 "products_per_sec": curves * mm2 * W / (s2_ms / 1e3), "frac_of_imad_peak": curves * mm2 * W / (s2_ms / 1e3) / peak_prod,
                                 "table_read_GBs_algorithmic": table_bytes / (s2_ms / 1e3) / 1e9,
                                 "hbm_peak_GBs": 6540.8}
        # special-form input (N | 2^415-1): the shift-and-fold kernels next to the Montgomery kernels on the same base
        try:
            sbase, sfb1 = (1 << 415) - 1, 30000
            sf = {"base": "2^415-1", "b1": sfb1, "curves": curves}
            ref_xz = None
            for mode in ("fold", "montgomery"):
                if mode == "montgomery":
                    os.environ["ECM_B200_NO_FOLD"] = "1"
                c3 = E.EcmContext(sbase, curves, device=local_rank, base=sbase)
                try:
                    best = None
                    for _ in range(2):
                        c3.build_curves(sig)
                        c3.stage1(sfb1)
                        ms3, _l = c3.last_timing()
                        best = ms3 if best is None else min(best, ms3)
                    xz = c3.read_stage1()[:2]
                    sf[mode + "_curves_per_sec"] = curves / (best / 1e3)
                    sf["uses_fold_" + mode] = c3.uses_fold
                finally:
                    c3.close()
                if ref_xz is None:
                    ref_xz = xz
                sf["identical_residues"] = bool(xz == ref_xz)
            sf["speedup"] = sf["fold_curves_per_sec"] / sf["montgomery_curves_per_sec"]
            also["special_form_sample"] = sf
        except Exception as e:          # a side measurement must never take the headline line down
            also["special_form_sample"] = {"error": repr(e)}
        finally:
            os.environ.pop("ECM_B200_NO_FOLD", None)

    # ---- strong-scaling sweep sample (every rank takes part) and the CLI's own multi-GPU path -----------------------
    sweep = None
    if not args.no_e2e:
        try:
            sweep = sweep_sample(E, torch, dist, rank, world, local_rank, 8 * 14208, 10000)
        except Exception as e:
            sweep = {"error": repr(e)}
        if rank == 0 and also is not None:
            also["sweep"] = sweep
            if world > 1:
                also["cli_multi_gpu"] = cli_multi_gpu_check(world)
        barrier()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, kind, cores, desc = run_reference_sample(B1, host_threads())
        cpu_baseline = {"value": rate, "unit": "curves/s", "cores": cores, "kind": kind, "sample": desc}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "curves/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": dev_s * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": "synthetic 415-bit composite (13x32-bit limbs), B1=1e6, %d curves per GPU, stage 1 only, sigma=7+i" % curves,
                       "composite": COMPOSITE, "b1": B1, "curves_per_gpu": curves, "limbs": nl,
                       "step": "one kernel launch of the stage-1 schedule", "steps_per_full_job": steps_per_job,
                       "job_fraction_timed": done_frac, "l2": "256 MiB memset between steps (inside the timed region)",
                       "residue_check_vs_oracle": check},
            "clocks": clocks, "e2e": e2e, "gpu_launches": total_launches,
            "roofline": {"bound": "imad", "achieved": prod_rate / 1e9, "peak": peak_all / 1e9, "unit": "Gprod/s",
                         "frac": prod_rate / peak_all,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel, read from the committed
                         # ncu summary of the same command (the batch state re-read after the L2 flush + write-backs;
                         # incidental for an IMAD-bound kernel); null when that file is absent
                         "traffic": profile_traffic(), "traffic_source": PROFILE_SUMMARY,
                         "note": "achieved = curves/s x 12974547 modmul/curve x (2n^2+n) products, n=%d; peak = fastest of three live probes on this GPU "
                                 "(IMAD.WIDE.U32.X chains with uniform / per-thread multiplier, register-resident 32-limb Montgomery loop) "
                                 "at %.0f MHz; the pipe's arithmetic ceiling is 32 products/clk/SM" % (nl, peak_clk)},
            "cpu_baseline": cpu_baseline,
            "also": also,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
