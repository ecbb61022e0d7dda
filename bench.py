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
            on all host threads
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
    b1s = B1 if (steps + warm) * 8.0 <= 180 else 100000
    for _ in range(warm):
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
                   "composite": COMPOSITE, "b1": B1, "sample_b1": b1s},
        "cpu_baseline": {"value": v, "unit": "curves/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": v, "unit": "curves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--curves", type=int, default=CURVES_PER_GPU, help="curves per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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

    # ---- the metric's second operand size: 1024-bit N, same B1, a few launches of the same schedule ----
    also = None
    if rank == 0 and not args.no_e2e:
        N2 = int(json.load(open(os.path.join(ROOT, "tests", "golden", "composites.json")))["syn1024"])
        c2 = E.EcmContext(N2, curves, device=local_rank)
        c2.build_curves(sig)
        c2.stage1_begin(B1)
        tot2, _ = c2.stage1_launches()
        c2.stage1_step(1); c2.sync()
        c2.build_curves(sig); c2.stage1_begin(B1)
        c2.timer_start(); c2.stage1_step(3); frac2 = c2.stage1_progress(); c2.timer_stop(); c2.sync()
        ms2 = c2.timer_ms()
        v2 = curves * frac2 / (ms2 / 1e3)
        W2 = 2 * c2.nl * c2.nl + c2.nl
        also = {"stage1_curves_per_sec_B1_1e6_1024bit": v2, "limbs": c2.nl, "launches_timed": 3, "launches_per_full_job": tot2,
                "products_per_sec": v2 * MODMUL_PER_CURVE * W2, "frac_of_imad_peak": v2 * MODMUL_PER_CURVE * W2 / peak_prod,
                "composite": "syn1024", "curves": curves}
        c2.close()
        # stage-2 sample on the bench composite: B1=1e5 -> B2=1e7 (D=2310, U=16), all curves of this GPU
        sb1, sb2 = 100000, 10000000
        ctx.build_curves(sig)
        ctx.stage1(sb1)
        t0 = time.time()
        ctx.stage2(sb1, sb2)
        s2_wall = time.time() - t0
        s2_ms, s2_launches = ctx.last_timing()
        cnt = ctx.stage2_counters()
        table_bytes = (2 * cnt["s2_paired"]) * 4 * nl * curves          # Pa_inv + Pb operand of every pair step
        also["stage2_sample"] = {"b1": sb1, "b2": sb2, "curves": curves, "device_s": s2_ms / 1e3, "wall_s": s2_wall,
                                 "curves_per_sec": curves / (s2_ms / 1e3), "kernel_launches": s2_launches,
                                 "pair_steps": cnt["s2_paired"], "point_adds": cnt["s2_ptadds"], "inversions": cnt["s2_numinv"],
                                 "products_per_sec_adds_and_pairs": curves * (6 * cnt["s2_ptadds"] + cnt["s2_paired"]) * W / (s2_ms / 1e3),
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
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch (33.47 + 56.40 MB: the batch state is
                         # re-read after the L2 flush between steps and dirty lines are written back),
                         # profiles/r1_final_stage1_ncu_full_summary.txt; incidental for an IMAD-bound kernel
                         "traffic": 89868544,
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
