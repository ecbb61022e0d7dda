"""avx-ecm_b200: B200-native engine for avx-ecm's hot path (batched Montgomery arithmetic ->
stage 1 PRAC -> stage 2 pairing), reached through the C ABI in include/ecm_b200.h.

This module is the thin Python host layer over libecm_b200.so (ctypes).  It mirrors the
reference's driver vocabulary (vececm / ecm_stage1 / save_b1.txt lines, ecm.c:1077-1544) so that
tests read like the reference's own runs.  There is NO CPU fallback: if the CUDA library is
missing or no GPU is present the calls raise.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.realpath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("ECM_B200_LIB", "libecm_b200.so"))     # ECM_B200_LIB: A/B builds of the same library
_lib = None

u8p = ctypes.POINTER(ctypes.c_uint8)
u32p = ctypes.POINTER(ctypes.c_uint32)
u64p = ctypes.POINTER(ctypes.c_uint64)


class EcmError(RuntimeError):
    pass


def build(limbs=None, force=False):
    """Compile libecm_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile)."""
    cmd = ["make", "-C", _HERE]
    if limbs:
        cmd.append("LIMBS=" + limbs)
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    subprocess.run(cmd, check=True, capture_output=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EcmError("libecm_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                       "there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp = c.c_void_p
    sig = {
        "ecm_b200_create": (c.c_int, [c.POINTER(vp), c.c_int, u32p, c.c_int, c.c_uint32]),
        "ecm_b200_create_special": (c.c_int, [c.POINTER(vp), c.c_int, u32p, c.c_int, u32p, c.c_int, c.c_uint32]),
        "ecm_b200_uses_fold": (c.c_int, [vp]),
        "ecm_b200_destroy": (None, [vp]),
        "ecm_b200_last_error": (c.c_char_p, []),
        "ecm_b200_limbs": (c.c_int, [vp]),
        "ecm_b200_build_curves": (c.c_int, [vp, c.c_uint32, u64p]),
        "ecm_b200_load_curves": (c.c_int, [vp, c.c_uint32, u32p, u32p]),
        "ecm_b200_stage1": (c.c_int, [vp, c.c_uint64]),
        "ecm_b200_stage1_ranges": (c.c_int, [c.c_uint64, c.POINTER(c.c_uint32)]),
        "ecm_b200_stage1_range": (c.c_int, [vp, c.c_uint64, c.c_uint32, c.POINTER(c.c_uint64)]),
        "ecm_b200_stage1_begin": (c.c_int, [vp, c.c_uint64]),
        "ecm_b200_stage1_step": (c.c_int, [vp, c.c_uint32, c.POINTER(c.c_int)]),
        "ecm_b200_stage1_launches": (c.c_int, [vp, u32p, u32p]),
        "ecm_b200_sync": (c.c_int, [vp]),
        "ecm_b200_stage1_progress": (c.c_int, [vp, c.POINTER(c.c_double)]),
        "ecm_b200_flush_l2": (c.c_int, [vp]),
        "ecm_b200_timer": (c.c_int, [vp, c.c_int, c.POINTER(c.c_float)]),
        "ecm_b200_read_stage1": (c.c_int, [vp, u32p, u32p, u8p, u32p]),
        "ecm_b200_stage2": (c.c_int, [vp, c.c_uint64, c.c_uint64]),
        "ecm_b200_stage2_init": (c.c_int, [vp, c.c_uint64, c.POINTER(c.c_int)]),
        "ecm_b200_stage2_range": (c.c_int, [vp, c.c_uint32, u32p, u32p, c.c_uint32]),
        "ecm_b200_read_stage2": (c.c_int, [vp, u32p, u8p, u32p, u8p]),
        "ecm_b200_stage2_counters": (c.c_int, [vp, u64p, u64p, u64p, u64p]),
        "ecm_b200_plan_stage1": (c.c_uint64, [c.c_uint64, u8p, c.c_uint64, u64p]),
        "ecm_b200_plan_stage2": (c.c_uint64, [c.c_uint64, c.c_uint64, u64p]),
        "ecm_b200_stage2_program": (c.c_uint64, [c.c_uint64, c.c_uint64, c.c_int, u64p, c.c_uint64, u32p]),
        "ecm_b200_stage2_pairmap_program": (c.c_uint64, [c.c_uint64, c.c_uint32, u32p, u32p, c.c_uint32, u64p, c.c_uint64]),
        "ecm_b200_rv_program": (c.c_int, [c.c_int, u32p, c.c_int]),
        "ecm_b200_pair": (c.c_uint32, [c.c_uint64, c.c_uint64, c.c_uint32, c.c_uint32, u32p, u32p, c.c_uint32, u32p, u32p]),
        "ecm_b200_stage2_params": (None, [c.c_uint64, u32p, u32p, u32p, u32p]),
        "ecm_b200_fieldop": (c.c_int, [vp, c.c_int, c.c_uint32, u32p, u32p, u32p, c.c_int]),
        "ecm_b200_launch_count": (c.c_uint64, []),
        "ecm_b200_last_timing": (c.c_int, [vp, c.POINTER(c.c_float), u32p]),
        "ecm_b200_measure_imad_peak": (c.c_int, [c.c_int, c.POINTER(c.c_double), c.POINTER(c.c_double)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


EXPORTS = ["ecm_b200_create", "ecm_b200_create_special", "ecm_b200_uses_fold", "ecm_b200_destroy", "ecm_b200_last_error", "ecm_b200_limbs", "ecm_b200_build_curves",
           "ecm_b200_load_curves", "ecm_b200_stage1", "ecm_b200_stage1_ranges", "ecm_b200_stage1_range", "ecm_b200_stage1_begin", "ecm_b200_stage1_step",
           "ecm_b200_stage1_launches", "ecm_b200_sync", "ecm_b200_stage1_progress", "ecm_b200_flush_l2", "ecm_b200_timer", "ecm_b200_read_stage1", "ecm_b200_stage2",
           "ecm_b200_stage2_init", "ecm_b200_stage2_range",
           "ecm_b200_read_stage2", "ecm_b200_stage2_counters", "ecm_b200_plan_stage1", "ecm_b200_plan_stage2", "ecm_b200_stage2_program", "ecm_b200_stage2_pairmap_program", "ecm_b200_rv_program", "ecm_b200_pair", "ecm_b200_stage2_params",
           "ecm_b200_fieldop", "ecm_b200_launch_count", "ecm_b200_last_timing", "ecm_b200_measure_imad_peak"]


def _check(rc):
    if rc != 0:
        raise EcmError("ecm_b200 error %d: %s" % (rc, lib().ecm_b200_last_error().decode()))


# ---- limb packing: batch buffers are [limb][curve] uint32 (include/ecm_b200.h) ---------------------
def pack(values, nl):
    n = len(values)
    buf = (ctypes.c_uint32 * (nl * n))()
    for i, v in enumerate(values):
        for k in range(nl):
            buf[k * n + i] = (v >> (32 * k)) & 0xFFFFFFFF
    return buf


def unpack(buf, nl, n):
    out = []
    for i in range(n):
        v = 0
        for k in range(nl):
            v |= buf[k * n + i] << (32 * k)
        out.append(v)
    return out


def save_line(sigma, b1, N, x, z):
    """One save_b1.txt line, byte for byte what the reference appends (ecm.c:1372-1380)."""
    return "METHOD=ECM; SIGMA=%d; B1=%d; N=0x%x; X=0x%x; Z=0x%x; PROGRAM=AVX-ECM;\n" % (sigma, b1, N, x, z)


class EcmContext:
    """One batch of curves sharing N on one GPU (the analogue of the reference's thread_data_t[],
    avx_ecm.h:265-287; one context per GPU replaces one pthread per 8 lanes)."""

    def __init__(self, N, max_curves, device=0, base=None):
        """base: for a special-form input (see special_form) the number 2^k-1 / 2^k+1 / 2^k-c that N divides;
        the curve arithmetic is then done modulo base and only the factor checks use N (main.c:597-616)."""
        L = lib()
        self.N = N
        self.base = base

        def limbs(v):
            nl = max(1, (v.bit_length() + 31) // 32)
            return (ctypes.c_uint32 * nl)(*[(v >> (32 * k)) & 0xFFFFFFFF for k in range(nl)]), nl
        nbuf, nlimbs = limbs(N)
        self._h = ctypes.c_void_p()
        if base is None:
            _check(L.ecm_b200_create(ctypes.byref(self._h), device, nbuf, nlimbs, max_curves))
        else:
            bbuf, blimbs = limbs(base)
            _check(L.ecm_b200_create_special(ctypes.byref(self._h), device, bbuf, blimbs, nbuf, nlimbs, max_curves))
        self.nl = L.ecm_b200_limbs(self._h)
        self.uses_fold = bool(L.ecm_b200_uses_fold(self._h))     # shift-and-fold kernels instead of Montgomery
        self.max_curves = max_curves
        self.count = 0
        self.sigmas = []

    def close(self):
        if self._h:
            lib().ecm_b200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # phase 0 (build_one_curve, ecm.c:1548-1803)
    def build_curves(self, sigmas):
        s = (ctypes.c_uint64 * len(sigmas))(*sigmas)
        _check(lib().ecm_b200_build_curves(self._h, len(sigmas), s))
        self.count, self.sigmas = len(sigmas), list(sigmas)

    def load_curves(self, xs, ss, sigmas=None):
        _check(lib().ecm_b200_load_curves(self._h, len(xs), pack(xs, self.nl), pack(ss, self.nl)))
        self.count, self.sigmas = len(xs), list(sigmas or range(len(xs)))

    # phase 1 (ecm_stage1, ecm.c:1806-1854)
    def stage1(self, b1):
        _check(lib().ecm_b200_stage1(self._h, b1))
        self.b1 = b1

    def stage1_range(self, b1, r):
        """One 1e8 prime range of stage 1 (ecm.c:1207-1234); -> the last prime used (the B1 of the checkpoint line)."""
        last = ctypes.c_uint64()
        _check(lib().ecm_b200_stage1_range(self._h, b1, r, ctypes.byref(last)))
        return last.value

    def stage1_begin(self, b1):
        _check(lib().ecm_b200_stage1_begin(self._h, b1))
        self.b1 = b1

    def stage1_step(self, max_launches):
        done = ctypes.c_int(0)
        _check(lib().ecm_b200_stage1_step(self._h, max_launches, ctypes.byref(done)))
        return bool(done.value)

    def stage1_launches(self):
        t, i = ctypes.c_uint32(), ctypes.c_uint32()
        _check(lib().ecm_b200_stage1_launches(self._h, ctypes.byref(t), ctypes.byref(i)))
        return t.value, i.value

    def sync(self):
        _check(lib().ecm_b200_sync(self._h))

    def stage1_progress(self):
        f = ctypes.c_double()
        _check(lib().ecm_b200_stage1_progress(self._h, ctypes.byref(f)))
        return f.value

    def timer_start(self):
        _check(lib().ecm_b200_timer(self._h, 0, None))

    def timer_stop(self):
        _check(lib().ecm_b200_timer(self._h, 1, None))

    def timer_ms(self):
        ms = ctypes.c_float()
        _check(lib().ecm_b200_timer(self._h, 2, ctypes.byref(ms)))
        return ms.value

    def flush_l2(self):
        _check(lib().ecm_b200_flush_l2(self._h))

    def read_stage1(self):
        """-> (X, Z, factors): residues as written to save_b1.txt and gcd(Z,N) when it is a proper factor."""
        n, nl = self.count, self.nl
        x = (ctypes.c_uint32 * (nl * n))()
        z = (ctypes.c_uint32 * (nl * n))()
        g = (ctypes.c_uint32 * (nl * n))()
        fl = (ctypes.c_uint8 * n)()
        _check(lib().ecm_b200_read_stage1(self._h, x, z, fl, g))
        gs = unpack(g, nl, n)
        return unpack(x, nl, n), unpack(z, nl, n), [gs[i] if fl[i] else 0 for i in range(n)]

    # phases 2+3 (ecm_stage2_init / ecm_stage2_pair, ecm.c:2201-2540)
    def stage2(self, b1, b2):
        _check(lib().ecm_b200_stage2(self._h, b1, b2))

    def stage2_init(self, b1):
        """ecm_stage2_init (ecm.c:2201-2340) for the whole batch; -> foundDuringInv"""
        f = ctypes.c_int()
        _check(lib().ecm_b200_stage2_init(self._h, b1, ctypes.byref(f)))
        return bool(f.value)

    def stage2_range(self, amin, pm_v, pm_u):
        """ecm_stage2_pair (ecm.c:2342-2540) on a caller-supplied pairmap starting at window index amin"""
        n = len(pm_v)
        v = (ctypes.c_uint32 * max(1, n))(*pm_v)
        u = (ctypes.c_uint32 * max(1, n))(*pm_u)
        _check(lib().ecm_b200_stage2_range(self._h, amin, v, u, n))

    def read_stage2(self):
        n, nl = self.count, self.nl
        acc = (ctypes.c_uint32 * (nl * n))()
        g = (ctypes.c_uint32 * (nl * n))()
        fl = (ctypes.c_uint8 * n)()
        iv = (ctypes.c_uint8 * n)()
        _check(lib().ecm_b200_read_stage2(self._h, acc, fl, g, iv))
        gs = unpack(g, nl, n)
        return unpack(acc, nl, n), [gs[i] if fl[i] else 0 for i in range(n)], list(iv)

    def stage2_counters(self):
        v = [ctypes.c_uint64() for _ in range(4)]
        _check(lib().ecm_b200_stage2_counters(self._h, *[ctypes.byref(x) for x in v]))
        return dict(zip(("s2_ptadds", "s2_numinv", "s2_paired", "pairmap_steps"), [x.value for x in v]))

    def fieldop(self, op, a, b, repeat=1):
        n, nl = len(a), self.nl
        r = (ctypes.c_uint32 * (nl * n))()
        _check(lib().ecm_b200_fieldop(self._h, op, n, pack(a, nl), pack(b, nl), r, repeat))
        return unpack(r, nl, n)

    def last_timing(self):
        ms, ln = ctypes.c_float(), ctypes.c_uint32()
        _check(lib().ecm_b200_last_timing(self._h, ctypes.byref(ms), ctypes.byref(ln)))
        return ms.value, ln.value


def measure_imad_peak(device=0):
    """-> (32x32->64 products per second, SM clock MHz) of dependency-free IMAD.WIDE chains, measured now."""
    r, clk = ctypes.c_double(), ctypes.c_double()
    _check(lib().ecm_b200_measure_imad_peak(device, ctypes.byref(r), ctypes.byref(clk)))
    return r.value, clk.value


def stage1_ranges(b1):
    """Number of prime ranges the reference (and ecm_b200_stage1) splits stage 1 into: ceil(B1 / 1e8)."""
    n = ctypes.c_uint32()
    _check(lib().ecm_b200_stage1_ranges(b1, ctypes.byref(n)))
    return n.value


def plan_stage1(b1):
    """-> (ops bytes, point adds, point doublings) for B1 (prac planner, ecm.c:479-884,1815-1832)."""
    L = lib()
    cnt = (ctypes.c_uint64 * 2)()
    n = L.ecm_b200_plan_stage1(b1, None, 0, cnt)
    buf = (ctypes.c_uint8 * max(1, n))()
    L.ecm_b200_plan_stage1(b1, buf, n, cnt)
    return bytes(buf[:n]), cnt[0], cnt[1]


def plan_stage2(b1, b2):
    """Compile (not run) the stage-2 program; -> dict of the reference's counters + program length."""
    cnt = (ctypes.c_uint64 * 6)()
    n = lib().ecm_b200_plan_stage2(b1, b2, cnt)
    return {"instructions": n, "s2_ptadds": cnt[0], "s2_numinv": cnt[1], "s2_paired": cnt[2], "pairmap_steps": cnt[3],
            "last_amin": cnt[4], "table_entries": cnt[5]}


def stage2_program(b1, b2, which):
    """-> (list of 64-bit instruction words, layout dict) of the compiled stage-2 program (which = -1: init)."""
    L = lib()
    lay = (ctypes.c_uint32 * 13)()
    n = L.ecm_b200_stage2_program(b1, b2, which, None, 0, lay)
    buf = (ctypes.c_uint64 * max(1, n))()
    L.ecm_b200_stage2_program(b1, b2, which, buf, n, lay)
    names = ("npb", "pbx", "pbz", "pba", "pax", "paz", "pai", "paa", "qx", "qz", "pdx", "pdz", "entries")
    return list(buf[:n]), dict(zip(names, lay))


def rv_program(macro_op):
    """Phase words of one stage-1 macro-op of the register-resident kernel (rv.cuh)."""
    buf = (ctypes.c_uint32 * 16)()
    n = lib().ecm_b200_rv_program(macro_op, buf, 16)
    return list(buf[:n])


def stage2_pairmap_program(b1, amin, pm_v, pm_u):
    """-> instruction words of the program ecm_b200_stage2_range runs for a caller pairmap; [] when it is rejected."""
    L = lib()
    n = len(pm_v)
    v = (ctypes.c_uint32 * max(1, n))(*pm_v)
    u = (ctypes.c_uint32 * max(1, n))(*pm_u)
    k = L.ecm_b200_stage2_pairmap_program(b1, amin, v, u, n, None, 0)
    buf = (ctypes.c_uint64 * max(1, k))()
    L.ecm_b200_stage2_pairmap_program(b1, amin, v, u, n, buf, k)
    return list(buf[:k])


def pair(lo, hi, D, U=16):
    L = lib()
    amin, npairs = ctypes.c_uint32(), ctypes.c_uint32()
    steps = L.ecm_b200_pair(lo, hi, D, U, None, None, 0, ctypes.byref(amin), ctypes.byref(npairs))
    v = (ctypes.c_uint32 * max(1, steps))()
    u = (ctypes.c_uint32 * max(1, steps))()
    L.ecm_b200_pair(lo, hi, D, U, v, u, steps, ctypes.byref(amin), ctypes.byref(npairs))
    return list(v[:steps]), list(u[:steps]), amin.value, npairs.value


def stage2_params(b1):
    L = lib()
    D, U, Lw, R = (ctypes.c_uint32() for _ in range(4))
    L.ecm_b200_stage2_params(b1, ctypes.byref(D), ctypes.byref(U), ctypes.byref(Lw), ctypes.byref(R))
    return D.value, U.value, Lw.value, R.value


_SMALL_P = [p for p in range(2, 1000) if all(p % q for q in range(2, int(p ** 0.5) + 1))]


def primitive_part(k, sign):
    """The part of 2^k + sign that main.c:187-358 keeps (find_primitive_factor with base 2): with m = k divided
    by its distinct odd prime factors q_1..q_r, the alternating product over the subsets T of {q_i} of
    (2^(m*prod T) + sign)^((-1)^(r-|T|)).  Up to three distinct odd primes, like the reference."""
    odd = []
    e = k
    for q in _SMALL_P:
        while e % q == 0:
            e //= q
            if q & 1 and q not in odd:
                odd.append(q)
    if len(odd) > 3:
        raise ValueError("too many distinct odd factors in exponent")
    m = k
    for q in odd:
        m //= q
    num = den = 1
    for mask in range(1 << len(odd)):
        t, bits = 1, 0
        for i, q in enumerate(odd):
            if mask >> i & 1:
                t *= q
                bits += 1
        term = (1 << (m * t)) + sign
        if (len(odd) - bits) % 2 == 0:
            num *= term
        else:
            den *= term
    return num // den


def special_form(N, digitbits=52):
    """Input classification of main.c:405-521.  Returns dict(kind, k, c, base, n): kind 0 = generic REDC
    (base None), 1 = 2^k-1, -1 = 2^k+1, c > 1 = pseudo-Mersenne 2^k-c; n = N with the algebraic factors of
    2^k+-1 removed (main.c:444-457).  A form whose base is so much longer than N that REDC on N is cheaper
    (word ratio < 0.7, main.c:505-521) is reported as generic, like the reference does."""
    from math import gcd
    size_n = N.bit_length()
    kind, k = 0, size_n
    for i in range(size_n - 1, 2048):
        r = pow(2, i, N)
        if (r - 1) % N == 0:
            kind, k = 1, i
            break
        if (r + 1) % N == 0:
            kind, k = -1, i
            break
        if r.bit_length() < digitbits:
            if r.bit_length() < 32:           # the reference keeps c in an int (main.c:391,437)
                kind, k = r, i
            break
    n = N
    if abs(kind) == 1:
        n = gcd(N, primitive_part(k, -kind))
    block = 208 if digitbits == 52 else 128

    def maxbits(bits):
        mb = block
        while mb <= bits:
            mb += block
        return mb
    if kind and maxbits(n.bit_length()) / maxbits(k) < 0.7:
        kind = 0
    if kind == 0:
        return {"kind": 0, "k": n.bit_length(), "c": 0, "base": None, "n": n}
    base = (1 << k) - kind if kind > 0 else (1 << k) + 1
    return {"kind": kind, "k": k, "c": kind, "base": base, "n": n}


def vececm(N, curves, b1, b2=None, sigma=7, device=0, ctx=None, base=None):
    """Driver for one batch, mirroring vececm() (ecm.c:1077-1544) with threads=1 semantics:
    curve i runs sigma+i (main.c:757-763).  Returns dict(save_lines, factors=[(sigma, stage, factor)],
    x, z, acc).  base: see EcmContext (special-form inputs; X, Z, acc are then residues mod base)."""
    own = ctx is None
    if own:
        ctx = EcmContext(N, curves, device, base=base)
    try:
        sig = [sigma + i for i in range(curves)]
        ctx.build_curves(sig)
        ctx.stage1(b1)
        x, z, f1 = ctx.read_stage1()
        out = {"save_lines": [save_line(s, b1, N, xi, zi) for s, xi, zi in zip(sig, x, z)],
               "factors": [(s, 1, f) for s, f in zip(sig, f1) if f], "x": x, "z": z, "acc": None}
        if b2 is None:
            b2 = 100 * b1                       # main.c:462
        if b2 > b1:                             # main.c:548-552
            ctx.stage2(b1, b2)
            acc, f2, inv_fail = ctx.read_stage2()
            out["acc"], out["inv_fail"] = acc, inv_fail
            out["factors"] += [(s, 2, f) for s, f in zip(sig, f2) if f]
        return out
    finally:
        if own:
            ctx.close()
