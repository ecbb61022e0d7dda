"""ctypes binding of the CPU oracle (oracle/ecm_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes, os, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "libecm_oracle.so")

_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "_build/libecm_oracle.so", "_build/ecm_oracle"], check=True)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ORACLE_DIR, "ecm_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            build()
        L = ctypes.CDLL(LIB)
        u64, u32, cp = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_char_p
        L.oracle_ecm_curve.argtypes = [cp, u64, u64, u64, cp, cp, cp, cp, cp, ctypes.POINTER(u32)]
        L.oracle_ecm_curve.restype = ctypes.c_int
        L.oracle_ecm_curve_special.argtypes = [cp, cp, u64, u64, u64, cp, cp, cp, cp, cp, ctypes.POINTER(u32)]
        L.oracle_ecm_curve_special.restype = ctypes.c_int
        L.oracle_set_prime_range.argtypes = [u64]
        L.oracle_set_checkpoint.argtypes = [ctypes.c_int]
        L.oracle_build_curve.argtypes = [cp, u64, cp, cp]
        L.oracle_save_line.argtypes = [cp, u64, u64, cp, cp, cp, ctypes.c_size_t]
        L.oracle_pair.argtypes = [u64, u64, u32, u32, ctypes.POINTER(u32), ctypes.POINTER(u32), u32,
                                  ctypes.POINTER(u32), ctypes.POINTER(u32)]
        L.oracle_pair.restype = u32
        L.oracle_stage2_D.argtypes = [u64]
        L.oracle_stage2_D.restype = u32
        L.oracle_stage2_map.argtypes = [u64, ctypes.POINTER(u32), u32]
        L.oracle_stage2_map.restype = u32
        L.oracle_stage1_trace.argtypes = [u64, ctypes.c_char_p, u64]
        L.oracle_stage1_trace.restype = u64
        L.oracle_prac_best.argtypes = [u64]
        L.oracle_prac_best.restype = ctypes.c_int
        _lib = L
    return _lib


def ecm_curve(N, b1, b2, sigma, M=None):
    """Run one curve. Returns dict(x, z, f1, acc, f2, counters, save_line).
    M: base number of a special-form input (arithmetic mod M, checks and save line with N)."""
    L = lib()
    nh = ("%x" % N).encode()
    cap = len(nh if M is None else "%x" % M) + 16
    x, z, acc = (ctypes.create_string_buffer(cap) for _ in range(3))
    f1, f2 = (ctypes.create_string_buffer(2 * cap) for _ in range(2))
    cnt = (ctypes.c_uint32 * 8)()
    if M is None:
        rc = L.oracle_ecm_curve(nh, b1, b2, sigma, x, z, f1, acc, f2, cnt)
    else:
        rc = L.oracle_ecm_curve_special(nh, ("%x" % M).encode(), b1, b2, sigma, x, z, f1, acc, f2, cnt)
    assert rc == 0
    line = ctypes.create_string_buffer(4 * cap + 256)
    L.oracle_save_line(nh, b1, sigma, x.value, z.value, line, len(line))
    return {"x": int(x.value, 16), "z": int(z.value, 16), "f1": int(f1.value),
            "acc": int(acc.value, 16) if acc.value else None, "f2": int(f2.value),
            "counters": list(cnt), "save_line": line.value.decode()}


def build_curve(N, sigma):
    L = lib()
    nh = ("%x" % N).encode()
    x, s = (ctypes.create_string_buffer(len(nh) + 16) for _ in range(2))
    L.oracle_build_curve(nh, sigma, x, s)
    return int(x.value, 16), int(s.value, 16)


def pair(lo, hi, D, U=16):
    L = lib()
    cap = int((hi - lo) // 8 + 2 * (hi - lo) // D + 4096)
    v = (ctypes.c_uint32 * cap)()
    u = (ctypes.c_uint32 * cap)()
    amin, npairs = ctypes.c_uint32(), ctypes.c_uint32()
    steps = L.oracle_pair(lo, hi, D, U, v, u, cap, ctypes.byref(amin), ctypes.byref(npairs))
    assert steps <= cap
    return list(v[:steps]), list(u[:steps]), amin.value, npairs.value


def stage2_map(b1):
    L = lib()
    n = L.oracle_stage2_map(b1, None, 0)
    m = (ctypes.c_uint32 * n)()
    L.oracle_stage2_map(b1, m, n)
    return list(m)


def stage1_trace(b1):
    L = lib()
    n = L.oracle_stage1_trace(b1, None, 0)
    buf = ctypes.create_string_buffer(n + 1)
    L.oracle_stage1_trace(b1, buf, n)
    return buf.raw[:n]


def set_prime_range(r):
    """Width of the prime ranges stage 1 is run in (0 = the reference's 1e8)."""
    lib().oracle_set_prime_range(r)


def set_checkpoint(ranges):
    """Stop stage 1 after that many prime ranges (0 = run all): the state of checkpoint.txt."""
    lib().oracle_set_checkpoint(ranges)
