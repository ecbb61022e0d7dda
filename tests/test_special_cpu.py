"""Special-form (Mersenne-like) inputs, host side: the classification of main.c:405-521 as done by the
Python mirror (avx_ecm_b200.special_form) and by the command-line driver (--classify), against what the
compiled reference printed for the same expressions (tests/golden/special_*.json)."""
import os, re, subprocess
import pytest
from conftest import GOLDEN, ROOT, golden_base
import avx_ecm_b200 as E

SPECIAL = sorted(k for k, g in GOLDEN.items() if "expr" in g)
CLI = os.path.join(ROOT, "avx-ecm_b200", "avx-ecm-b200")


def evaluate(expr):
    return eval(expr.replace("^", "**").replace("/", "//"), {"__builtins__": {}})


def test_there_are_special_cases():
    kinds = {GOLDEN[k]["kind"] for k in SPECIAL}
    assert 1 in kinds and -1 in kinds and 0 in kinds and any(k > 1 for k in kinds)


@pytest.mark.parametrize("name", SPECIAL)
def test_python_classification_matches_reference(name):
    g = GOLDEN[name]
    f = E.special_form(evaluate(g["expr"]))
    assert f["n"] == int(g["n"])                       # "commencing parallel ecm on ..." after factor removal
    assert f["kind"] == g["kind"] and f["base"] == golden_base(g)
    if g["kind"]:
        assert f["k"] == g["k"]


@pytest.mark.parametrize("name", SPECIAL)
def test_cli_classification_matches_reference(name):
    g = GOLDEN[name]
    if not os.path.exists(CLI):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "avx-ecm_b200"), "cli"], check=True)
    out = subprocess.run([CLI, "--classify", g["expr"]], capture_output=True, text=True, check=True).stdout
    m = re.search(r"kind (-?\d+) k (\d+) n (\d+) base (\d+)", out)
    assert int(m.group(1)) == g["kind"] and int(m.group(3)) == int(g["n"])
    assert int(m.group(4)) == (golden_base(g) or 0)
    assert ("determined to be faster by REDC" in out) == g["redc_forced"]
    if g["kind"]:
        assert int(m.group(2)) == g["k"]
        assert ("Using special pseudo-Mersenne mod" in out) == (g["kind"] > 1)


def test_primitive_part():
    # 2^15-1 = 7 * 31 * 151: the primitive part w.r.t. the odd primes 3 and 5 is 151
    assert E.primitive_part(15, -1) == 151
    assert E.primitive_part(277, -1) == 2 ** 277 - 1
    assert E.primitive_part(523, 1) == (2 ** 523 + 1) // 3
    assert E.primitive_part(1024, 1) == 2 ** 1024 + 1
    with pytest.raises(ValueError):
        E.primitive_part(3 * 5 * 7 * 11, -1)


def test_generic_inputs_stay_generic():
    from conftest import composites
    for N in composites().values():
        f = E.special_form(N)
        assert f["kind"] == 0 and f["base"] is None and f["n"] == N


def test_create_special_rejects_bad_arguments():
    import ctypes
    L = E.lib()
    h = ctypes.c_void_p()
    n = (ctypes.c_uint32 * 2)(7, 1)
    big = (ctypes.c_uint32 * 1)(15)
    # the input number must not be wider than its base, and both must be odd
    assert L.ecm_b200_create_special(ctypes.byref(h), 0, big, 1, n, 2, 8) == -1
    even = (ctypes.c_uint32 * 2)(8, 1)
    assert L.ecm_b200_create_special(ctypes.byref(h), 0, n, 2, even, 2, 8) == -1
    assert L.ecm_b200_create_special(ctypes.byref(h), 0, n, 2, None, 0, 8) == -1


# ---- classification only: more shapes, as decided by the compiled reference (tools/gen_classify_golden.py) ------
import json
CLASSIFY = json.load(open(os.path.join(ROOT, "tests", "golden_classify.json")))


@pytest.mark.parametrize("rec", CLASSIFY, ids=[r["expr"][:40] for r in CLASSIFY])
def test_classification_of_more_shapes_matches_reference(rec):
    N = evaluate(rec["expr"])
    if rec["error"]:                                   # the reference gives up: more than three distinct odd primes in k
        with pytest.raises(ValueError):
            E.special_form(N)
        r = subprocess.run([CLI, "--classify", rec["expr"]], capture_output=True, text=True)
        assert r.returncode == 1 and "too many distinct odd factors" in r.stdout
        return
    f = E.special_form(N)
    assert f["n"] == int(rec["n"]) and f["kind"] == rec["kind"]
    out = subprocess.run([CLI, "--classify", rec["expr"]], capture_output=True, text=True, check=True).stdout
    m = re.search(r"kind (-?\d+) k (\d+) n (\d+) base (\d+)", out)
    assert int(m.group(1)) == rec["kind"] and int(m.group(3)) == int(rec["n"])
    assert ("determined to be faster by REDC" in out) == rec["redc_forced"]
    if rec["kind"]:
        assert f["k"] == rec["k"] == int(m.group(2))
        assert int(m.group(4)) == f["base"]
