#!/usr/bin/env python3
"""tests/golden/classify.json: what the compiled reference (oracle/_ref/avx-ecm-ref) decides for Mersenne-like
input expressions -- the number it goes on to factor after removing algebraic factors, the special form it
selects, or REDC (main.c:405-521).  TEST INFRASTRUCTURE; needs the reference built (oracle/build_ref.sh)."""
import json, os, re, subprocess, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "avx-ecm-ref")
EXPRS = ["2^315-1", "2^330+1", "2^255+1", "2^405-1", "(2^429-1)/(2^143-1)", "2^512+1", "2^1155+1", "2^226-5", "2^521-1",
         "2^607-1", "2^300-1", "3*(2^200-119)", "2^127-1", "(2^128+1)/59649589127497217", "2^89-1", "2^1279-1", "2^2203-1",
         "(2^1061-1)/46817226351072265620777670675006972301618979214252832875068976303839400413682313921168154465151768472420980044715745858522803980473207943564433",
         "2^96-17", "2^64+13"]

out = []
for e in EXPRS:
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([REF, e, "8", "3", "1", "3", "1000003"], cwd=d, capture_output=True, text=True).stdout
    rec = {"expr": e, "error": "too many distinct odd factors" in r}
    if not rec["error"]:
        rec["n"] = re.search(r"commencing parallel ecm on (\d+)", r).group(1)
        m = re.search(r"Using special (?:pseudo-)?Mersenne mod for factor of: 2\^(\d+)([-+])(\d+)", r)
        rec["kind"] = 0 if not m else (int(m.group(3)) if m.group(2) == "-" else -1)
        rec["k"] = int(m.group(1)) if m else None
        rec["redc_forced"] = "determined to be faster by REDC" in r
    out.append(rec)
    print(rec)
json.dump(out, open(os.path.join(ROOT, "tests", "golden_classify.json"), "w"), indent=1)
