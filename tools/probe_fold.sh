#!/bin/bash
# fold vs Montgomery kernels on special-form bases (stage 1, and stage 1+2 for one size)
python tools/perf_probe_special.py 277 1 1 65536 30000 2>&1 | grep -E "^fold|^mont|speed"
python tools/perf_probe_special.py 415 1 1 65536 30000 2>&1 | grep -E "^fold|^mont|speed"
python tools/perf_probe_special.py 523 -1 1 65536 20000 2>&1 | grep -E "^fold|^mont|speed"
python tools/perf_probe_special.py 1000 -1 1 65536 5000 2>&1 | grep -E "^fold|^mont|speed"
python tools/perf_probe_special.py 220 1 69 65536 30000 2>&1 | grep -E "^fold|^mont|speed"
python tools/perf_probe_special.py 415 1 1 16384 20000 2000000 2>&1 | grep -E "^fold|^mont|speed"
