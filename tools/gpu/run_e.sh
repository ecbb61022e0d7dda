set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_stage2.py -m gpu -x -q -k "cooperative" > gpurun_out/r2e_tests_coop.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_tests_coop.log
grep -E "^E|passed|failed" gpurun_out/r2e_tests_coop.log | head -20
timeout 300 python tools/perf_probe3.py syn415 65536 1000000 100000000 > gpurun_out/r2e_s2_415.log 2>&1
timeout 300 python tools/perf_probe3.py syn1024 32768 1000000 100000000 > gpurun_out/r2e_s2_1024.log 2>&1
timeout 300 python tools/perf_probe3.py syn2048 14208 100000 10000000 > gpurun_out/r2e_s2_2048.log 2>&1
tail -n1 gpurun_out/r2e_s2_*.log
timeout 300 python tools/perf_probe_special.py 415 1 1 65536 30000 3000000 > gpurun_out/r2e_fold_415.log 2>&1
timeout 300 python tools/perf_probe_special.py 277 1 1 65536 30000 > gpurun_out/r2e_fold_277.log 2>&1
timeout 300 python tools/perf_probe_special.py 523 -1 1 65536 20000 > gpurun_out/r2e_fold_523.log 2>&1
tail -n3 gpurun_out/r2e_fold*.log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2e_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_gputests.log
tail -25 gpurun_out/r2e_gputests.log
