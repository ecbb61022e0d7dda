set -x
cd /root/repo
( time timeout 900 python -m pytest tests/test_gpu_stage2.py tests/test_gpu_known_answers.py -m gpu -x -q ) > gpurun_out/r2q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_tests.log
tail -4 gpurun_out/r2q_tests.log
ECM_B200_S2_TRACE=1 timeout 300 python tools/perf_probe3.py syn415 65536 1000000 100000000 > gpurun_out/r2q_s2_415.log 2>&1
ECM_B200_S2_TRACE=1 timeout 300 python tools/perf_probe3.py syn1024 32768 3000000 20000000 > gpurun_out/r2q_s2_1024.log 2>&1
tail -n2 gpurun_out/r2q_s2_*.log
