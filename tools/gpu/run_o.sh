set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_stage2.py -m gpu -x -q -k "cooperative" > gpurun_out/r2o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_tests.log
tail -3 gpurun_out/r2o_tests.log
# 1024-bit stage 2: one thread per curve against two lanes per curve (same library, env switch)
ECM_B200_S2_TRACE=1 ECM_B200_S2_KERNEL=solo timeout 300 python tools/perf_probe3.py syn1024 32768 3000000 20000000 > gpurun_out/r2o_s2_1024_solo.log 2>&1
ECM_B200_S2_TRACE=1 timeout 300 python tools/perf_probe3.py syn1024 32768 3000000 20000000 > gpurun_out/r2o_s2_1024_coop2.log 2>&1
ECM_B200_S2_TRACE=1 ECM_B200_S2_KERNEL=solo timeout 300 python tools/perf_probe3.py syn880 32768 3000000 20000000 > gpurun_out/r2o_s2_880_solo.log 2>&1
ECM_B200_S2_TRACE=1 timeout 300 python tools/perf_probe3.py syn880 32768 3000000 20000000 > gpurun_out/r2o_s2_880_coop2.log 2>&1
tail -n2 gpurun_out/r2o_s2_*.log
timeout 300 python tools/perf_probe.py syn880 65536 100000 > gpurun_out/r2o_s1_880.log 2>&1
tail -n1 gpurun_out/r2o_s1_880.log
