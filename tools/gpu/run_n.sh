set -x
cd /root/repo
# in-between limb counts (28 one thread per curve; 40, 56 four lanes per curve): goldens, both machines, cooperative stage 2
timeout 1200 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_stage2.py -m gpu -x -q -k "syn880 or syn1250 or syn1750 or cooperative or syn1024 or syn2048" > gpurun_out/r2n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_tests.log
tail -3 gpurun_out/r2n_tests.log
timeout 300 python tools/perf_probe.py syn880 65536 10000 > gpurun_out/r2n_s1_880.log 2>&1
timeout 300 python tools/perf_probe.py syn1250 14208 10000 > gpurun_out/r2n_s1_1250.log 2>&1
timeout 300 python tools/perf_probe.py syn1750 14208 10000 > gpurun_out/r2n_s1_1750.log 2>&1
timeout 300 python tools/perf_probe.py csh_line19 65536 20000 > gpurun_out/r2n_s1_n20.log 2>&1
timeout 300 python tools/perf_probe.py csh_line02 65536 20000 > gpurun_out/r2n_s1_n24.log 2>&1
tail -n1 gpurun_out/r2n_s1_*.log
