set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_stage2.py -m gpu -x -q > gpurun_out/r2h_tests_s2.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_tests_s2.log
grep -E "^E|passed|failed|skipped" gpurun_out/r2h_tests_s2.log | head
for lib in libecm_b200.so libecm_b200_hyb13.so; do
  ECM_B200_LIB=$lib timeout 300 python tools/perf_probe3.py syn415 65536 1000000 100000000 > gpurun_out/r2h_s2_415_$lib.log 2>&1
  ECM_B200_LIB=$lib timeout 300 python tools/perf_probe3.py syn415 65536 100000 10000000 > gpurun_out/r2h_s2_415small_$lib.log 2>&1
done
timeout 300 python tools/perf_probe3.py syn2048 14208 100000 10000000 > gpurun_out/r2h_s2_2048.log 2>&1
timeout 300 python tools/perf_probe3.py syn1024 32768 100000 10000000 > gpurun_out/r2h_s2_1024.log 2>&1
tail -n1 gpurun_out/r2h_s2_*.log
for lib in libecm_b200.so libecm_b200_rv24.so; do
  ECM_B200_LIB=$lib timeout 300 python tools/perf_probe.py csh_line19 65536 20000 > gpurun_out/r2h_n20_$lib.log 2>&1
  ECM_B200_LIB=$lib timeout 300 python tools/perf_probe.py csh_line02 65536 20000 > gpurun_out/r2h_n24_$lib.log 2>&1
done
tail -n1 gpurun_out/r2h_n2*.log
ECM_B200_SLOW=1 timeout 1200 python -m pytest tests/test_gpu_known_answers.py -m gpu -q -o timeout=0 > gpurun_out/r2h_known_answers_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_known_answers_all.log
tail -4 gpurun_out/r2h_known_answers_all.log
