set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_stage2.py tests/test_gpu_stage1.py -m gpu -x -q -k "cooperative or syn2048 or reference_granularity or found_during or state_checks" > gpurun_out/r2c_tests_coop.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_tests_coop.log
tail -5 gpurun_out/r2c_tests_coop.log
for k in solo coop; do
  ECM_B200_S2_KERNEL=$k timeout 300 python tools/perf_probe2.py syn2048 14208 3000 300000 > gpurun_out/r2c_probe2_2048_$k.log 2>&1
  ECM_B200_S2_KERNEL=$k timeout 300 python tools/perf_probe2.py slow_csh_line07 14208 3000 300000 > gpurun_out/r2c_probe2_1165_$k.log 2>&1
done
timeout 300 python tools/perf_probe.py syn2048 28416 3000 > gpurun_out/r2c_probe_2048_rv.log 2>&1
tail -n1 gpurun_out/r2c_probe*.log
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2c_bench.err
timeout 900 python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/r2c_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_gputests.log
tail -40 gpurun_out/r2c_gputests.log
