set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_stage1.py -m gpu -x -q -k "register_machine or cooperative or golden" > gpurun_out/r2b_tests_s1.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_tests_s1.log
tail -5 gpurun_out/r2b_tests_s1.log
for k in vm rv; do
  ECM_B200_S1_KERNEL=$k timeout 300 python tools/perf_probe.py syn415 65536 30000 > gpurun_out/r2b_probe_415_$k.log 2>&1
  ECM_B200_S1_KERNEL=$k timeout 300 python tools/perf_probe.py syn2048 18944 3000 > gpurun_out/r2b_probe_2048_$k.log 2>&1
  ECM_B200_S1_KERNEL=$k timeout 300 python tools/perf_probe.py slow_csh_line07 28416 3000 > gpurun_out/r2b_probe_1165_$k.log 2>&1
  ECM_B200_S1_KERNEL=$k timeout 300 python tools/perf_probe.py readme508 65536 20000 > gpurun_out/r2b_probe_508_$k.log 2>&1
  ECM_B200_S1_KERNEL=$k timeout 300 python tools/perf_probe.py t35 65536 30000 > gpurun_out/r2b_probe_297_$k.log 2>&1
done
tail -n1 gpurun_out/r2b_probe_*.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_gputests.log
tail -5 gpurun_out/r2b_gputests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stage1_rv -s 2 -c 1 -f -o gpurun_out/r2b_rv13 python tools/perf_probe.py syn415 65536 30000 > gpurun_out/r2b_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stage1_rv -s 1 -c 1 -f -o gpurun_out/r2b_rv64 python tools/perf_probe.py syn2048 18944 1500 > gpurun_out/r2b_ncu2.log 2>&1
