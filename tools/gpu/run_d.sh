set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_stage2.py tests/test_gpu_stage1.py -m gpu -x -q -k "cooperative or syn2048 or reference_granularity or found_during or state_checks" > gpurun_out/r2d_tests_coop.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_tests_coop.log
tail -5 gpurun_out/r2d_tests_coop.log
for k in solo coop; do
  ECM_B200_S2_KERNEL=$k timeout 300 python tools/perf_probe2.py syn2048 14208 3000 300000 > gpurun_out/r2d_probe2_2048_$k.log 2>&1
  ECM_B200_S2_KERNEL=$k timeout 300 python tools/perf_probe2.py slow_csh_line07 14208 3000 300000 > gpurun_out/r2d_probe2_1165_$k.log 2>&1
done
tail -n1 gpurun_out/r2d_probe2*.log
timeout 300 python tools/perf_probe3.py syn415 65536 1000000 100000000 > gpurun_out/r2d_s2_415_default.log 2>&1
ECM_B200_PAIR_THREADS=384 timeout 300 python tools/perf_probe3.py syn415 65536 1000000 100000000 > gpurun_out/r2d_s2_415_pt384.log 2>&1
ECM_B200_PAIR_CHUNK=2048 timeout 300 python tools/perf_probe3.py syn415 65536 1000000 100000000 > gpurun_out/r2d_s2_415_chunk2048.log 2>&1
ECM_B200_PAIR_THREADS=384 ECM_B200_PAIR_CHUNK=2048 timeout 300 python tools/perf_probe3.py syn415 65536 1000000 100000000 > gpurun_out/r2d_s2_415_pt384_chunk2048.log 2>&1
tail -n1 gpurun_out/r2d_s2_415*.log
for k in vm rv; do
  ECM_B200_S1_KERNEL=$k timeout 300 python tools/perf_probe_special.py 415 1 1 65536 30000 > gpurun_out/r2d_fold_415_$k.log 2>&1
  ECM_B200_S1_KERNEL=$k timeout 300 python tools/perf_probe_special.py 277 1 1 65536 30000 > gpurun_out/r2d_fold_277_$k.log 2>&1
done
tail -n3 gpurun_out/r2d_fold*.log
for t in 384 512; do ECM_B200_THREADS=$t timeout 300 python tools/perf_probe.py syn415 75776 30000 > gpurun_out/r2d_probe_415_t$t.log 2>&1; done
tail -n1 gpurun_out/r2d_probe_415_t*.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 40 -c 1 -f -o gpurun_out/r2d_pair13 python tools/perf_probe3.py syn415 65536 100000 10000000 > gpurun_out/r2d_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_vm2 -s 6 -c 1 -f -o gpurun_out/r2d_vm2_13 python tools/perf_probe3.py syn415 65536 100000 10000000 > gpurun_out/r2d_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair_coop -s 4 -c 1 -f -o gpurun_out/r2d_pair_coop64 python tools/perf_probe2.py syn2048 14208 3000 300000 > gpurun_out/r2d_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_vm2_coop -s 1 -c 1 -f -o gpurun_out/r2d_vm2_coop64 python tools/perf_probe2.py syn2048 14208 1000 30000 > gpurun_out/r2d_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
