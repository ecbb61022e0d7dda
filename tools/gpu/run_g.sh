set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_stage2.py -m gpu -x -q -k "cooperative" > gpurun_out/r2g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_tests.log
grep -E "^E|passed|failed|skipped" gpurun_out/r2g_tests.log | head -20
for lib in libecm_b200.so libecm_b200_spdual.so; do for t in 384 512; do
  ECM_B200_LIB=$lib ECM_B200_THREADS=$t timeout 300 python tools/perf_probe_special.py 415 1 1 65536 30000 > gpurun_out/r2g_fold_415_${lib}_t$t.log 2>&1
done; done
tail -n3 gpurun_out/r2g_fold_415_*.log
tools/gpu/ncu_cap.sh r2g_pair13 k_pair 40 -- python tools/perf_probe3.py syn415 65536 100000 10000000
tools/gpu/ncu_cap.sh r2g_vm2_13 k_vm2 6 -- python tools/perf_probe3.py syn415 65536 100000 10000000
tools/gpu/ncu_cap.sh r2g_pair_coop64 k_pair_coop 4 -- python tools/perf_probe2.py syn2048 14208 3000 300000
tools/gpu/ncu_cap.sh r2g_vm2_coop64 k_vm2_coop 1 -- python tools/perf_probe2.py syn2048 14208 1000 30000
tools/gpu/ncu_cap.sh r2g_pair32 k_pair 20 -- python tools/perf_probe3.py syn1024 32768 100000 10000000
timeout 1700 python bench.py --config syn1024_s12 > gpurun_out/r2g_config2_full.json 2> gpurun_out/r2g_config2_full.err; echo "config2 rc=$?"
tail -c 400 gpurun_out/r2g_config2_full.err; cut -c1-1500 gpurun_out/r2g_config2_full.json
