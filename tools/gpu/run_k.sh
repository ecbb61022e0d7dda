set -x
cd /root/repo
# A/B: dedicated dual squaring in the register machine (libecm_b200_rvsqr.so) against the default build
for lib in libecm_b200.so libecm_b200_rvsqr.so; do
  ECM_B200_LIB=$lib timeout 300 python tools/perf_probe.py syn415 65536 100000 > gpurun_out/r2k_s1_415_$lib.log 2>&1
  ECM_B200_LIB=$lib timeout 300 python tools/perf_probe.py readme508 65536 50000 > gpurun_out/r2k_s1_508_$lib.log 2>&1
  ECM_B200_LIB=$lib timeout 300 python tools/perf_probe.py t35 65536 100000 > gpurun_out/r2k_s1_297_$lib.log 2>&1
done
tail -n1 gpurun_out/r2k_s1_*.log
# correctness of the A/B build: goldens of 10/13/16 limbs
ECM_B200_LIB=libecm_b200_rvsqr.so timeout 600 python -m pytest tests/test_gpu_stage1.py -m gpu -x -q > gpurun_out/r2k_tests_rvsqr.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_tests_rvsqr.log
tail -3 gpurun_out/r2k_tests_rvsqr.log
# full-grid capture of the four-lanes-per-curve stage-1 kernel
bash tools/gpu/ncu_cap.sh r2k_stage1_coop64 k_stage1_rv 0 -- python tools/perf_probe.py syn2048 14208 1500
