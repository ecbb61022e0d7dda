set -x
cd /root/repo
# correctness of the single-body cooperative kernels (48/64 limbs, both stages)
timeout 900 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_stage2.py -m gpu -x -q -k "cooperative or golden or machines" > gpurun_out/r2l_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_tests.log
tail -3 gpurun_out/r2l_tests.log
timeout 300 python tools/perf_probe.py syn2048 14208 10000 > gpurun_out/r2l_s1_2048.log 2>&1
ECM_B200_S2_TRACE=1 timeout 300 python tools/perf_probe3.py syn2048 14208 11000000 20000000 > gpurun_out/r2l_s2_2048.log 2>&1
ECM_B200_S2_TRACE=1 ECM_B200_LIB=libecm_b200_hyb13.so timeout 300 python tools/perf_probe3.py syn415 65536 1000000 100000000 > gpurun_out/r2l_s2_415_hyb13.log 2>&1
tail -n2 gpurun_out/r2l_s*.log
bash tools/gpu/ncu_cap.sh r2l_stage1_n32 k_stage1 0 -- python tools/perf_probe.py syn1024 65536 1500
