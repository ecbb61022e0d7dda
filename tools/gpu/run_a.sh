set -x
cd /root/repo
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_gputests.log
timeout 300 python tools/perf_probe.py syn2048 18944 3000 > gpurun_out/r2a_probe_2048.log 2>&1
timeout 300 python tools/perf_probe.py syn1024 65536 5000 > gpurun_out/r2a_probe_1024.log 2>&1
timeout 300 python tools/perf_probe.py syn415 65536 30000 > gpurun_out/r2a_probe_415.log 2>&1
timeout 300 python tools/perf_probe2.py syn1024 32768 20000 2000000 > gpurun_out/r2a_probe2_1024.log 2>&1
timeout 300 python tools/perf_probe2.py syn2048 9472 3000 300000 > gpurun_out/r2a_probe2_2048.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stage1 -s 1 -c 1 -f -o gpurun_out/r2a_stage1_64 python tools/perf_probe.py syn2048 18944 1500 > gpurun_out/r2a_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stage1 -s 1 -c 1 -f -o gpurun_out/r2a_stage1_32 python tools/perf_probe.py syn1024 65536 1500 > gpurun_out/r2a_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 3 -c 1 -f -o gpurun_out/r2a_pair_32 python tools/perf_probe2.py syn1024 32768 3000 300000 > gpurun_out/r2a_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_vm2 -s 3 -c 1 -f -o gpurun_out/r2a_vm2_64 python tools/perf_probe2.py syn2048 9472 1000 30000 > gpurun_out/r2a_ncu4.log 2>&1
tail -3 gpurun_out/r2a_gputests.log; cat gpurun_out/r2a_probe*.log
