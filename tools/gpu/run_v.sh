set -x
cd /root/repo
for t in 128 256; do
  ECM_B200_S1_KERNEL=vm ECM_B200_THREADS=$t timeout 300 python tools/perf_probe.py csh_line02 65536 50000 2>&1 | tail -1 > gpurun_out/r2v_s1_n24_vm_t$t.log
  ECM_B200_S1_KERNEL=vm ECM_B200_THREADS=$t timeout 300 python tools/perf_probe.py csh_line19 65536 50000 2>&1 | tail -1 > gpurun_out/r2v_s1_n20_vm_t$t.log
done
for t in 256 384; do
  ECM_B200_THREADS=$t timeout 300 python tools/perf_probe.py csh_line02 65536 50000 2>&1 | tail -1 > gpurun_out/r2v_s1_n24_rv_t$t.log
  ECM_B200_THREADS=$t timeout 300 python tools/perf_probe.py csh_line19 65536 50000 2>&1 | tail -1 > gpurun_out/r2v_s1_n20_rv_t$t.log
done
ECM_B200_THREADS=128 timeout 300 python tools/perf_probe.py syn880 65536 50000 2>&1 | tail -1 > gpurun_out/r2v_s1_880_t128.log
ECM_B200_THREADS=128 timeout 300 python tools/perf_probe.py syn1024 65536 30000 2>&1 | tail -1 > gpurun_out/r2v_s1_1024_t128.log
ECM_B200_THREADS=256 timeout 300 python tools/perf_probe.py syn1024 65536 30000 2>&1 | tail -1 > gpurun_out/r2v_s1_1024_t256.log
cat gpurun_out/r2v_*.log
