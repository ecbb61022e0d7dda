set -x
cd /root/repo
for t in 256 384; do
  ECM_B200_THREADS=$t timeout 300 python tools/perf_probe.py syn880 65536 50000 2>&1 | tail -1 > gpurun_out/r2u_s1_880_t$t.log
done
ECM_B200_S1_KERNEL=vm timeout 300 python tools/perf_probe.py csh_line02 65536 20000 2>&1 | tail -1 > gpurun_out/r2u_s1_n24_vm.log
timeout 300 python tools/perf_probe.py csh_line02 65536 20000 2>&1 | tail -1 > gpurun_out/r2u_s1_n24_rv.log
cat gpurun_out/r2u_*.log
