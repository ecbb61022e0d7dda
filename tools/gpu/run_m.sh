set -x
cd /root/repo
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_tests.log
tail -5 gpurun_out/r2m_tests.log
ECM_B200_S2_TRACE=1 timeout 300 python tools/perf_probe3.py syn415 65536 1000000 100000000 > gpurun_out/r2m_s2_415.log 2>&1
tail -n2 gpurun_out/r2m_s2_415.log
bash tools/probe_fold.sh > gpurun_out/r2m_fold.log 2>&1
tail -n 20 gpurun_out/r2m_fold.log
