set -x
cd /root/repo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/r2t_s2_415_launches_raw.csv python tools/perf_probe3.py syn415 65536 1000000 100000000 > gpurun_out/r2t_s2_415_launches.log 2>&1
tail -2 gpurun_out/r2t_s2_415_launches.log
