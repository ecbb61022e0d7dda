# End-of-round evidence run: GPU suite, smoke, default bench line, ncu launch list of the bench command, full captures of
# the dominant kernels.  Everything lands in gpurun_out/r2_final_*; summaries are copied to profiles/ by hand.
set -x
cd /root/repo
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_final_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final_gputests.log
tail -5 gpurun_out/r2_final_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final_smoke.log
tail -2 gpurun_out/r2_final_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_final_bench_1gpu.err
# launch list of the bench command (short: 2 timed launches), then full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_final_stage1_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_final_launches_bench.log 2>&1
bash tools/gpu/ncu_cap.sh r2_final_stage1 k_stage1_rv 2 -- python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
bash tools/gpu/ncu_cap.sh r2_final_stage1_coop64 k_stage1_rv 0 -- python tools/perf_probe.py syn2048 14208 1500
bash tools/gpu/ncu_cap.sh r2_final_pair13 k_pair 40 -- python tools/perf_probe3.py syn415 65536 100000 10000000
bash tools/gpu/ncu_cap.sh r2_final_vm2_13 k_vm2 6 -- python tools/perf_probe3.py syn415 65536 100000 10000000
ls -la gpurun_out | tail -20
