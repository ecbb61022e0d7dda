set -x
cd /root/repo
( time timeout 150 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r2_final_bench_reference_arm.json 2> gpurun_out/r2_final_bench_reference_arm.err
tail -3 gpurun_out/r2_final_bench_reference_arm.err; cut -c1-600 gpurun_out/r2_final_bench_reference_arm.json
