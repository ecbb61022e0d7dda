set -x
cd /root/repo
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2i_tests.log 2>&1; echo rc=$? >> gpurun_out/r2i_tests.log
tail -5 gpurun_out/r2i_tests.log
( time timeout 900 python bench.py ) > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2i_bench.err
