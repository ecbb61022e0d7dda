set -x
cd /root/repo
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_stage2.py -m gpu -x -q -k "two_gpus" > gpurun_out/r2w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_tests.log
grep -E "^E|passed|failed|skipped" gpurun_out/r2w_tests.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2w_bench_n2.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2w_bench_n2.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['frac'])
a=l['also']
print('sweep', a.get('sweep')); print('cli', a.get('cli_multi_gpu'))
print('c2', a['config2_1024bit'].get('error') or a['config2_1024bit']['stage2']['range_head']['frac_of_imad_peak'])
print('c4', a['config4_2048bit'].get('error') or (a['config4_2048bit']['stage1']['frac_of_imad_peak'], a['config4_2048bit']['stage2']['init']['frac_of_imad_peak'], a['config4_2048bit']['stage2']['range_head']['frac_of_imad_peak']))
PY
