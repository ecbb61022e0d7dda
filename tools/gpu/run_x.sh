set -x
cd /root/repo
ECM_B200_SLOW=1 timeout 1000 python -m pytest tests/test_gpu_known_answers.py -m gpu -q -o timeout=0 > gpurun_out/r2_final_known_answers_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final_known_answers_all.log
tail -4 gpurun_out/r2_final_known_answers_all.log
