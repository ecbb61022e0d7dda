#!/bin/bash
# ncu_cap.sh <name> <kernel-regex> <skip> -- <command...> : one `ncu --set full` capture, summarised ON THE BOX (the .ncu-rep
# files are tens of MB each and gpurun brings back at most 64 MiB); keeps gpurun_out/<name>_ncu_summary.txt only.
name=$1; regex=$2; skip=$3; shift 4
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
python tools/ncu_summary.py /tmp/$name.ncu-rep "$name: ncu -k regex:$regex -s $skip -c 1 ; $*" > gpurun_out/${name}_ncu_summary.txt 2>> gpurun_out/${name}_ncu.log
tail -2 gpurun_out/${name}_ncu.log
