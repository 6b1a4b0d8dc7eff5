#!/bin/bash
# ncu --set full capture of two posterior_window_kernel launches of a pipelined pass (the pipeline's first two window
# launches work on empty slots and are skipped), after the same command has exited 0 without ncu.
tag=${1:-prof}
out=gpurun_out
mkdir -p $out
timeout 120 python tools/step_profile.py --steps 2 --posterior > $out/${tag}_plain_window.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:posterior_window -s 2 -c 2 -f \
    -o $out/${tag}_window python tools/step_profile.py --steps 2 --posterior > $out/${tag}_ncu_window.log 2>&1
echo "ncu window rc $?"
ncu -i $out/${tag}_window.ncu-rep --page raw --csv 2>/dev/null | gzip > $out/${tag}_window_raw.csv.gz
ls -la $out
