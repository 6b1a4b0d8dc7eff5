#!/bin/bash
# One gpurun call that regenerates the round's profiler evidence (profiles/README.md says which file came from what):
#   per-stream CUPTI timelines of a replayed step and of a posterior pass, `ncu --set full` captures of the step's and
#   the posterior's kernels (cudaProfilerStart/Stop ranges inside tools/step_profile.py), and the ncu launch list of a
#   short bench.py run.  Each ncu command runs only after the same command has exited 0 without ncu.
#     gpurun --timeout 900 -- 'bash tools/profile_round.sh s4'
tag=${1:-prof}
out=gpurun_out
mkdir -p $out
timeout 120 python tools/step_timeline.py --out $out/${tag}_step_timeline.txt > $out/${tag}_step_timeline.log 2>&1
echo "step_timeline rc $?"
timeout 120 python tools/posterior_profile.py --cells 8000 --reps 2 --timeline $out/${tag}_posterior_timeline.txt > $out/${tag}_posterior_timeline.log 2>&1
echo "posterior timeline rc $?"
K='obs_|gemm_tf32|gather_rows|bn0_forward'
timeout 120 python tools/step_profile.py --steps 5 > $out/${tag}_plain_step.log 2>&1 &&
timeout 420 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$K" -c 14 -f \
    -o $out/${tag}_step python tools/step_profile.py --steps 5 > $out/${tag}_ncu_step.log 2>&1
echo "ncu step rc $?"
ncu -i $out/${tag}_step.ncu-rep --page raw --csv 2>/dev/null | gzip > $out/${tag}_step_raw.csv.gz
rm -f $out/${tag}_step.ncu-rep  # gpurun_out/ travels back only below 64 MiB
timeout 120 python tools/step_profile.py --steps 2 --posterior > $out/${tag}_plain_post.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    --kernel-name-base demangled -k regex:'posterior_window|col_lse|gemm_tf32_kernel<.{0,8}2,.{0,8}256' -c 10 -f -o $out/${tag}_posterior python tools/step_profile.py --steps 2 --posterior \
    > $out/${tag}_ncu_post.log 2>&1
echo "ncu posterior rc $?"
ncu -i $out/${tag}_posterior.ncu-rep --page raw --csv 2>/dev/null | gzip > $out/${tag}_posterior_raw.csv.gz
rm -f $out/${tag}_posterior.ncu-rep
B="python bench.py --steps 2 --warmup 3 --no-posterior --no-cpu-baseline --no-kernels --short-warmup"
timeout 120 $B > $out/${tag}_plain_bench.log 2>&1 &&
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/${tag}_launches.csv $B \
    > $out/${tag}_ncu_bench.log 2>&1
echo "ncu launches rc $?"
ls -la $out | tail -20
