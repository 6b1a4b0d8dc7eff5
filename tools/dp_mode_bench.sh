# usage: bash tools/dp_mode_bench.sh N mode...   (one bench line per data-parallel exchange mode on N GPUs)
N=$1; shift
for m in "$@"; do
  CB_DP_MODE=$m timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 100 --warmup 10 --no-posterior 2> gpurun_out/dp_${m}_$N.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', $N, 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 's/epoch', round(d['value'],5))" || tail -5 gpurun_out/dp_${m}_$N.err
done
grep -h "p2p exchange" gpurun_out/dp_*_$N.err 2>/dev/null
