run() {
  tag="$1"; shift
  env "$@" timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-posterior 2> gpurun_out/nccl_$tag.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"
}
run base NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT
grep -i "nvls\|channels\|nchannels" gpurun_out/nccl_base.err | head -12
run minch32 NCCL_MIN_NCHANNELS=32
run maxch8 NCCL_MAX_NCHANNELS=8
run simple_ring NCCL_PROTO=Simple NCCL_ALGO=Ring
run nonvls NCCL_NVLS_ENABLE=0
run ctas NCCL_MIN_CTAS=32
ls -la gpurun_out/nccl_*.err | head; for f in gpurun_out/nccl_*.err; do head -c 20000 $f > $f.tmp; mv $f.tmp $f; done
