# final build on an 8-GPU box: STUNet-B and BASELINE config 4 at N = 8
cd $GRAFT_REPO_ROOT
run() {
  N=$1; FLAGS=$2; TAG=$3; STEPS=$4
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N"
  echo "=== bench $TAG N=$N"
  timeout 300 $TR bench.py --gpus $N $FLAGS --steps $STEPS --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2z_bench_${TAG}_${N}gpu.json 2> gpurun_out/r2z_bench_${TAG}_${N}gpu.err; echo rc=$?
  python -c "import sys,json; d=json.loads(open('gpurun_out/r2z_bench_${TAG}_${N}gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])" || tail -5 gpurun_out/r2z_bench_${TAG}_${N}gpu.err
}
run 8 "" B 30
run 8 "--model L --sbn" Lsbn 10
