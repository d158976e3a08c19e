# final build on N GPUs: distributed checks, STUNet-B, BASELINE config 4 (STUNet-L + decoder SyncBN)
cd $GRAFT_REPO_ROOT
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
line() { python -c "import sys,json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['clocks'])" || tail -5 $1.err; }
if [ "$N" = "2" ]; then
  echo "=== dist_checks"; timeout 400 $TR tests/dist_checks.py > gpurun_out/r2g_dist_checks_${N}gpu.log 2>&1; echo rc=$?; grep -E "RESULT|Error|error|assert" gpurun_out/r2g_dist_checks_${N}gpu.log | tail -12
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "two_gpu" 2>&1 | tail -2
fi
echo "=== bench B N=$N"; timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2g_bench_B_${N}gpu.json 2> gpurun_out/r2g_bench_B_${N}gpu.json.err; line gpurun_out/r2g_bench_B_${N}gpu.json
echo "=== bench L sbn N=$N"; timeout 400 $TR bench.py --gpus $N --model L --sbn --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2g_bench_Lsbn_${N}gpu.json 2> gpurun_out/r2g_bench_Lsbn_${N}gpu.json.err; line gpurun_out/r2g_bench_Lsbn_${N}gpu.json
echo "=== reference arm under torchrun"; timeout 300 $TR bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-400
