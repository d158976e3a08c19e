# 2-GPU A/B of the gradient exchange: bucket split (default), NCCL CTA limits, exchange outside the graph
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
echo "=== staged-input test (1 GPU)"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "staged or transpose" 2>&1 | tail -3
echo "=== dist_checks"; timeout 400 $TR tests/dist_checks.py > gpurun_out/r2i_dist_checks_2gpu.log 2>&1; echo rc=$?; grep -E "RESULT|Error|error|assert" gpurun_out/r2i_dist_checks_2gpu.log | tail -12
line() { python -c "import sys,json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['clocks'])" || tail -5 $1.err; }
for sw in "X=0" "NCCL_MAX_CTAS=8" "NCCL_MAX_CTAS=4" "NCCL_MAX_CTAS=2" "AMB_NCCL_OUTSIDE_GRAPH=1" "AMB_BENCH_NO_PREFETCH=1"; do
  echo "=== bench B N=2 $sw"
  env $sw timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2i_ab_$sw.json 2> gpurun_out/r2i_ab_$sw.json.err
  line gpurun_out/r2i_ab_$sw.json
done
