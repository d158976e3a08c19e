# last 2-GPU pass on the committed tree: distributed checks + STUNet-B line
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 300 $TR tests/dist_checks.py > gpurun_out/r2z_dist_checks_2gpu.log 2>&1; echo rc=$?; grep -E "RESULT all|Error|assert" gpurun_out/r2z_dist_checks_2gpu.log | tail -4
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2z_bench_B_2gpu.json 2> gpurun_out/r2z_bench_B_2gpu.err
python -c "import sys,json; d=json.loads(open('gpurun_out/r2z_bench_B_2gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])" || tail -5 gpurun_out/r2z_bench_B_2gpu.err
