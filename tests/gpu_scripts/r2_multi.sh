# 2-GPU pass: distributed checks, STUNet-B scaling point, BASELINE config 4 (STUNet-L + decoder SyncBN) at N GPUs
cd $GRAFT_REPO_ROOT
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "=== dist_checks"; date; timeout 400 $TR tests/dist_checks.py > gpurun_out/r2_dist_checks_${N}gpu.log 2>&1; echo rc=$?; grep -E "RESULT|Error|error|assert" gpurun_out/r2_dist_checks_${N}gpu.log | tail -20
echo "=== bench B N=$N"; date; timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_B_${N}gpu.json 2> gpurun_out/bench_r2_B_${N}gpu.err; echo rc=$?; tail -c 300 gpurun_out/bench_r2_B_${N}gpu.json; tail -3 gpurun_out/bench_r2_B_${N}gpu.err
echo "=== bench L sbn N=$N"; date; timeout 400 $TR bench.py --gpus $N --model L --sbn --steps 6 --warmup 3 > gpurun_out/bench_r2_Lsbn_${N}gpu.json 2> gpurun_out/bench_r2_Lsbn_${N}gpu.err; echo rc=$?; tail -c 300 gpurun_out/bench_r2_Lsbn_${N}gpu.json; tail -3 gpurun_out/bench_r2_Lsbn_${N}gpu.err
echo "(reference arm under torchrun: verified earlier in the round)"
date
