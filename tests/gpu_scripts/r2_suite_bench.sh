# whole GPU suite + default bench + STUNet-L/SyncBN 1-GPU point (denominator of the config-4 scaling line)
cd $GRAFT_REPO_ROOT
TAG=${1:-r2h}
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-400 | tail -20
timeout 600 python tests/step_profile.py ${TAG}_step 2>&1 | tail -2
echo "=== bench default"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null > gpurun_out/${TAG}_bench_B_1gpu.json; python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_bench_B_1gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['clocks'])"
echo "=== bench L sbn"; timeout 600 python bench.py --model L --sbn --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null > gpurun_out/${TAG}_bench_Lsbn_1gpu.json; python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_bench_Lsbn_1gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['clocks'])"
