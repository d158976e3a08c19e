# ConvTranspose forward on the kz-stacked halo-plane kernel: exactness vs the per-tap kernel, timings, convT tests, bench
cd $GRAFT_REPO_ROOT
timeout 300 python tests/convT_bench.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "transpose or spark_step or golden" 2>&1 | tail -5
echo "=== bench default"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['clocks'])"
echo "=== bench AMB_DISABLE_V4T=1"; AMB_DISABLE_V4T=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['clocks'])"
