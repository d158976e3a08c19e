cd $GRAFT_REPO_ROOT
timeout 300 python tests/convT_bench.py 2>&1 | grep RESULT | cut -c1-260
AMB_CB_LAYERS=6 timeout 300 python tests/conv_bench.py v1 2>&1 | tail -8
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or transpose or spark_step or golden" 2>&1 | tail -3
echo "=== bench default"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['clocks'])"
