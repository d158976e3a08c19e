cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv_fwd_dgrad_wgrad or integers" 2>&1 | tail -6
echo "=== wns default"; AMB_CB_LAYERS=2 timeout 300 python tests/conv_bench.py wgrad 2>&1 | grep -v convT | tail -3
echo "=== wns PW=16"; AMB_WNS_PW=16 AMB_CB_LAYERS=2 timeout 300 python tests/conv_bench.py wgrad 2>&1 | grep -v convT | tail -3
echo "=== wns disabled"; AMB_DISABLE_WNS=1 AMB_CB_LAYERS=2 timeout 300 python tests/conv_bench.py wgrad 2>&1 | grep -v convT | tail -3
bash tests/gpu_scripts/r2_bench_dump.sh r2d
