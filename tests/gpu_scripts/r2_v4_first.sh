set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv_fwd_dgrad_wgrad or masked_conv_over or epilogue_statistics or integers" 2>&1 | tail -15
echo "=== v4 default"; AMB_CB_LAYERS=4 timeout 300 python tests/conv_bench.py v3 2>&1 | tail -8
echo "=== v4 T=6"; AMB_V4_T=6 AMB_CB_LAYERS=2 timeout 300 python tests/conv_bench.py v3 2>&1 | grep -v convT | tail -3
echo "=== v4 order 1"; AMB_V4_ORDER=1 AMB_CB_LAYERS=2 timeout 300 python tests/conv_bench.py v3 2>&1 | grep -v convT | tail -3
echo "=== v3 (V4 disabled)"; AMB_DISABLE_V4=1 AMB_CB_LAYERS=4 timeout 300 python tests/conv_bench.py v3 2>&1 | tail -8
