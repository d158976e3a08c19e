cd $GRAFT_REPO_ROOT
for co in 64 32; do
  echo "=== 64->$co"; AMB_V4_DBG=1 AMB_NT_CO=$co timeout 120 python tests/ncu_target.py 2>&1 | grep -E "V4DBG|rror" | sed -n '2p;5p'
done
echo "=== 128->128 @64"; AMB_V4_DBG=1 AMB_NT_CI=128 AMB_NT_CO=128 AMB_NT_S=64 timeout 120 python tests/ncu_target.py 2>&1 | grep -E "V4DBG|rror" | sed -n '2p'
echo "=== conv_bench"; AMB_CB_LAYERS=4 timeout 300 python tests/conv_bench.py v3 2>&1 | tail -6
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_fwd_dgrad or fused_inference or epilogue_stat or masked_conv or integers" 2>&1 | tail -3
