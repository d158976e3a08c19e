# round 2, 1 GPU: the remaining sparse-layer API + checkpointed variants + device-side input pipeline
cd $GRAFT_REPO_ROOT
echo "=== f4/f2 tests"; timeout 1200 python -m pytest tests/test_kernels_gpu.py -q -k "remaining_sparse or mednext or gradient_checkpointed or input_pipeline" -s 2>&1 | grep -E "RESULT|passed|failed|Error|error|assert|FAILED" | tail -40
