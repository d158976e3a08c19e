cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -s -k "lean_zero or zero_shell or add_parity0 or conv_pair" 2>&1 | grep -E "RESULT|^E |passed|failed|Error|FAILED" | cut -c1-600 | head -60
