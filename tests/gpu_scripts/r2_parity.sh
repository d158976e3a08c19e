cd $GRAFT_REPO_ROOT
timeout 900 python -m tests.model_checks spark_report 2>&1 | grep RESULT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "sparse_batchnorm or script_step" 2>&1 | tail -30
