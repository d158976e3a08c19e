# whole GPU suite + per-kernel table of one step (torch.profiler)
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -5
timeout 600 python tests/step_profile.py ${1:-r2_step} 2>&1 | tail -5
