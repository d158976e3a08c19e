# whole GPU suite (what the driver runs at round end)
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/ -q -m gpu -s 2>&1 | grep -E "RESULT (augment|checkpoint|gradient)|passed|failed|Error|FAILED|^E  " | tail -40
