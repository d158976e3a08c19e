cd $GRAFT_REPO_ROOT
for sk in 0 1 2 3; do
  echo "=== 64->64 skip=$sk"; AMB_V4_SKIP=$sk AMB_V4_DBG=1 timeout 120 python tests/ncu_target.py 2>&1 | grep -E "V4DBG|rror" | sed -n '2p'
done
for sk in 0 1 2 3; do
  echo "=== 128->128@64 skip=$sk"; AMB_V4_SKIP=$sk AMB_V4_DBG=1 AMB_NT_CI=128 AMB_NT_CO=128 AMB_NT_S=64 timeout 120 python tests/ncu_target.py 2>&1 | grep -E "V4DBG|rror" | sed -n '2p'
done
