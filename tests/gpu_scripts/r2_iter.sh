# iteration pass: new tests, whole GPU suite, step profile, bench A/B over env switches given as arguments ("VAR=1" ...)
cd $GRAFT_REPO_ROOT
TAG=${1:-iter}; shift
timeout 1500 python -m pytest tests/ -q -m gpu -s 2>&1 | grep -E "RESULT lean|^E  |passed|failed|FAILED" | cut -c1-500 | tail -30
timeout 600 python tests/step_profile.py ${TAG} 2>&1 | tail -2
echo "=== bench default"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['clocks'])"
for sw in "$@"; do
  echo "=== bench $sw"; env $sw timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['ms_per_step_median'], d['e2e']['value'], d['clocks'])"
done
