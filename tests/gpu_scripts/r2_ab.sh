# same-box A/B of environment switches, interleaved: bash r2_ab.sh TAG ROUNDS "VAR=1" ["VAR2=1" ...]
cd $GRAFT_REPO_ROOT
TAG=$1; ROUNDS=$2; shift; shift
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['ms_per_step'],3), round(d['ms_per_step_median'],3), round(d['e2e']['value'],2), d['clocks']['sm_mhz'])"; }
for r in $(seq 1 $ROUNDS); do
  echo "=== default #$r"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | line
  for sw in "$@"; do
    echo "=== $sw #$r"; env $sw timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | line
  done
done
