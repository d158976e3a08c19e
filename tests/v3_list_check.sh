#!/bin/bash
cd "$(dirname "$0")/.."
run() { timeout 120 python -m tests.kernel_checks "$1" "$2" 2>&1 | tail -1; }
run conv '{"Cin":32,"Cout":32,"S":32,"impl":2,"masked":true}'
run conv '{"Cin":64,"Cout":64,"S":32,"impl":2,"masked":true,"f":2}'
run conv '{"Cin":64,"Cout":32,"S":64,"N":1,"impl":2,"masked":true,"f":4}'
run conv '{"Cin":32,"Cout":32,"S":64,"N":1,"impl":2,"masked":true,"f":2}'
run conv '{"Cin":64,"Cout":64,"S":32,"impl":2,"masked":true,"f":8}'
run conv '{"Cin":64,"Cout":64,"S":32,"impl":2}'
b() { timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; }
echo "== v3 list"; b
echo "== per-tap list"; AMB_V3_NO_LIST=1 b
