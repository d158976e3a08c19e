#!/bin/bash
# bring-up of the halo-plane wgrad kernel: parity in pair / single-tap mode, then A/B timing against the per-tap kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 120 python -m tests.kernel_checks conv "$1" 2>&1 | tail -2; }
for mode in PAIR SINGLE; do
  if [ $mode = SINGLE ]; then export AMB_WH_SINGLE=1; else unset AMB_WH_SINGLE; fi
  echo "== mode $mode"
  run '{"Cin":64,"Cout":64,"S":16,"impl":2}'
  run '{"Cin":64,"Cout":32,"S":16,"impl":2,"bias":false}'
  run '{"Cin":64,"Cout":64,"S":32,"impl":2}'
  run '{"Cin":128,"Cout":64,"S":32,"N":1,"impl":2}'
  run '{"Cin":64,"Cout":32,"S":24,"impl":2}'
  run '{"Cin":32,"Cout":32,"S":16,"impl":2}'
done
unset AMB_WH_SINGLE
echo "== halo"; timeout 300 python tests/conv_bench.py wgrad
echo "== per-tap"; AMB_DISABLE_WH=1 timeout 300 python tests/conv_bench.py wgrad
