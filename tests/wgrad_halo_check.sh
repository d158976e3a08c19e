#!/bin/bash
# bring-up of the halo-plane wgrad kernel: parity, then A/B timing of its variants against the per-tap kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 120 python -m tests.kernel_checks conv "$1" 2>&1 | tail -1; }
run '{"Cin":64,"Cout":64,"S":16,"impl":2}'
run '{"Cin":64,"Cout":32,"S":16,"impl":2,"bias":false}'
run '{"Cin":128,"Cout":128,"S":16,"impl":2}'
run '{"Cin":256,"Cout":256,"S":16,"impl":2}'
run '{"Cin":64,"Cout":128,"S":32,"impl":2,"bias":false}'
run '{"Cin":512,"Cout":512,"S":16,"N":1,"impl":2}'
export AMB_CB_LAYERS=6
echo "== halo (wide, 64-column accumulators)"; timeout 300 python tests/conv_bench.py wgrad
echo "== halo (wide, 128-column accumulators)"; AMB_WH_WIDE_NT=128 timeout 300 python tests/conv_bench.py wgrad
echo "== narrow only"; AMB_WH_NO_WIDE=1 timeout 300 python tests/conv_bench.py wgrad
