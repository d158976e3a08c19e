#!/bin/bash
# bring-up of the halo-plane wgrad kernel: parity, then A/B timing of its variants against the per-tap kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 120 python -m tests.kernel_checks "$1" "$2" 2>&1 | tail -1; }
run conv '{"Cin":64,"Cout":64,"S":16,"impl":2}'
run conv '{"Cin":64,"Cout":32,"S":16,"impl":2,"bias":false}'
run conv '{"Cin":128,"Cout":128,"S":16,"impl":2}'
run convT '{"Cin":64,"Cout":64,"S":16,"impl":2}'
run convT '{"Cin":128,"Cout":128,"S":16,"N":1,"impl":2}'
run convT '{"Cin":256,"Cout":256,"S":16,"N":1,"impl":2}'
run convT '{"Cin":128,"Cout":64,"S":16,"N":1,"impl":2}'
export AMB_CB_LAYERS=0
echo "== halo"; timeout 300 python tests/conv_bench.py wgrad
echo "== per-tap convT"; AMB_WH_NO_CONVT=1 timeout 300 python tests/conv_bench.py wgrad
