#!/bin/bash
# A/B of the plane-ring depth in the halo-plane conv kernel (igemm3): parity first, then isolated layer timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 120 python -m tests.kernel_checks "$1" "$2" 2>&1 | tail -1; }
run conv '{"Cin":64,"Cout":64,"S":32,"impl":2}'
run conv '{"Cin":64,"Cout":32,"S":24,"impl":2}'
run conv '{"Cin":128,"Cout":64,"S":32,"N":1,"impl":2}'
run conv_stats '{"Cin":64,"Cout":64,"S":32}'
export AMB_CB_LAYERS=4
for sl in 10 8 11; do echo "== plane slots $sl"; AMB_V3_A_SLOTS=$sl timeout 300 python tests/conv_bench.py v3; done
