#!/bin/bash
# bring-up of the halo-plane wgrad kernel: parity, then A/B timing (one vs two issuing warps, per-tap kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 120 python -m tests.kernel_checks conv "$1" 2>&1 | tail -1; }
run '{"Cin":64,"Cout":64,"S":16,"impl":2}'
run '{"Cin":64,"Cout":32,"S":16,"impl":2,"bias":false}'
run '{"Cin":128,"Cout":64,"S":32,"N":1,"impl":2}'
run '{"Cin":64,"Cout":32,"S":24,"impl":2}'
echo "== halo, 2 issuers"; timeout 300 python tests/conv_bench.py wgrad
echo "== halo, 1 issuer"; AMB_WH_ISSUERS=1 timeout 300 python tests/conv_bench.py wgrad
if [ -n "$WH_PER_TAP" ]; then echo "== per-tap"; AMB_DISABLE_WH=1 timeout 300 python tests/conv_bench.py wgrad; fi
