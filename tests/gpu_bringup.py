"""Runs every kernel check in its own process (a trapped kernel must not poison the others).
   python tests/gpu_bringup.py [filter]  → gpurun_out/bringup.log"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D, T = 1, 2
CASES = [
    ('layout', {}), ('loss', {}), ('hard_mask', {}), ('ema_adamw', {}), ('proj', {}), ('stem', {}),
    ('norm', dict(mode='sparse', act=1, residual=True)), ('norm', dict(mode='sparse', act=0, residual=False, Cc=32, S=32)),
    ('norm', dict(mode='dense', act=2, residual=False)), ('norm', dict(mode='dense', act=0, residual=False, Cc=512, S=8)),
    ('norm', dict(mode='fill', act=0, residual=False)),
    ('conv', dict(Cin=16, Cout=16, S=8, impl=D)), ('conv', dict(Cin=16, Cout=24, S=8, stride=2, impl=D)),
    ('conv', dict(Cin=8, Cout=16, S=8, k=1, stride=2, impl=D)), ('conv', dict(Cin=16, Cout=8, S=8, k=1, impl=D)),
    ('conv', dict(Cin=16, Cout=16, S=16, impl=D, masked=True)), ('conv', dict(Cin=16, Cout=16, S=16, stride=2, impl=D, masked=True)),
    ('convT', dict(Cin=16, Cout=8, S=4, impl=D)),
    ('conv', dict(Cin=64, Cout=64, S=16, impl=T)), ('conv', dict(Cin=32, Cout=32, S=16, impl=T)),
    ('conv', dict(Cin=16, Cout=16, S=16, impl=T)), ('conv', dict(Cin=128, Cout=256, S=8, impl=T)),
    ('conv', dict(Cin=512, Cout=512, S=8, impl=T)), ('conv', dict(Cin=64, Cout=32, S=16, impl=T, bias=False)),
    ('conv', dict(Cin=64, Cout=64, S=12, impl=T)), ('conv', dict(Cin=256, Cout=128, S=4, impl=T)),
    ('conv', dict(Cin=64, Cout=128, S=16, stride=2, impl=T)), ('conv', dict(Cin=32, Cout=64, S=16, k=1, stride=2, impl=T)),
    ('conv', dict(Cin=64, Cout=64, S=16, k=1, impl=T)),
    ('conv', dict(Cin=32, Cout=32, S=32, impl=T, masked=True)), ('conv', dict(Cin=64, Cout=64, S=16, impl=T, masked=True)),
    ('conv', dict(Cin=128, Cout=128, S=8, impl=T, masked=True)), ('conv', dict(Cin=32, Cout=64, S=32, stride=2, impl=T, masked=True)),
    ('conv', dict(Cin=32, Cout=64, S=32, k=1, stride=2, impl=T, masked=True)),
    ('convT', dict(Cin=64, Cout=64, S=8, impl=T)), ('convT', dict(Cin=512, Cout=512, S=4, impl=T)),
    ('convT', dict(Cin=128, Cout=128, S=8, impl=T)), ('convT', dict(Cin=32, Cout=32, S=8, impl=T)),
    ('conv_stats', {}),
    # halo-plane kernel (v2): dense 3x3x3 s1, Cin % 64 == 0, H >= 16
    ('conv', dict(Cin=64, Cout=64, S=16, impl=T)), ('conv', dict(Cin=64, Cout=64, S=32, impl=T)),
    ('conv', dict(Cin=128, Cout=64, S=16, impl=T)), ('conv', dict(Cin=64, Cout=128, S=32, impl=T, bias=False)),
    ('conv', dict(Cin=256, Cout=256, S=16, impl=T)), ('conv', dict(Cin=64, Cout=32, S=24, impl=T)),
    ('conv', dict(Cin=64, Cout=64, S=32, impl=T, masked=True, f=8)), ('conv', dict(Cin=512, Cout=512, S=16, N=1, impl=T)),
    ('conv_stats', dict(Cin=64, Cout=64, S=32)),
    ('conv', dict(Cin=32, Cout=32, S=32, impl=T)), ('conv', dict(Cin=64, Cout=16, S=20, impl=T)),
    ('conv', dict(Cin=128, Cout=64, S=32, N=1, impl=T)),
]

if __name__ == '__main__':
    flt = sys.argv[1] if len(sys.argv) > 1 else ''
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    log = open(os.path.join(ROOT, 'gpurun_out', 'bringup.log'), 'a')
    npass = nfail = 0
    for name, kw in CASES:
        if flt and flt not in name:
            continue
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, '-m', 'tests.kernel_checks', name, json.dumps(kw)], cwd=ROOT,
                               capture_output=True, text=True, timeout=180)
            ok = p.returncode == 0
            tail = (p.stdout + p.stderr).strip().splitlines()[-6:]
        except subprocess.TimeoutExpired:
            ok, tail = False, ['TIMEOUT']
        npass += ok
        nfail += not ok
        line = f"{'PASS' if ok else 'FAIL'} {name} {json.dumps(kw)} ({time.time() - t0:.1f}s)"
        print(line)
        log.write(line + '\n')
        if not ok:
            for l in tail:
                print('    ' + l[:300])
                log.write('    ' + l[:300] + '\n')
        else:
            print('    ' + tail[-1][:300])
            log.write('    ' + tail[-1][:300] + '\n')
        log.flush()
    print(f'{npass} passed, {nfail} failed')
