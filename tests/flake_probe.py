"""Repeats the model-level parity checks to expose tolerance margins that sit inside the run-to-run noise of the
atomics-based reductions.   python tests/flake_probe.py [repeats]"""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import model_checks as mc  # noqa: E402

if __name__ == '__main__':
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    for name, seed in [('tiny', 3), ('S64', 5), ('B64', 5), ('S_aniso', 4), ('L32', 2)]:
        for r in range(reps):
            try:
                mc.check_spark(name, seed=seed, verbose=False)
            except AssertionError as e:
                print('FAIL spark', name, r, str(e)[:600])
            except Exception:
                traceback.print_exc()
    for r in range(reps):
        try:
            mc.check_anatomask_steps()
            print('ok anatomask_steps', r)
        except AssertionError as e:
            print('FAIL anatomask_steps', r, str(e)[:600])
