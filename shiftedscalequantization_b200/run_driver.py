"""Run one of the reference's driver scripts UNCHANGED against the B200-native implementation:

    python -m shiftedscalequantization_b200.run_driver /path/to/main_cifar10.py --iters_w 2000 ...

The script's own directory normally shadows everything on sys.path (so `from quant import *` would pick up the
reference's pure-PyTorch package next to it); this launcher puts compat/ first and removes the script directory."""
import os
import runpy
import sys

COMPAT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat")


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    sys.argv = [script] + sys.argv[2:]
    root = os.path.dirname(COMPAT)
    sys.path[:] = [COMPAT, root] + [p for p in sys.path if os.path.abspath(p or ".") not in (os.path.dirname(script),)]
    for stale in [m for m in sys.modules if m == "quant" or m.startswith("quant.")]:
        del sys.modules[stale]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
