"""Run an UNMODIFIED reference script (src/train.py, src/eval.py, src/inference.py) on the gaviko_b200 drop-in modules:

    python -m gaviko_b200.launch /path/to/GAViKO/src/train.py --config cfg.yaml --method gaviko --results_dir out

`sys.path` is ordered [gaviko_b200/dropin, <script dir>, ...] so that the reference's `from model.gaviko import Gaviko` (namespace packages, no
__init__.py anywhere under src/) resolves to the drop-in while `utils.logging`, `data.dataset` stay the reference's own files; the script then
runs as `__main__` with its own argv.  Optional environment: GAVIKO_COMPUTE_DTYPE = fp32 | bf16 selects the compute mode for every model the
script constructs (default: follow the parameter dtype, i.e. exact fp32 kernels for the scripts' fp32 models).
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DROPIN = os.path.join(HERE, 'dropin')


def prepare(script):
    script = os.path.abspath(script)
    root = os.path.dirname(HERE)
    src = os.path.dirname(script)
    for p in (root, src, DROPIN):                 # final order: dropin, script dir, repo root, rest
        while p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in [m for m in sys.modules if m.split('.')[0] in ('model', 'losses', 'utils', 'data')]:
        del sys.modules[name]                     # a previous import under another path order must not leak in
    return script


def run(script, argv):
    script = prepare(script)
    old = sys.argv
    sys.argv = [script] + list(argv)
    try:
        return runpy.run_path(script, run_name='__main__')
    finally:
        sys.argv = old


if __name__ == '__main__':
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    run(sys.argv[1], sys.argv[2:])
