"""Drop-in wiring: run the reference's own, UNMODIFIED entry points (`train.py`, `translate.py`, `trainer/*.py`) on the
modules of this package.

The reference has no plugin / FFI layer; its boundary is the Python module API.  This package ships
modules under the reference's own import names for exactly the files on the hot path,

    models/{Seq2seq,Las,Enc,Dec,TFEnc,TFDec}.py    modules/{layers,attention,loss,optim}.py

and nothing under `utils/` or `trainer/`.  `models/__init__.py` and `modules/__init__.py` extend their `__path__`
with the same-named directories found further along `sys.path`, so with the reference root BEHIND this package on
`sys.path` every other name keeps resolving to the reference's file: `modules.checkpoint`, `models.Act`,
`utils.misc`, `utils.dataset`, `utils.config`, `trainer.trainer_st`, ...  Where this package overrides a module
that the reference's callers import MORE names from than the hot path needs (`modules/loss.py`: the trainers import
BCELoss / CrossEntropyLoss / KLDivLoss / MSELoss next to NLLLoss, trainer_st.py:15, translate.py:19), the module calls
`inherit_shadowed()` to take the remaining names from the reference's shadowed file.

Use:
    python -m b200st.dropin --reference /path/to/reference [--dtype bf16] train.py --train_path_src ...
    python -m b200st.dropin --reference /path/to/reference translate.py --test_path_src ...
or, from Python, `b200st.dropin.install('/path/to/reference')` before importing `train` / `translate` / `trainer.*`.
tests/test_dropin_reference.py drives `Trainer_ST._train_batch` and `translate.translate` of the unmodified
reference this way and compares them with the reference running on its own modules.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from typing import List, Optional

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # holds models/, modules/, b200st/
OVERRIDDEN_PACKAGES = ('models', 'modules')


def extend_path(path: List[str], name: str) -> List[str]:
    """`__path__` of package `name` + every `<sys.path entry>/<name>` directory not already on it (in sys.path order).
    Like pkgutil.extend_path, but it also takes directories without an `__init__.py` (the reference's packages are
    namespace packages)."""
    out = list(path)
    seen = {os.path.realpath(p) for p in out}
    for entry in sys.path:
        if not isinstance(entry, str):
            continue
        cand = os.path.join(entry or os.getcwd(), *name.split('.'))
        if os.path.isdir(cand) and os.path.realpath(cand) not in seen:
            seen.add(os.path.realpath(cand))
            out.append(cand)
    return out


def inherit_shadowed(module_globals: dict) -> Optional[str]:
    """Called at the END of a module of this package that shadows a same-named reference module: executes the
    reference's file (the next one along the parent package's `__path__`) under a private name and copies every public
    name this module did not define itself.  Returns the file it took them from (None: no reference on the path, the
    module then offers the hot-path names only)."""
    name = module_globals['__name__']
    pkg_name, _, leaf = name.rpartition('.')
    pkg = sys.modules.get(pkg_name)
    if pkg is None:
        return None
    here = os.path.realpath(module_globals['__file__'])
    for d in extend_path(list(pkg.__path__), pkg_name):
        cand = os.path.join(d, leaf + '.py')
        if os.path.isfile(cand) and os.path.realpath(cand) != here:
            spec = importlib.util.spec_from_file_location(f'_b200st_shadowed.{name}', cand)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            for k, v in vars(mod).items():
                if not k.startswith('__') and k not in module_globals:
                    module_globals[k] = v
            module_globals['__shadowed_file__'] = cand
            return cand
    return None


def install(reference_root: str) -> None:
    """Put this package in front of, and the reference root behind it on, sys.path; re-extend the `__path__` of the
    overridden packages if they were imported already.  Idempotent."""
    reference_root = os.path.abspath(reference_root)
    if not os.path.isdir(os.path.join(reference_root, 'models')):
        raise FileNotFoundError(f'{reference_root} does not look like the reference checkout (no models/)')
    for p in (reference_root, _PKG_ROOT):
        while p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [_PKG_ROOT, reference_root]
    sys.dont_write_bytecode = True            # never write .pyc files into the reference checkout
    for name in OVERRIDDEN_PACKAGES:
        pkg = sys.modules.get(name)
        if pkg is not None:
            if not os.path.realpath(getattr(pkg, '__file__', '') or '').startswith(os.path.realpath(_PKG_ROOT)):
                raise RuntimeError(f"'{name}' was already imported from {getattr(pkg, '__file__', pkg.__path__)}: call "
                                   f'b200st.dropin.install() before importing the reference')
            pkg.__path__ = extend_path(list(pkg.__path__), name)
    # modules that shadow a reference file and were imported before the reference was on the path: take the rest now
    for name in ('modules.loss',):
        mod = sys.modules.get(name)
        if mod is not None and '__shadowed_file__' not in vars(mod):
            inherit_shadowed(vars(mod))


def main(argv=None):
    import argparse
    import runpy
    ap = argparse.ArgumentParser(prog='python -m b200st.dropin', description=__doc__.split('\n\n')[0])
    ap.add_argument('--reference', required=True, help='root of the reference checkout (holds train.py, translate.py)')
    ap.add_argument('--dtype', default=os.environ.get('B200ST_DTYPE', 'bf16'), choices=['bf16', 'fp32'])
    ap.add_argument('script', help='reference entry point, e.g. train.py or translate.py')
    ap.add_argument('args', nargs=argparse.REMAINDER)
    ns = ap.parse_args(argv)
    install(ns.reference)
    from b200st import runtime
    runtime.set_compute_dtype(ns.dtype)
    script = ns.script if os.path.isabs(ns.script) else os.path.join(os.path.abspath(ns.reference), ns.script)
    sys.argv = [script] + ns.args
    runpy.run_path(script, run_name='__main__')


if __name__ == '__main__':
    main()
