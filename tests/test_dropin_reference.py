"""The drop-in boundary, tested with the reference's own code: its UNMODIFIED `train.py`, `translate.py`
and `trainer/trainer_st.py` are imported with this repo's modules in front (b200st/dropin.py), `Trainer_ST(expt_dir=...)`
is constructed, its `_train_batch` runs three optimizer steps and `translate.translate` writes its text file for greedy /
beam-3 and both histories (HYP = forward_translate, REF = forward_translate_refen) — and everything is compared with the
SAME script running on the reference's own modules (tests/dropin_driver.py, one subprocess per implementation).

The reference tree is /root/reference in the build container, or the staged copy oracle/_ref/ (oracle/build_ref.sh)
which travels to the GPU box; without either the tests skip.  CPU: kernels replaced by tests/fake_kernels.py (host logic,
import wiring, optimizer protocol).  GPU: the real sm_100a kernels in fp32 mode against the reference on stock PyTorch CUDA.
"""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200')


def _reference_root():
    for cand in ('/root/reference', os.path.join(ROOT, 'oracle', '_ref')):
        if os.path.isfile(os.path.join(cand, 'trainer', 'trainer_st.py')):
            return cand
    return None


REF = _reference_root()
needs_ref = pytest.mark.skipif(REF is None, reason='no reference tree (/root/reference or oracle/_ref)')


def _drive(impl, out, device):
    os.makedirs(out, exist_ok=True)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1')
    env.pop('PYTHONPATH', None)
    r = subprocess.run([sys.executable, os.path.join(HERE, 'dropin_driver.py'), '--impl', impl, '--reference', REF,
                        '--out', out, '--device', device], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-4000:]
    return json.load(open(os.path.join(out, 'result.json')))


def _compare(ref, got, tol):
    # which file every module name resolved to: the reference's own for everything but the hot-path overrides
    for name in ('train', 'translate', 'trainer.trainer_st', 'trainer.trainer_base', 'utils.misc', 'utils.dataset',
                 'modules.checkpoint'):
        assert got['origin'][name] == ref['origin'][name], name
        assert os.path.realpath(REF) in got['origin'][name], name
    for name in ('modules.loss', 'modules.optim', 'modules.layers', 'models.Seq2seq', 'models.Dec'):
        assert got['origin'][name].startswith(os.path.realpath(PKG)), (name, got['origin'][name])
    # Trainer_ST._train_batch: first loss == the golden (pre-update), all three == the reference's, weights after 3 steps
    assert abs(ref['losses'][0] - ref['golden_loss']) < tol * abs(ref['golden_loss'])
    for a, b in zip(got['losses'], ref['losses']):
        assert abs(a - b) < tol * abs(b), (got['losses'], ref['losses'])
    assert ref['losses'][2] < ref['losses'][0]
    for n, w in ref['wsum'].items():
        assert abs(got['wsum'][n] - w) <= 10 * tol * max(w, 1e-6), n
    # translate.translate: identical output files
    for key, lines in ref['translate'].items():
        assert got['translate'][key] == lines, key


@needs_ref
def test_reference_entry_points_run_on_repo_modules_cpu(tmp_path):
    ref = _drive('reference', str(tmp_path / 'ref'), 'cpu')
    got = _drive('b200', str(tmp_path / 'b200'), 'cpu')
    _compare(ref, got, 2e-5)


@needs_ref
@pytest.mark.gpu
def test_reference_entry_points_run_on_cuda_kernels(tmp_path):
    ref = _drive('reference', str(tmp_path / 'ref'), 'cpu')       # the reference's own CPU path
    got = _drive('b200', str(tmp_path / 'b200'), 'cuda')          # unmodified trainers / translate on libb200st.so
    _compare(ref, got, 1e-4)


@needs_ref
def test_loss_module_inherits_the_rest_from_the_reference(tmp_path):
    code = ("import sys, types; sys.dont_write_bytecode = True\n"
            f"sys.path.insert(0, {PKG!r})\n"
            "from b200st import dropin\n"
            "import modules.loss as L0\n"
            "assert not hasattr(L0, 'BCELoss')\n"                  # standalone: hot-path names only
            f"dropin.install({REF!r})\n"
            "from modules.loss import NLLLoss, BCELoss, CrossEntropyLoss, KLDivLoss, MSELoss\n"
            "import modules.loss as L, modules.checkpoint as C\n"
            f"assert L.__file__.startswith({PKG!r}) and L.__shadowed_file__.startswith({os.path.realpath(REF)!r})\n"
            f"assert C.__file__.startswith({os.path.realpath(REF)!r})\n"
            "assert NLLLoss.__module__ == 'modules.loss' and BCELoss().name == 'BCELoss'\n"
            "print('ok')\n")
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-2000:]
