"""CPU: the script-equivalent sequence of tests/test_script_equivalent_gpu.py with the kernels replaced by their torch restatement
(tests/ops_double.py), so the host side of that sequence is covered in the `-m "not gpu"` suite as well."""
import pytest

import ops_double
import test_script_equivalent_gpu as T


@pytest.mark.parametrize('method', ['gaviko', 'dvpt', 'deep_vpt', 'melo', 'bitfit', 'evp'])
def test_train_eval_checkpoint_sequence_host_side(method, tmp_path, monkeypatch):
    monkeypatch.setattr(T, 'DEVICE', 'cpu')
    with ops_double.install():
        T.test_train_eval_checkpoint_sequence(method, tmp_path)
