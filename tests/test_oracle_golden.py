"""Pins the oracle (oracle/tedeous_oracle.py) AND this repo's front end (Domain / Conditions / Equation)
to the golden fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import problems
from helpers import load_golden, oracle_eval

CASES = sorted(problems.ZOO)


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_fp64(name):
    g = load_golden(name, 'float64')
    sol, loss, loss_n, grad = oracle_eval(name, 'float64', g['weights'])
    assert sol.op.shape[0] == int(g['op_rows'])
    assert list(sol.bval_keys) == [str(k) for k in g['bval_keys']]
    assert list(sol.bval_length) == [int(x) for x in g['bval_length']]
    # same algorithm, same dtype, same torch: agreement to fp64 rounding
    assert loss == pytest.approx(float(g['loss']), rel=1e-11)
    assert loss_n == pytest.approx(float(g['loss_normalized']), rel=1e-11)
    np.testing.assert_allclose(sol.op.detach().numpy()[:256], g['op_head'], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(sol.bval.detach().numpy(), g['bval'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(sol.true_bval.detach().numpy(), g['true_bval'], rtol=1e-12, atol=1e-14)
    assert np.linalg.norm(grad - g['grad']) <= 1e-9 * np.linalg.norm(g['grad'])


@pytest.mark.parametrize('name', [c for c in CASES if c not in problems.LARGE])
def test_oracle_matches_reference_fp32(name):
    g = load_golden(name, 'float32')
    sol, loss, loss_n, grad = oracle_eval(name, 'float32', g['weights'])
    # fp32: identical call sequence -> differences only from summation order inside torch kernels
    assert loss == pytest.approx(float(g['loss']), rel=2e-6)
    assert loss_n == pytest.approx(float(g['loss_normalized']), rel=2e-6)
    assert np.linalg.norm(grad - g['grad']) <= 2e-5 * np.linalg.norm(g['grad'])
