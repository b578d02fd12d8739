"""Host logic of mat mode: the banded derivative operators built in torch_de_solver_b200/mat.py must equal the
oracle's (= the reference's) rolled finite differences, including the special edge rows, and boundary points
must map to the right cells."""
import numpy as np
import pytest
import torch

from oracle.tedeous_oracle import MatDerivative
from torch_de_solver_b200.mat import cell_indices, derivative_band, first_derivative_matrix


def dense_from_band(band, b, E, n):
    w = 2 * b + 1
    interior, lo, hi = band[:w], band[w:w + E * w].reshape(E, w), band[w + E * w:].reshape(E, w)
    D = np.zeros((n, n))
    for i in range(n):
        row = lo[i] if i < E else hi[n - 1 - i] if i >= n - E else interior
        for m in range(-b, b + 1):
            if 0 <= i + m < n:
                D[i, i + m] = row[m + b]
    return D


@pytest.mark.parametrize('p', [2, 3, 4])
@pytest.mark.parametrize('order', [1, 2, 3])
@pytest.mark.parametrize('n', [9, 16, 40])
def test_band_equals_reference_operator(p, order, n):
    if n < 2 * p:
        pytest.skip('axis too short')
    h = 0.125
    band, b, E = derivative_band(n, p, order, h)
    D = dense_from_band(band.astype(np.float64), b, E, n)
    md = MatDerivative(p)
    # apply the oracle's first derivative `order` times to the identity, along the last axis of a [2, n] field
    eye = torch.eye(n, dtype=torch.float64)
    cols = []
    for j in range(n):
        u = torch.stack([eye[j], 2 * eye[j]])          # 2 rows so the 2-D code path is taken
        for _ in range(order):
            u = md.d1(u, torch.tensor(h, dtype=torch.float64), 1)
        cols.append(u[0].numpy())
    ref = np.stack(cols, 1)
    np.testing.assert_allclose(D, ref, rtol=2e-6, atol=2e-6 * np.abs(ref).max())


def test_first_derivative_rows_p2():
    D = first_derivative_matrix(6, 2)
    np.testing.assert_allclose(D[0, :2], [-1, 1])
    np.testing.assert_allclose(D[2, 1:4], [-0.5, 0, 0.5])
    np.testing.assert_allclose(D[5, 4:], [-1, 1])


def test_cell_indices():
    x = torch.linspace(0, 1, 5)
    y = torch.linspace(-1, 1, 9)
    grid = torch.stack(torch.meshgrid(x, y, indexing='ij'))
    bnd = torch.tensor([[0.25, -1.0], [1.0, 0.5], [0.0, 1.0]])
    idx = cell_indices(grid, bnd)
    assert idx.tolist() == [1 * 9 + 0, 4 * 9 + 6, 0 * 9 + 8]
    with pytest.raises(ValueError):
        cell_indices(grid, torch.tensor([[0.3, 0.0]]))


@pytest.mark.parametrize('name', ['legendre_mat_1d', 'lotka_mat_1d'])
def test_one_dimensional_grid_is_lifted(name):
    """1-D mat-mode grids ([1, N0], the ODE examples) are lowered as [N0, 1] grids with a dummy second axis: the dense
    fp64 interpretation of the lowered problem reproduces the reference's loss and gradient."""
    import numpy as np
    from mat_interp import evaluate_mat_ir
    from test_distributed_cpu import _mat_ir
    g, ir, u = _mat_ir(name, (0, 1))
    assert ir.lifted and ir.shape_ext[2] == 1
    out, grad = evaluate_mat_ir(ir, u.unsqueeze(-1))
    # (the lowering divides by the fp32 grid step the fp32 reference computes, derivative.py:229-247; h = 1/48 is not a
    # binary fraction, so the fp64 fixture differs at the 1e-8 level)
    assert float(out[0]) == pytest.approx(float(g['loss']), rel=1e-6)
    gref = g['grad']
    assert np.linalg.norm(grad.reshape(-1).numpy() - gref) <= 1e-6 * np.linalg.norm(gref)


def test_robin_rows_in_mat_mode():
    """`robin` conditions in mat mode (tedeous/eval.py:357-388, quirk q5: the alpha term counted again in every beta term)
    are lowered as boundary-operator rows; the dense fp64 interpretation of the IR reproduces the reference fixture."""
    import numpy as np
    from mat_interp import evaluate_mat_ir
    from test_distributed_cpu import _mat_ir
    g, ir, u = _mat_ir('poisson_robin_mat', (0, 1))
    assert ir.bnd_types == ['robin', 'dirichlet']
    out, grad = evaluate_mat_ir(ir, u)
    assert float(out[0]) == pytest.approx(float(g['loss']), rel=1e-6)
    gref = g['grad']
    assert np.linalg.norm(grad.reshape(-1).numpy() - gref) <= 1e-6 * np.linalg.norm(gref)
