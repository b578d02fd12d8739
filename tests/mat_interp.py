"""Test helper: evaluates a lowered mat-mode problem (torch_de_solver_b200.mat.MatIR) with dense fp64 operators on the
CPU.  It mirrors what the CUDA kernels do with the IR - banded derivative fields on the (extended) slab, term table,
residual, loss over the owned rows, gradient of the owned rows from the seeds of the rows around them - so the
*lowering*, the slab decomposition and the halo exchange can be validated without a GPU.  It lives under tests/ on
purpose: the product has no CPU path."""
import numpy as np
import torch

from test_mat_cpu import dense_from_band


def _fields(ir, u):
    n_var, n_ext, n1 = ir.shape_ext
    out = []
    for q, (var, axis, order) in enumerate(ir.fields):
        if order == 0:
            out.append(u[var])
            continue
        f = ir.fld[q]
        b, E, off = int(f['half_width']), int(f['n_edge']), int(f['coef_off'])
        n = (n_ext, n1)[axis]
        w = 2 * b + 1
        band = ir.band[off:off + w * (1 + 2 * E)].astype(np.float64)
        D = torch.as_tensor(dense_from_band(band, b, E, n))
        out.append(D @ u[var] if axis == 0 else u[var] @ D.T)
    return out


def _terms(ir, F, tb, te, coeffs, cells=None):
    n_var, n_ext, n1 = ir.shape_ext
    val = 0.0
    for t in range(tb, te):
        tm = ir.terms[t]
        if tm['kind'] == 1:
            c = coeffs[int(tm['idx']):int(tm['idx']) + n_ext * n1].reshape(n_ext, n1)
        else:
            c = float(tm['coeff'])
        prod = c if cells is None or tm['kind'] != 1 else c.reshape(-1)[cells]
        for fi in range(int(tm['fac_begin']), int(tm['fac_end'])):
            fc = ir.factors[fi]
            x = F[int(fc['chan'])]
            if cells is not None:
                x = x.reshape(-1)[cells]
            prod = prod * x ** float(fc['pow'])
        val = val + prod
    return val


def evaluate_mat_ir(ir, u_ext):
    """u_ext: [n_var, n_ext, n1] float64.  -> (out [2 + n_slots] partial sums of this rank, gradient of the owned rows)"""
    u = u_ext.detach().double().clone().requires_grad_(True)
    n_var, n_ext, n1 = ir.shape_ext
    (r0, r1), (e0, e1) = ir.rows, ir.ext
    lo, hi = r0 - e0, r1 - e0
    s_lo, s_hi = max(0, lo - ir.reach0), min(n_ext, hi + ir.reach0)     # rows whose seeds reach the owned rows
    coeffs = ir.coeffs.detach().double().cpu()
    F = _fields(ir, u)
    lam, ln = ir.slot_lambda, ir.slot_len
    mse, obj = [], 0.0
    for e, (tb, te) in enumerate(ir.eq_ranges):
        r = _terms(ir, F, tb, te, coeffs)
        if not torch.is_tensor(r):
            r = torch.full((n_ext, n1), float(r), dtype=torch.float64)
        mse.append((r[lo:hi] ** 2).sum() / ln[e])
        obj = obj + lam[e] * (r[s_lo:s_hi] ** 2).sum() / ln[e]
    cells_all = ir.cells.cpu().long()
    tg_all = ir.targets.detach().double().cpu()
    sums = [torch.zeros((), dtype=torch.float64) for _ in ir.bnd_types]
    for bc in ir.bcs:
        n, K = int(bc['n_rows']), int(bc['K'])
        if n == 0:
            continue
        cells = cells_all[int(bc['cell_off']):int(bc['cell_off']) + n * K].reshape(n, K)
        val = 0.0
        for k in range(K):
            if bc['term_begin'] == bc['term_end']:
                v = u[int(bc['var'])].reshape(-1)[cells[:, k]]
            else:
                v = _terms(ir, F, int(bc['term_begin']), int(bc['term_end']), coeffs, cells[:, k])
            val = val + float(bc['sign'][k]) * v
        res = val - tg_all[int(bc['tgt_off']):int(bc['tgt_off']) + n]
        sums[int(bc['slot'])] = sums[int(bc['slot'])] + (res ** 2).sum()
    n_eq = ir.n_eq
    bm = [s / ln[n_eq + i] for i, s in enumerate(sums)]
    obj = obj + sum(lam[n_eq + i] * m for i, m in enumerate(bm))
    g, = torch.autograd.grad(obj, u)
    loss = sum(lam[e] * mse[e] for e in range(n_eq)) + sum(lam[n_eq + i] * m for i, m in enumerate(bm))
    loss_n = sum(mse) + sum(bm)
    out = torch.stack([loss, loss_n] + mse + bm).detach()
    return out, g[:, lo:hi].detach()
