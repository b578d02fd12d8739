"""Test helper: executes a lowered ProblemIR (torch_de_solver_b200.plan) with plain torch autograd on the CPU.

It mirrors what the CUDA kernel does with the IR (jets -> virtual channels -> term table -> residual ->
slot sums) so the *lowering* can be validated against the golden fixtures without a GPU.  It lives under
tests/ on purpose: the product has no CPU path."""
import torch

from torch_de_solver_b200.plan import ProblemIR


def _jets(model, pts, jet):
    """[n_pts, J, n_out] values of every jet channel."""
    pts = pts.clone().requires_grad_(True)
    out = model(pts)
    n_out = out.shape[1]
    chans = [out]
    for i, (_, order) in enumerate(jet.dirs):
        vec = torch.as_tensor(jet.vector(i, pts.shape[1]), dtype=pts.dtype)     # unit vector for a pure partial
        cur = out
        for _ in range(order):
            cols = []
            for v in range(n_out):
                g, = torch.autograd.grad(cur[:, v].sum(), pts, create_graph=True)
                cols.append(g @ vec)
            cur = torch.stack(cols, 1)
            chans.append(cur)
    return torch.stack(chans, 1)


def evaluate_ir(ir: ProblemIR, model, dtype=torch.float64, tol=0):
    """-> (loss, loss_normalized, slot_mse list, fields per segment).  tol != 0: causal weights on the rows of
    segment 0 (what Solution._causal_weights feeds the kernels; tedeous/losses.py:137-182), lambda_operator unused."""
    sums = [torch.zeros((), dtype=dtype) for _ in range(ir.n_slots)]
    fields = []
    for s in ir.segments:
        if s.n_groups == 0:          # a rank may own no row of a small segment
            fields.append(torch.zeros(0, len(s.cols), dtype=dtype))
            continue
        pts = s.points.to(dtype)
        J, K, M = s.jet.J, s.K, s.M
        jets = _jets(model, pts, s.jet)                           # [n*K, J, n_out]
        n = s.n_groups
        jets = jets.reshape(n, K * J, -1)
        if s.comb is None:
            V = jets
        else:
            V = torch.einsum('mq,nqv->nmv', torch.as_tensor(s.comb, dtype=dtype), jets)
        cols = []
        for terms, slot in zip(s.cols, s.slots):
            val = torch.zeros(n, dtype=dtype)
            for t in terms:
                c = t.coeff
                prod = c.to(dtype).reshape(-1) if isinstance(c, torch.Tensor) else torch.full((n,), float(c), dtype=dtype)
                for f in t.factors:
                    prod = prod * V[:, s.chan_of(f), f.var] ** f.pow
                val = val + prod
            cols.append(val)
        vals = torch.stack(cols, 1)
        if s is ir.segments[0] and getattr(ir, 'host_terms', None):
            # callable-'pow' terms: assembled from the factor columns as Solution._assemble_host does
            eqs = [vals[:, e] for e in range(ir.n_eq)]
            for e, coeff, chain in ir.host_terms:
                der = 1.
                for col, pw in chain:
                    der = pw(der * vals[:, col]) if callable(pw) else der * vals[:, col] ** pw
                c = coeff.to(dtype).reshape(-1) if isinstance(coeff, torch.Tensor) and coeff.numel() > 1 else coeff
                eqs[e] = eqs[e] + c * der
            vals = torch.stack(eqs, 1)
        fields.append(vals)
        res = vals - (s.targets.to(dtype) if s.targets is not None else 0.)
        w = 1.
        if tol != 0 and s is ir.segments[0]:
            with torch.no_grad():
                n_t = ir.n_t if n % ir.n_t == 0 else n
                r2 = (vals.detach() ** 2).sum(1).reshape(n_t, -1)
                w = torch.exp(-tol * (torch.cumsum(r2, 0) - r2)).reshape(-1)
        for ci, slot in enumerate(s.slots[:vals.shape[1]]):
            sums[slot] = sums[slot] + (w * res[:, ci] ** 2).sum()
    mse = [sm / ln for sm, ln in zip(sums, ir.slot_len)]
    lam = [1. if (tol != 0 and i < ir.n_eq) else l for i, l in enumerate(ir.slot_lambda)]
    loss = sum(l * m for l, m in zip(lam, mse))
    loss_n = sum(mse)
    return loss, loss_n, mse, fields
