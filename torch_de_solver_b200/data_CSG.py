"""Point-mask shapes for domains with holes (tedeous/data_CSG.py).  Grid construction only: the fused
path takes an arbitrary [N, d] point list, so CSG grids need no kernel support."""
import torch


class Rectangle:
    def __init__(self, lower, upper, dims=None):
        self.lower = torch.as_tensor(lower, dtype=torch.float32)
        self.upper = torch.as_tensor(upper, dtype=torch.float32)
        self.dims = dims if dims is not None else list(range(self.lower.numel()))

    def contains(self, pts):
        sub = pts[:, self.dims]
        lo, up = self.lower.to(sub.device), self.upper.to(sub.device)
        return ((sub >= lo) & (sub <= up)).all(dim=1)

    def boundary(self, pts, rtol=1e-4, atol=0.0):
        sub = pts[:, self.dims]
        lo, up = self.lower.to(sub), self.upper.to(sub)
        on_lo = torch.isclose(sub, lo.unsqueeze(0), rtol=rtol, atol=atol).any(dim=1)
        on_up = torch.isclose(sub, up.unsqueeze(0), rtol=rtol, atol=atol).any(dim=1)
        return self.contains(pts) & (on_lo | on_up)


class Circle:
    def __init__(self, center, radius, dims=None):
        self.center = torch.as_tensor(center, dtype=torch.float32)
        self.radius_sq = float(radius) ** 2
        self.dims = dims if dims is not None else list(range(self.center.numel()))

    def _sqd(self, pts):
        sub = pts[:, self.dims]
        return ((sub - self.center.to(sub.device)) ** 2).sum(dim=1)

    def contains(self, pts):
        return self._sqd(pts) <= self.radius_sq

    def boundary(self, pts, rtol=1e-4, atol=0.0):
        sqd = self._sqd(pts)
        return torch.isclose(sqd, torch.tensor(self.radius_sq, dtype=pts.dtype, device=pts.device),
                             rtol=rtol, atol=atol)


def csg_difference(grid, shape):
    return grid[~shape.contains(grid)]


def csg_boundary(grid, shape, rtol=1e-4, atol=0.0):
    return grid[shape.boundary(grid, rtol=rtol, atol=atol)]
