"""NN-mode stencil / shift generator.

Same contract as tedeous/finite_diffs.py (First_order_scheme 9-117, Second_order_scheme 120-223,
Finite_diffs.scheme_choose 244-268): a derivative multi-index such as [0, 0] (= d2/dx0^2) becomes a list
of integer shift vectors (in units of the step h) and a list of matching weights.  The implementation is
a single table-driven expansion: every differentiation step replaces each (shift, weight) pair by the
pairs of a 1-D rule applied along that axis.

Only the *setup* of NN mode uses this.  For interior ('central') points the fused kernel evaluates the
exact derivative jets instead of the 2^k shifted forwards (the central rule is consistent to O(h^2)); for
one-sided boundary operators the literal shifted evaluations are kept, because the reference's
"second-order" rule (3u(x+-2h) - 4u(x+-h) + u(x)) / (+-2h) is u' + 2h u'' and not u' (SURVEY B.1 q3)."""
from typing import List, Tuple


def _rule(direction: str, variant: str, h: float) -> List[Tuple[int, float]]:
    """1-D first-derivative rule as (shift, weight-multiplier-for-sign=+1) pairs, in the reference's order."""
    if direction == 'central':
        c = 1 / (2 * h)
        return [(1, c), (-1, -c)]
    if variant == '1':
        if direction == 'f':
            return [(1, 1 / h), (0, -1 / h)]
        if direction == 'b':
            return [(0, 1 / h), (-1, -1 / h)]
    elif variant == '2':
        c = 1 / (2 * h)
        if direction == 'f':
            return [(2, 3 * c), (1, -4 * c), (0, c)]
        if direction == 'b':
            return [(-2, -3 * c), (-1, 4 * c), (0, -c)]
    raise ValueError(f'no finite-difference rule for direction={direction!r} variant={variant!r}')


class Finite_diffs:
    """scheme_choose(scheme_label, h) -> [shifts, signs]   (tedeous/finite_diffs.py:226-268)."""

    def __init__(self, term: list, nvars: int, axes_scheme_type: str):
        self.term = term
        self.nvars = nvars
        self.axes_scheme_type = axes_scheme_type

    def directions(self) -> List[str]:
        if self.axes_scheme_type == 'central':
            return ['central'] * len(self.term)
        return [self.axes_scheme_type[a] for a in self.term]

    def scheme_choose(self, scheme_label: str, h: float = 1 / 2) -> list:
        if self.term == [None]:
            return [[None], [1]]
        if scheme_label not in ('1', '2'):
            raise ValueError("scheme_label must be '1' or '2'")
        shifts = [[0] * self.nvars]
        signs = [1]
        for axis, direction in zip(self.term, self.directions()):
            # the reference's Second_order_scheme is only defined for one-sided points
            rule = _rule(direction, scheme_label if direction != 'central' else '1', h)
            new_shifts, new_signs = [], []
            for s, w in zip(shifts, signs):
                for delta, mult in rule:
                    s2 = list(s)
                    s2[axis] += delta
                    new_shifts.append(s2)
                    new_signs.append(_times(w, mult))
            shifts, signs = new_shifts, new_signs
        return [shifts, signs]


def _times(w, mult):
    # keeps the reference's arithmetic order ("sign * (1 / (2 * h))") so weights agree to the last bit
    return w * mult
