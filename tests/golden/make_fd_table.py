"""Dump Finite_diffs.scheme_choose of the UNMODIFIED reference (tedeous/finite_diffs.py:244-268) for a sweep of
(term, nvars, axes_scheme_type, scheme order, h) into tests/golden/finite_diffs_table.json.
Run in the build container:  python tests/golden/make_fd_table.py"""
import itertools
import json
import os
import sys

sys.path.insert(0, '/root/reference')
from tedeous.finite_diffs import Finite_diffs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    for nvars in (1, 2, 3):
        terms = [[None]]
        for k in (1, 2, 3, 4):
            for t in itertools.product(range(nvars), repeat=k):
                if k <= 2 or len(set(t)) == 1 or (k == 3 and nvars == 2):
                    terms.append(list(t))
        types = ['central'] + [''.join(c) for c in itertools.product('fb', repeat=nvars)]
        for term, typ, order, h in itertools.product(terms, types, ('1', '2'), (0.5, 0.01, 0.001)):
            if typ == 'central' and order == '2':
                continue          # Second_order_scheme is only defined for one-sided points
            yield term, nvars, typ, order, h


def main():
    out = []
    for term, nvars, typ, order, h in cases():
        res = Finite_diffs(term, nvars, typ).scheme_choose(order, h=h)
        out.append({'term': term, 'nvars': nvars, 'type': typ, 'order': order, 'h': h, 'scheme': res[0], 'sign': res[1]})
    with open(os.path.join(HERE, 'finite_diffs_table.json'), 'w') as f:
        json.dump(out, f, separators=(',', ':'))
    print(len(out), 'cases')


if __name__ == '__main__':
    main()
