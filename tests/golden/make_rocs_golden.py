"""Golden vectors for the shape Tanimoto from the UNMODIFIED reference functions VAB_2nd_order /
shape_tanimoto / get_ROCS (utils/evaluation/shaep_utils.py:59-83).  The module itself imports RDKit at the top, which this
image does not have, so the three (pure torch) function definitions are compiled from the reference source file at
generation time; nothing is copied into the repository.

Run in the build container only:   python tests/golden/make_rocs_golden.py   ->  tests/golden/rocs.pt
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_loader  # noqa: E402


def reference_functions():
    path = os.path.join(ref_loader.REF_ROOT, 'utils', 'evaluation', 'shaep_utils.py')
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ('VAB_2nd_order', 'shape_tanimoto', 'get_ROCS')]
    assert len(keep) == 3
    ns = {'torch': torch, 'np': np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, 'exec'), ns)
    return ns['get_ROCS']


def main():
    from oracle import shapemol_oracle as orc
    get_ROCS = reference_functions()
    g = torch.Generator().manual_seed(4)
    cases = []
    for n, r, scale in ((9, 22, 1.5), (27, 27, 2.0), (1, 5, 1.0), (20, 13, 0.5), (60, 60, 3.0), (14, 14, 0.0)):
        a = scale * torch.randn(n, 3, generator=g, dtype=torch.float64)
        b = a[:r].clone() + 0.3 * torch.randn(min(n, r), 3, generator=g, dtype=torch.float64) if n >= r else scale * torch.randn(r, 3, generator=g, dtype=torch.float64)
        if scale == 0.0:
            a = torch.randn(n, 3, generator=g, dtype=torch.float64)
            b = a.clone()                                  # identical sets: Tanimoto 1
        val = get_ROCS(a, b)
        got = orc.get_rocs(a, b)
        assert abs(float(val) - float(got)) < 1e-12, (float(val), float(got))
        cases.append(dict(a=a, b=b, rocs=val.clone()))
        print('n=%d r=%d rocs=%.6f' % (n, b.shape[0], float(val)))
    torch.save(cases, os.path.join(HERE, 'rocs.pt'))


if __name__ == '__main__':
    main()
