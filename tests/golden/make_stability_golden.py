"""Golden vectors for the stability check from the UNMODIFIED reference check_stability / get_bond_order
(utils/evaluation/analyze.py:249-297).  The module imports matplotlib at the top (absent here), so the definitions it needs are
compiled from the reference source file at generation time; nothing is copied into the repository.  Also asserts that
shapemol_b200/chem_tables.py equals the reference's dictionaries.

Run in the build container only:   python tests/golden/make_stability_golden.py   ->  tests/golden/stability.pt
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


def reference_namespace():
    path = os.path.join(ref_loader.REF_ROOT, 'utils', 'evaluation', 'analyze.py')
    tree = ast.parse(open(path).read())
    want_assign = {'atom_encoder', 'atom_decoder', 'bonds1', 'bonds2', 'bonds3', 'allowed_bonds'}
    keep = []
    for n in tree.body:
        if isinstance(n, ast.FunctionDef) and n.name in ('get_bond_order', 'check_stability'):
            keep.append(n)
        elif isinstance(n, ast.Assign):
            names = set()
            for t in n.targets:
                names |= {e.id for e in ast.walk(t) if isinstance(e, ast.Name)}
            if names & (want_assign | {'margin1', 'margin2', 'margin3'}):
                keep.append(n)
    ns = {'np': np, 'torch': torch}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, 'exec'), ns)
    return ns


def main():
    from oracle import shapemol_oracle as orc
    from shapemol_b200 import chem_tables as ct
    ns = reference_namespace()
    # ---- the restated tables equal the reference's ----
    b = ct.bond_tables()
    for k, name in enumerate(('bonds1', 'bonds2', 'bonds3')):
        for i, a in enumerate(ct.ELEMENTS):
            for j, c in enumerate(ct.ELEMENTS):
                assert int(b[k][i, j]) == ns[name][a][c], (name, a, c)
    assert ct.MARGINS == (ns['margin1'], ns['margin2'], ns['margin3'])
    assert all(ns['allowed_bonds'][e] == ct.ALLOWED_BONDS[i] and ns['atom_encoder'][e] == ct.ATOMIC_NUMBERS[i] for i, e in enumerate(ct.ELEMENTS))
    thr, allowed = ct.thresholds(), torch.tensor(ct.ALLOWED_BONDS, dtype=torch.int32)
    rng = np.random.RandomState(7)
    cases = []
    for n, scale, pool in ((12, 1.3, (6, 7, 8, 1)), (27, 1.6, (6, 6, 6, 7, 8, 9, 16, 17)), (1, 1.0, (6,)), (20, 1.1, (6, 7, 8, 15, 16, 35, 53)),
                           (40, 2.0, (1, 6, 7, 8)), (9, 0.9, (6, 8))):
        pos = (rng.randn(n, 3) * scale).astype(np.float32)
        z = np.array([pool[i] for i in rng.randint(0, len(pool), n)])
        for hs in (False, True):
            stable, nr_stable, total, nr_bonds = ns['check_stability'](pos, z, hs=hs, return_nr_bonds=True)
            got = orc.check_stability(torch.from_numpy(pos), ct.element_index(torch.from_numpy(z)), thr, allowed, hs=hs)
            assert (bool(stable), int(nr_stable), int(total)) == got[:3] and np.array_equal(nr_bonds, got[3].numpy()), (n, hs)
            cases.append(dict(pos=torch.from_numpy(pos), z=torch.from_numpy(z), hs=hs, stable=bool(stable), nr_stable=int(nr_stable),
                              nr_bonds=torch.from_numpy(np.asarray(nr_bonds)).long()))
        print('n=%d: stable atoms %d / %d, bonds %s' % (n, nr_stable, total, nr_bonds.tolist()[:10]))
    torch.save(cases, os.path.join(HERE, 'stability.pt'))


if __name__ == '__main__':
    main()
