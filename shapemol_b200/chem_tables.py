"""Element tables of the stability check (reference utils/evaluation/analyze.py:11-54): typical single / double / triple bond
lengths in picometres (source quoted there: wiredchemist.com bond energies & lengths), the tuned margins and the allowed
valences.  Stored as the upper triangle of symmetric matrices over ELEMENTS; -1 = no such bond.
tests/golden/make_stability_golden.py asserts that these equal the reference's dictionaries."""
import torch

ELEMENTS = ('H', 'C', 'N', 'O', 'F', 'P', 'S', 'Cl', 'Br', 'I')
ATOMIC_NUMBERS = (1, 6, 7, 8, 9, 15, 16, 17, 35, 53)
ALLOWED_BONDS = (1, 4, 3, 2, 1, 5, 4, 1, 1, 1)
MARGINS = (10, 5, 3)

_SINGLE = """
 74 109 101  96  92 144 134 127 141 161
    154 147 143 135 184 182 177 194 214
        145 140 136 177 168 175  -1  -1
            148 142 163 151 164  -1  -1
                142 156 158 166  -1 191
                    221 210 203  -1  -1
                        204 207  -1  -1
                            199  -1 232
                                228  -1
                                    267
"""
_DOUBLE = {('C', 'C'): 134, ('C', 'N'): 129, ('C', 'O'): 120, ('C', 'S'): 160, ('N', 'N'): 125, ('N', 'O'): 121, ('O', 'O'): 121,
           ('O', 'P'): 150, ('P', 'S'): 186}
_TRIPLE = {('C', 'C'): 120, ('C', 'N'): 116, ('C', 'O'): 113, ('N', 'N'): 110}


def _sym_from_triangle(text):
    n = len(ELEMENTS)
    m = [[-1] * n for _ in range(n)]
    rows = [r.split() for r in text.strip().splitlines()]
    for i, r in enumerate(rows):
        for k, val in enumerate(r):
            m[i][i + k] = m[i + k][i] = int(val)
    return m


def _sym_from_pairs(pairs):
    n = len(ELEMENTS)
    m = [[-1] * n for _ in range(n)]
    for (a, b), val in pairs.items():
        i, j = ELEMENTS.index(a), ELEMENTS.index(b)
        m[i][j] = m[j][i] = val
    return m


def bond_tables():
    """(bonds1, bonds2, bonds3) as int32 [10,10] tensors in pm, -1 where absent."""
    return tuple(torch.tensor(t, dtype=torch.int32) for t in (_sym_from_triangle(_SINGLE), _sym_from_pairs(_DOUBLE), _sym_from_pairs(_TRIPLE)))


def thresholds():
    """int32 [3,10,10]: bonds_k + margin_k -- a pair at distance d (in pm) has order >= k iff d < thr[k-1] for all k' <= k
    (get_bond_order, analyze.py:249-261)."""
    return torch.stack([t + m for t, m in zip(bond_tables(), MARGINS)])


def element_index(atomic_numbers):
    """atomic numbers [N] (int tensor) -> element index 0..9 (raises on an element outside the table, like the reference's KeyError)."""
    lut = torch.full((64,), -1, dtype=torch.int32)
    for i, z in enumerate(ATOMIC_NUMBERS):
        lut[z] = i
    z = atomic_numbers.to(torch.long).cpu()
    if bool(((z < 0) | (z >= 64)).any()) or bool((lut[z] < 0).any()):
        raise KeyError('atomic number outside the stability tables: %s' % sorted(set(z.tolist()) - set(ATOMIC_NUMBERS)))
    return lut[z].to(atomic_numbers.device)
