"""Golden vectors for the knowledge-gradient function from the REFERENCE's own
``util.misc_util.knowledge_gradient`` (/root/reference/src, imported where it lies; build
container only).  ``np.float`` (removed from numpy) is restored for the import, nothing else:
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_kg_golden.py
"""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402

from oracle import refshim  # noqa: E402


def cases():
    rs = np.random.RandomState(7)
    out = []
    for n in (2, 3, 5, 17, 100, 250):
        for scale in (1.0, 1e-3):
            a = rs.normal(size=n)
            b = scale * rs.normal(size=n)
            out.append((a, b))
    # lines that never reach the envelope, and a near-degenerate fan
    out.append((np.asarray([0.0, -5.0, -5.0, 0.1]), np.asarray([-1.0, -0.5, 0.5, 1.0])))
    out.append((np.linspace(0, 1, 40) ** 2, np.linspace(-1, 1, 40)))
    return out


def main():
    refshim.install()
    if not hasattr(np, 'float'):
        np.float = float  # numpy >= 1.24 removed the alias the reference uses at :192
    from util.misc_util import knowledge_gradient
    vecs = []
    for a, b in cases():
        vecs.append({'a': [float(v) for v in a], 'b': [float(v) for v in b],
                     'kg': float(knowledge_gradient(np.array(a), np.array(b)))})
    path = os.path.join(ROOT, 'tests', 'golden', 'kg_vectors.json')
    with open(path, 'w') as fh:
        json.dump(vecs, fh)
    print('wrote', path, len(vecs), 'vectors')


if __name__ == '__main__':
    main()
