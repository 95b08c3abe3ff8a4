"""Put this directory on sys.path BEFORE the reference root to make `import models.molopt_score_model`
resolve to the B200-native drop-in."""
import os
import sys


def install(reference_root=None):
    """Prepends the drop-in `models` package to sys.path (and remembers where the reference lives so
    that modules we do not override, e.g. models.shape_modelAE, still resolve)."""
    here = os.path.dirname(os.path.abspath(__file__))
    if reference_root:
        os.environ['SHAPEMOL_REFERENCE_ROOT'] = reference_root
    for name in [m for m in sys.modules if m == 'models' or m.startswith('models.')]:
        del sys.modules[name]
    if here in sys.path:
        sys.path.remove(here)
    sys.path.insert(0, here)
