"""Drop-in `models` package.  Modules not overridden here (shape_modelAE, ...) fall through to the
reference checkout named by $SHAPEMOL_REFERENCE_ROOT, if any."""
import os

_ref = os.environ.get('SHAPEMOL_REFERENCE_ROOT')
if _ref and os.path.isdir(os.path.join(_ref, 'models')):
    __path__.append(os.path.join(_ref, 'models'))
