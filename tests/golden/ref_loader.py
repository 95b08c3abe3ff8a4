"""Imports the UNMODIFIED reference modules from /root/reference through the shims in
tests/golden/_shims.  Only usable in the build container (the GPU box has no /root/reference);
used by make_golden.py to create the committed fixtures, by optional CPU tests that are
skipped when the reference tree is absent, and by bench.py's CPU arm (which finds the staged copy
oracle/_ref on the GPU box)."""
import os
import sys
import types

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_shims')
# the reference checkout (build container) or the files oracle/stage_ref.py staged for the GPU box (oracle/_ref, git-ignored)
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'oracle', '_ref')
REF_ROOT = os.environ.get('SHAPEMOL_REFERENCE', '/root/reference')
if not os.path.isdir(os.path.join(REF_ROOT, 'models')) and os.path.isdir(os.path.join(_STAGED, 'models')):
    REF_ROOT = os.path.abspath(_STAGED)


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'models'))


def load():
    """Returns (molopt_score_model, shape_pointcloud_modelAE, EasyDict) reference modules."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    sys.dont_write_bytecode = True
    for p in (REF_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, _SHIMS)
    # a previously imported drop-in `models` package must not shadow the reference
    for name in [m for m in sys.modules if m == 'models' or m.startswith('models.')]:
        del sys.modules[name]
    if 'utils.covalent_graph' not in sys.modules:
        stub = types.ModuleType('utils.covalent_graph')

        def connect_covalent_graph(*a, **k):
            raise NotImplementedError('cov_radius graph is not on the hot path')
        stub.connect_covalent_graph = connect_covalent_graph
        sys.modules['utils.covalent_graph'] = stub
    import models.molopt_score_model as msm
    import models.shape_pointcloud_modelAE as spm
    from easydict import EasyDict
    return msm, spm, EasyDict


def model_config(EasyDict, **overrides):
    import yaml
    path = os.path.join(REF_ROOT, 'config/training',
                        'dgcnn_signeddist_512_attention_residue_uniform_pos0_10_pos1.e-7_0.01_6_v001.yml')
    with open(path) as f:
        cfg = EasyDict(yaml.safe_load(f))
    for k, v in overrides.items():
        cfg.model[k] = v
    return cfg.model
