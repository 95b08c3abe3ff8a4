"""Stages the UNMODIFIED reference files the CPU arm needs into oracle/_ref/ (git-ignored, travels to the GPU box with the
snapshot like the built .so does).  Test / benchmark infrastructure only: nothing under shapemol_b200/ imports it.

Called by __graft_entry__.build() in the build container (where /root/reference exists); a no-op elsewhere.
The files are copied byte for byte -- the reference is pure Python and has nothing to compile -- so `bench.py --impl
reference` and `cpu_baseline` time the reference's own code (kind "reference"), not the oracle port.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get('SHAPEMOL_REFERENCE', '/root/reference')
DST = os.path.join(HERE, '_ref')

# the six hot-path modules, the training YAML the constructor reads, the shipped shape-encoder checkpoint
FILES = [
    'models/__init__.py',            # (empty) makes `models` a regular package, like in the checkout
    'models/molopt_score_model.py', 'models/uni_transformer.py', 'models/common.py', 'models/diffusion.py',
    'models/shape_vn_layers.py', 'models/shape_pointcloud_modelAE.py',
    'config/training/dgcnn_signeddist_512_attention_residue_uniform_pos0_10_pos1.e-7_0.01_6_v001.yml',
    'trained_models/se_model.pt',
    # the caller of the hot path: tests/test_gpu_script_contract.py executes its sample_diffusion_ligand() unmodified against the drop-in
    'scripts/sample_diffusion.py',
]


def stage(verbose=True):
    if not os.path.isdir(os.path.join(REF_SRC, 'models')):
        return None
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
    # `from utils import *` in models/shape_vn_layers.py:5 resolves against a namespace package: an empty directory is enough
    os.makedirs(os.path.join(DST, 'utils'), exist_ok=True)
    with open(os.path.join(DST, 'utils', '.keep'), 'w') as f:
        f.write('')
    if verbose:
        print('staged reference files into', DST)
    return DST


if __name__ == '__main__':
    stage()
