"""ctypes binding of libshapemol_b200.so (the C ABI declared in include/shapemol_b200.h).

There is no fallback: if the shared library is missing, load() raises.  Build it with
`python -m shapemol_b200.build` (or __graft_entry__.build()).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libshapemol_b200.so')

SMB_MAX_LAYERS = 16
SMB_MAX_ATOMS_PER_MOL = 64
SMB_MAX_K = 63
PREC_BF16X3 = 0
PREC_BF16 = 1
PRECISIONS = {'bf16x3': PREC_BF16X3, 'fp32': PREC_BF16X3, 'bf16': PREC_BF16}

_fp = C.c_void_p   # device / host pointers travel as plain integers


class ModelDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('hidden', 'heads', 'layers', 'k', 'classes', 'time_dim', 'timesteps', 'precision')]


class Batch(C.Structure):
    _fields_ = [('n_atoms', C.c_int32), ('n_mols', C.c_int32), ('max_atoms_per_mol', C.c_int32),
                ('mol_ptr', _fp), ('atom_mol', _fp)]


class ForwardIO(C.Structure):
    _fields_ = [('pos', _fp), ('v', _fp), ('shape', _fp), ('t', _fp),
                ('pred_pos', _fp), ('pred_h', _fp), ('pred_v', _fp), ('h0', _fp), ('nbr', _fp),
                ('bn_weight', _fp * SMB_MAX_LAYERS), ('bn_bias', _fp * SMB_MAX_LAYERS),
                ('bn_running_mean', _fp * SMB_MAX_LAYERS), ('bn_running_var', _fp * SMB_MAX_LAYERS),
                ('bn_num_batches_tracked', _fp * SMB_MAX_LAYERS), ('training', C.c_int32),
                ('prof_kernel', C.c_int32), ('prof_capacity', C.c_int32), ('prof_events', C.POINTER(_fp)),
                ('reuse_static', C.c_int32)]

PROF = {'edge_k': 1, 'edge_v': 2, 'edge_xv': 3, 'node_pre': 4, 'node_out': 5, 'gate': 6, 'knn': 7, 'head': 8}


class PosteriorIO(C.Structure):
    _fields_ = [('pred_pos', _fp), ('pred_v', _fp), ('t', _fp), ('pos', _fp), ('v', _fp),
                ('noise_pos', _fp), ('noise_u', _fp), ('log_v0', _fp), ('log_post', _fp),
                ('seed', C.c_uint64), ('atom_offset', C.c_int64),
                ('posterior_mean_c0_coef', _fp), ('posterior_mean_ct_coef', _fp), ('posterior_logvar', _fp),
                ('log_alphas_v', _fp), ('log_one_minus_alphas_v', _fp), ('log_alphas_cumprod_v', _fp),
                ('log_one_minus_alphas_cumprod_v', _fp)]


class GuidanceIO(C.Structure):
    _fields_ = [('pos', _fp), ('cloud', _fp), ('cloud_ptr', _fp), ('n_cloud', C.c_int32), ('t', _fp), ('grad_step', C.c_int32),
                ('step', C.c_int32), ('radius', C.c_double), ('ratio', C.c_double), ('u', _fp), ('seed', C.c_uint64),
                ('atom_offset', C.c_int64)]


class EncoderWeights(C.Structure):
    _fields_ = [('hidden', C.c_int32), ('latent', C.c_int32), ('n_blocks', C.c_int32), ('num_k', C.c_int32),
                ('conv_pos_feat', _fp), ('conv_pos_dir', _fp), ('conv_pos_bn_w', _fp), ('conv_pos_bn_b', _fp),
                ('conv_pos_bn_rm', _fp), ('conv_pos_bn_rv', _fp),
                ('block_feat', _fp * 8), ('block_dir', _fp * 8), ('block_bn_w', _fp * 8), ('block_bn_b', _fp * 8),
                ('block_bn_rm', _fp * 8), ('block_bn_rv', _fp * 8),
                ('conv_c_feat', _fp), ('conv_c_dir', _fp), ('conv_c_bn_w', _fp), ('conv_c_bn_b', _fp),
                ('conv_c_bn_rm', _fp), ('conv_c_bn_rv', _fp), ('training', C.c_int32)]


EXPORTS = {
    'smb_abi_version': (C.c_int, []),
    'smb_last_error_string': (C.c_char_p, []),
    'smb_param_count': (C.c_int, [C.POINTER(ModelDims)]),
    'smb_param_name': (C.c_char_p, [C.POINTER(ModelDims), C.c_int]),
    'smb_packed_weights_bytes': (C.c_size_t, [C.POINTER(ModelDims)]),
    'smb_pack_weights': (C.c_int, [C.POINTER(ModelDims), C.POINTER(_fp), C.c_int, _fp, C.c_size_t]),
    'smb_workspace_bytes': (C.c_size_t, [C.POINTER(ModelDims), C.c_int32, C.c_int32]),
    'smb_knn_graph': (C.c_int, [_fp, C.POINTER(Batch), C.c_int32, _fp, _fp, _fp]),
    'smb_forward': (C.c_int, [C.POINTER(ModelDims), _fp, C.POINTER(Batch), C.POINTER(ForwardIO), _fp, C.c_size_t, _fp]),
    'smb_type_head': (C.c_int, [C.POINTER(ModelDims), _fp, C.POINTER(Batch), _fp, _fp, _fp]),
    'smb_posterior_step': (C.c_int, [C.POINTER(ModelDims), C.POINTER(Batch), C.POINTER(PosteriorIO), _fp]),
    'smb_decrement_t': (C.c_int, [_fp, C.c_int32, _fp]),
    'smb_pointcloud_guidance': (C.c_int, [C.POINTER(Batch), C.POINTER(GuidanceIO), _fp]),
    'smb_check_stability': (C.c_int, [C.POINTER(Batch), _fp, _fp, _fp, _fp, C.c_int32, C.c_int32, _fp, _fp, _fp]),
    'smb_shape_tanimoto': (C.c_int, [C.POINTER(Batch), _fp, _fp, _fp, C.c_int32, C.c_double, C.c_double, C.c_double, _fp, _fp]),
    'smb_encoder_workspace_bytes': (C.c_size_t, [C.POINTER(EncoderWeights), C.c_int32, C.c_int32]),
    'smb_vn_dgcnn_encode': (C.c_int, [C.POINTER(EncoderWeights), _fp, C.c_int32, C.c_int32, _fp, _fp, C.c_size_t, _fp]),
}

_lib = None


class SmbError(RuntimeError):
    pass


def load():
    """Loads the shared library (once) and declares every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SmbError('%s is missing: the CUDA library has not been built (python -m shapemol_b200.build); '
                       'shapemol_b200 has no CPU fallback' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.smb_abi_version() != 2:
        raise SmbError('ABI version mismatch')
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().smb_last_error_string()
        raise SmbError('%s failed (code %d): %s' % (what, rc, (msg or b'').decode()))


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
