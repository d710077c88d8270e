"""ctypes binding of libcatb200.so (the C ABI declared in include/catb200.h).

The library is built in-tree (cat_b200/lib/libcatb200.so) by ``__graft_entry__.build()`` or
``make -C cat_b200/csrc``.  There is no CPU or PyTorch fallback: if the library is missing, or no sm_100
device is present when a kernel is requested, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libcatb200.so')

ACT_NONE, ACT_RELU, ACT_LEAKY02, ACT_TANH, ACT_LEAKY001 = 0, 1, 2, 3, 4
PAD_ZERO, PAD_REFLECT = 0, 1
GAN_MODES = {'hinge': 0, 'lsgan': 1, 'vanilla': 2}
RECON_KINDS = {'l1': 0, 'l2': 1, 'smooth_l1': 2}


class GatherUnit(C.Structure):
    _fields_ = [('dr', C.c_int8), ('ds', C.c_int8), ('cu', C.c_int16)]


class WeightUnit(C.Structure):
    _fields_ = [('w_off', C.c_int32), ('sn_w', C.c_int32), ('sc_w', C.c_int32), ('nvalid', C.c_int32)]


class IgemmDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'N', 'H', 'W', 'ldx', 'x_coff', 'OH', 'OW', 'ldy', 'y_coff', 'o_step', 'o_ph', 'o_pw', 'OHs', 'OWs',
        'sn', 'sd', 'pad_mode', 'n_units', 'n_rows', 'n_tile', 'act', 'accumulate', 'y_is_f32', 'reserved')]


class HaloDesc(C.Structure):
    _fields_ = [('n_steps', C.c_int32), ('n_chunks', C.c_int32), ('n_planes', C.c_int32),
                ('plane_pa', C.c_int32 * 4), ('plane_pb', C.c_int32 * 4), ('plane_y0', C.c_int32 * 4),
                ('plane_x0', C.c_int32 * 4), ('mul', C.c_int32), ('TW', C.c_int32), ('n_strips', C.c_int32),
                ('Wf', C.c_int32), ('Lh', C.c_int32),
                ('Ymax', C.c_int32), ('Xmax', C.c_int32), ('m_sub', C.c_int32), ('b_budget', C.c_int32)]


class EpilogueStats(C.Structure):
    _fields_ = [('sums', C.c_void_p), ('C', C.c_int32), ('coff', C.c_int32), ('per_sample', C.c_int32),
                ('reserved', C.c_int32)]


class PackJob(C.Structure):
    _fields_ = [('wunits', C.c_void_p), ('packed', C.c_void_p), ('n_tile', C.c_int32), ('n_units', C.c_int32),
                ('n_chunks', C.c_int32), ('row0', C.c_int32), ('span', C.c_int32), ('nreal', C.c_int32)]


class UnpackJob(C.Structure):
    _fields_ = [('ws', C.c_void_p), ('wunits', C.c_void_p), ('grad', C.c_void_p), ('n_splits', C.c_int32),
                ('ws_rows', C.c_int32), ('ws_k', C.c_int32), ('row0', C.c_int32), ('n_rows', C.c_int32), ('n_units', C.c_int32)]


class CatbError(RuntimeError):
    pass


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float
_DP = C.POINTER(IgemmDesc)

# name -> argtypes (all return int unless listed in _SPECIAL)
_PROTOS = {
    'catb_init': [_I],
    'catb_debug_timeline': [_P],
    'catb_debug_mode': [_I],
    'catb_pack_weights': [_DP, _P, _P, _P, _P],
    'catb_pack_weights_rows': [_DP, _P, _P, _P, _I, _I, _I, _P],
    'catb_pack_weights_batch': [_P, _I, _I, _P, _P],
    'catb_igemm_fprop': [_DP, _P, _P, _P, _P, _P, _P],
    'catb_igemm_wgrad': [_DP, _P, _P, _P, _P, _P, _P],
    'catb_igemm_halo_fprop': [_DP, C.POINTER(HaloDesc), _P, _P, _P, _P, _P, _P, _P, _P],
    'catb_igemm_halo_fprop_persist': [_DP, C.POINTER(HaloDesc), _P, _P, _P, _P, _P, _P, _I, _I, _P, _P],
    'catb_igemm_halo_wgrad': [_DP, C.POINTER(HaloDesc), _P, _P, _P, _I, _P, _P, _P, _P, _P],
    'catb_igemm_wgrad_ws_shape': [_DP, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    'catb_igemm_wgrad_ws': [_DP, _P, _P, _P, _P, _P],
    'catb_igemm_halo_wgrad_ws_shape': [_DP, C.POINTER(HaloDesc), _I, _I, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    'catb_igemm_halo_wgrad_ws': [_DP, C.POINTER(HaloDesc), _P, _P, _P, _I, _P, _P, _P, _I, _I, _P],
    'catb_wgrad_unpack': [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    'catb_wgrad_unpack_batch': [_P, _I, _I, _P],
    'catb_ref_fprop': [_DP, _P, _P, _P, _P, _P, _P, _P],
    'catb_ref_wgrad': [_DP, _P, _P, _P, _P, _P, _P],
    'catb_dwconv_fwd': [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P],
    'catb_dwconv_bwd_data': [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P],
    'catb_dwconv_bwd_weight': [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P],
    'catb_norm_stats': [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    'catb_norm_finalize': [_P, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P],
    'catb_norm_apply': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P],
    'catb_norm_apply_fused': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P,
                              _I, _P],
    'catb_norm_bwd_reduce': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P],
    'catb_norm_bwd_apply': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _F, _I,
                            _P, _P, _P],
    'catb_nchw_to_nhwc': [_P, _I, _I, _I, _I, _P, _I, _I, _P],
    'catb_nhwc_to_nchw': [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    'catb_copy_channels': [_P, _I, _I, _P, _I, _I, _L, _I, _P],
    'catb_act_bwd': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _L, _I, _I, _P],
    'catb_channel_sum': [_P, _I, _I, _L, _I, _P, _P],
    'catb_reflect_fold': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'catb_add': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _L, _I, _P],
    'catb_gan_loss': [_P, _L, _I, _I, _I, _I, _F, _P, _P, _I, _I, _P],
    'catb_recon_loss': [_P, _I, _I, _P, _I, _I, _L, _I, _I, _I, _F, _P, _P, _I, _I, _P, _I, _I, _P],
    'catb_gram': [_P, _I, _I, _I, _L, _I, _P, _P],
    'catb_gram_ref': [_P, _I, _I, _I, _L, _I, _P, _P],
    'catb_ka_finish': [_P, _P, _I, _F, _P, _P, _P, _P],
    'catb_ka_bwd': [_P, _I, _I, _I, _L, _I, _P, _P, _I, _I, _I, _P],
    'catb_adam': [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _P, _P],
    # SPADE path
    'catb_resize_nearest': [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    'catb_upsample2x_bwd': [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    'catb_spade_modulate': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _L, _I, _P, _P, _I, _P],
    'catb_spade_modulate_bwd': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I,
                                _L, _I, _P, _P, _I, _P],
    'catb_act_fwd': [_P, _I, _I, _P, _I, _I, _L, _I, _I, _P],
    'catb_avgpool3s2': [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P],
    'catb_avgpool3s2_bwd': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    'catb_maxpool2': [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P],
    'catb_maxpool2_bwd': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    'catb_onehot_edges': [_P, _P, _I, _I, _I, _I, _P, _I, _I, _I, _P],
    'catb_gather_sum_f32': [_P, _P, _I, _I, _P, _P],
    'catb_scatter_add_f32': [_P, _P, _I, _I, _P, _P],
    'catb_fma_vec': [_P, _P, _P, _I, _P],
    'catb_sn_forward': [_P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P],
    'catb_sn_backward': [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    'catb_expand_x': [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'catb_shift_sum': [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P],
    'catb_shift_expand': [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'catb_tap_sum': [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    'catb_tap_expand': [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
}
_SPECIAL = {
    'catb_version': ([], C.c_char_p),
    'catb_last_error_string': ([], C.c_char_p),
    'catb_packed_weight_bytes': ([_I, _I, _I], C.c_size_t),
    'catb_igemm_halo_fits': ([_I, _I, _I, _I, _I, _I], C.c_int),
    'catb_igemm_halo_wgrad_fits': ([_I, _I], C.c_int),
    'catb_igemm_halo_persist_fits': ([_I] * 10, C.c_int),
    'catb_igemm_halo_wgrad_tma_fits': ([C.POINTER(HaloDesc)], C.c_int),
}
EXPORTED_SYMBOLS = sorted(list(_PROTOS) + list(_SPECIAL))

_lib = None
_inited_devices = set()


def load():
    """Load libcatb200.so and declare its prototypes (no GPU needed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CatbError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                        f'or `make -C cat_b200/csrc` -- cat_b200 has no fallback path')
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name, (argtypes, restype) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def init(device=0):
    lib = load()
    if device not in _inited_devices:
        check(lib.catb_init(int(device)), 'catb_init')
        _inited_devices.add(device)
    return lib


def check(status, what=''):
    if status != 0:
        msg = load().catb_last_error_string().decode()
        raise CatbError(f'{what} failed with status {status}: {msg}')


LAUNCH_COUNT = [0]  # kernels enqueued so far (one per entry-point call; catb_adam 2, catb_sn_forward 5, catb_sn_backward 2)


def call(name, *args):
    """Invoke an int-returning entry point and raise on a non-zero status."""
    LAUNCH_COUNT[0] += {'catb_adam': 2, 'catb_sn_forward': 5, 'catb_sn_backward': 2}.get(name, 1)
    check(getattr(load(), name)(*args), name)
