"""ctypes binding of libk4nerf.so -- the exact declarations of include/k4nerf.h."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libk4nerf.so')

K4_OK = 0
K4_KIND_DVGO, K4_KIND_DMPIGO, K4_KIND_DCVGO = 0, 1, 2
K4_MLP_FP32, K4_MLP_F16, K4_MLP_F16X3, K4_MLP_TCGEN05, K4_MLP_TCGEN05_WS = 0, 1, 2, 3, 4
K4_MAX_MLP_LAYERS = 8
K4_MAX_PEERS = 8
MLP_MODES = {'fp32': K4_MLP_FP32, 'f16': K4_MLP_F16, 'f16x3': K4_MLP_F16X3, 'tcgen05': K4_MLP_TCGEN05,
             'tc': K4_MLP_TCGEN05, 'ws': K4_MLP_TCGEN05_WS}


class SceneDesc(C.Structure):
    _fields_ = [
        ('kind', C.c_int32), ('world_size', C.c_int32 * 3), ('k0_dim', C.c_int32), ('mask_size', C.c_int32 * 3),
        ('xyz_min', C.c_float * 3), ('xyz_max', C.c_float * 3),
        ('xyz2ijk_scale', C.c_float * 3), ('xyz2ijk_shift', C.c_float * 3),
        ('act_shift', C.c_float), ('voxel_size', C.c_float), ('voxel_size_ratio', C.c_float),
        ('fast_color_thres', C.c_float),
        ('max_world_size', C.c_int32), ('mpi_depth', C.c_int32),
        ('rgbnet_depth', C.c_int32), ('rgbnet_width', C.c_int32), ('rgbnet_direct', C.c_int32),
        ('viewbase_pe', C.c_int32), ('spatial_pe', C.c_int32), ('reserved0', C.c_int32),
        ('d_density', C.c_void_p), ('d_k0', C.c_void_p), ('d_mask', C.c_void_p), ('d_act_shift_grid', C.c_void_p),
        ('d_rgbnet_weight', C.c_void_p * K4_MAX_MLP_LAYERS), ('d_rgbnet_bias', C.c_void_p * K4_MAX_MLP_LAYERS),
        ('scene_center', C.c_float * 3), ('scene_radius', C.c_float * 3), ('bg_len', C.c_float), ('world_len', C.c_int32),
    ]


class RenderArgs(C.Structure):
    _fields_ = [
        ('near_', C.c_float), ('far_', C.c_float), ('stepsize', C.c_float), ('bg', C.c_float),
        ('render_depth', C.c_int32), ('mlp_mode', C.c_int32), ('image_w', C.c_int32), ('image_h', C.c_int32),
        ('d_t_list', C.c_void_p), ('n_t', C.c_int32), ('dist_thres', C.c_float),
    ]


class RenderOut(C.Structure):
    _fields_ = [
        ('d_rgb_marched', C.c_void_p), ('d_depth', C.c_void_p), ('d_alphainv_last', C.c_void_p),
        ('d_ray_stats', C.c_void_p), ('d_t_minmax', C.c_void_p), ('d_counters', C.c_void_p),
    ]


class FrameDst(C.Structure):
    _fields_ = [
        ('n_dst', C.c_int32), ('rank', C.c_int32), ('world', C.c_int32), ('frame_w', C.c_int32),
        ('n_full', C.c_int64), ('d_frame', C.c_void_p * K4_MAX_PEERS),
    ]


class SrnetDesc(C.Structure):
    _fields_ = [
        ('n_in_colors', C.c_int32), ('scale', C.c_int32), ('num_feat', C.c_int32), ('num_block', C.c_int32),
        ('num_grow_ch', C.c_int32), ('num_cond', C.c_int32), ('n_params', C.c_int32), ('reserved0', C.c_int32),
        ('h_params', C.POINTER(C.c_void_p)),
    ]


EXPORTS = {
    'k4_abi_version': (C.c_int, []),
    'k4_status_string': (C.c_char_p, [C.c_int]),
    'k4_last_cuda_error': (C.c_char_p, []),
    'k4_device_check': (C.c_int, []),
    'k4_scene_create': (C.c_int, [C.POINTER(SceneDesc), C.c_void_p, C.POINTER(C.c_void_p)]),
    'k4_scene_destroy': (C.c_int, [C.c_void_p]),
    'k4_scene_device_bytes': (C.c_size_t, [C.c_void_p]),
    'k4_scene_best_mlp_mode': (C.c_int, [C.c_void_p]),
    'k4_scene_ws_config': (C.c_int, [C.c_void_p]),
    'k4_render_workspace_bytes': (C.c_size_t, [C.c_void_p, C.c_int64]),
    'k4_render_rays': (C.c_int, [C.c_void_p, C.POINTER(RenderArgs), C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int64, C.POINTER(RenderOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    'k4_render_rays_frames': (C.c_int, [C.c_void_p, C.POINTER(RenderArgs), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.POINTER(FrameDst), C.POINTER(RenderOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    'k4_peer_enable_all': (C.c_int, []),
    'k4_peer_alloc': (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    'k4_peer_free': (C.c_int, [C.c_void_p]),
    'k4_peer_export': (C.c_int, [C.c_void_p, C.c_char_p]),
    'k4_peer_open': (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    'k4_peer_close': (C.c_int, [C.c_void_p]),
    'k4_srnet_create': (C.c_int, [C.POINTER(SrnetDesc), C.c_void_p, C.POINTER(C.c_void_p)]),
    'k4_srnet_destroy': (C.c_int, [C.c_void_p]),
    'k4_srnet_workspace_bytes': (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32]),
    'k4_srnet_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.c_size_t, C.c_void_p]),
    'k4_srnet_forward_roi': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    'k4_srnet_forward_roi_peers': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_void_p),
                                             C.c_void_p, C.c_size_t, C.c_void_p]),
    'k4_op_infer_t_minmax': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'k4_op_infer_n_samples': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_void_p, C.c_void_p]),
    'k4_op_infer_ray_start_dir': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'k4_op_fill_ray_step_ids': (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'k4_op_sample_pts': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'k4_op_sample_ndc_pts': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'k4_op_sample_bg_pts': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    'k4_op_maskcache_lookup': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    'k4_op_raw2alpha': (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'k4_op_raw2alpha_backward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'k4_op_alpha2weight': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'k4_op_alpha2weight_backward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'k4_op_total_variation_add_grad': (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    'k4_op_adam_upd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_void_p]),
    'k4_op_grid_alpha': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    'k4_op_maxpool3_thres_and': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    'k4_op_cumdist_thres': (C.c_int, [C.c_void_p, C.c_float, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    'k4_op_resample_trilinear': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    'k4_op_grid_sample': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                    C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'k4_op_grid_sample_backward': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                             C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'k4_make_rays_rows': (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'k4_make_rays': (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

if not os.path.exists(LIB_PATH):
    raise RuntimeError(
        f'{LIB_PATH} is missing: build it with `python 4k-nerf_b200/csrc/build.py` '
        '(k4nerf has no CPU or PyTorch fallback for the rendering hot path)')

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in EXPORTS.items():
    _fn = getattr(lib, _name)           # AttributeError here == the library does not match the header
    _fn.restype = _res
    _fn.argtypes = _args

if lib.k4_abi_version() != 1:
    raise RuntimeError('libk4nerf.so ABI version mismatch')


class K4Error(RuntimeError):
    pass


def check(status, what=''):
    if status != K4_OK:
        msg = lib.k4_status_string(status).decode()
        if status == -3:
            msg += ': ' + lib.k4_last_cuda_error().decode()
        raise K4Error(f'{what}: {msg}' if what else msg)
