/*
 * k4nerf.h -- C ABI of libk4nerf.so, the B200 (sm_100a) implementation of the 4K-NeRF rendering
 * hot path.  This header is the drop-in boundary: plain pointers and sizes,
 * no torch types.  The reference has no C ABI of its own; every entry point below names the
 * reference interface it replaces (file:line under /root/reference).  The Python host mirror of
 * the reference API (k4nerf.dvgo.DirectVoxGO etc.) is a thin ctypes caller of these functions;
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *  - every function returns K4_OK (0) or a negative k4_status; nothing throws, nothing exits;
 *  - all pointers named d_* / inside descriptors are DEVICE pointers on the current device unless
 *    the name says host (h_*); fp32 unless stated;
 *  - the caller allocates and owns all outputs and the workspace; no hidden allocation, no hidden
 *    host synchronisation on the render calls (the reference syncs twice per call:
 *    lib/cuda/render_utils_kernel.cu:212 and :635);
 *  - all work is enqueued on the given stream (pass torch.cuda.current_stream().cuda_stream);
 *  - re-entrant across streams; a k4_scene is immutable after creation.
 */
#ifndef K4NERF_H_
#define K4NERF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define K4_ABI_VERSION 1

#if defined(__GNUC__)
#define K4_API __attribute__((visibility("default")))
#else
#define K4_API
#endif

typedef enum k4_status {
    K4_OK = 0,
    K4_ERR_INVALID_ARG = -1,   /* NULL pointer, bad shape, bad enum                     */
    K4_ERR_UNSUPPORTED = -2,   /* valid in the reference but not built here (see DESIGN) */
    K4_ERR_CUDA = -3,          /* a CUDA runtime call failed; see k4_last_cuda_error()   */
    K4_ERR_WORKSPACE = -4,     /* workspace NULL or smaller than k4_render_workspace_bytes */
    K4_ERR_NO_DEVICE = -5      /* no sm_100 device / library built for another arch      */
} k4_status;

typedef void* k4_stream_t;                 /* cudaStream_t */
typedef struct k4_scene k4_scene;          /* opaque: device-resident, repacked scene     */
typedef struct k4_srnet k4_srnet;          /* opaque: device-resident VC-Decoder weights  */

enum { K4_KIND_DVGO = 0,                   /* DirectVoxGO,            lib/dvgo.py:23-448   */
       K4_KIND_DMPIGO = 1,                 /* DirectMPIGO,            lib/dmpigo.py:18-427 */
       K4_KIND_DCVGO = 2 };                /* DirectContractedVoxGO,  lib/dcvgo.py:27-382  */

enum { K4_MLP_FP32 = 0,                    /* fp32 FFMA, sequential accumulation (exact mode)   */
       K4_MLP_F16 = 1,                     /* tensor cores, fp16 operands, fp32 accumulate       */
       K4_MLP_F16X3 = 2,                   /* tensor cores, error-compensated 3-term fp16 split  */
       K4_MLP_TCGEN05 = 3,                 /* tcgen05 + TMEM, fp16 operands (CTA-wide batches)   */
       K4_MLP_TCGEN05_WS = 4 };            /* same, MLP in its own warpgroup (warp-specialised)  */

#define K4_MAX_MLP_LAYERS 8
#define K4_MAX_PEERS 8            /* GPUs of one NVLink / NVSwitch node */

/*
 * Scene description = the tensors of a reference checkpoint's `model_state_dict` plus the derived
 * scalars its constructor computes (SURVEY.md section 5 "Checkpoint"):
 *   density.grid [1,1,X,Y,Z], k0.grid [1,C,X,Y,Z] (planar, z fastest: lib/grid.py:115),
 *   mask_cache.mask [mX,mY,mZ] bool + xyz2ijk_scale/shift (lib/grid.py:290-293),
 *   act_shift (DVGO scalar lib/dvgo.py:46 | MPI grid [1,1,1,1,D] lib/dmpigo.py:48-58),
 *   rgbnet.{0,1.0,...}.{weight,bias} (lib/dvgo.py:116-124; weight [out,in] row major).
 * k4_scene_create copies/repacks everything it needs; the caller may free its tensors afterwards.
 */
typedef struct k4_scene_desc {
    int32_t kind;                 /* K4_KIND_*                                                  */
    int32_t world_size[3];        /* X, Y, Z                                                    */
    int32_t k0_dim;               /* C (3 when rgbnet_depth == 0)                               */
    int32_t mask_size[3];         /* mX, mY, mZ                                                 */
    float xyz_min[3];
    float xyz_max[3];
    float xyz2ijk_scale[3];
    float xyz2ijk_shift[3];
    float act_shift;              /* DVGO: density bias (lib/dvgo.py:46); MPI: ignored          */
    float voxel_size;             /* DVGO: lib/dvgo.py:155; MPI: ignored                        */
    float voxel_size_ratio;       /* lib/dvgo.py:158 | lib/dmpigo.py:164                        */
    float fast_color_thres;       /* lib/dvgo.py:38                                             */
    int32_t max_world_size;       /* DVGO: world_size.max() (depth normalisation, dvgo.py:311)  */
    int32_t mpi_depth;            /* MPI: number of planes (lib/dmpigo.py:159)                  */
    int32_t rgbnet_depth;         /* number of Linear layers; 0 = no MLP, rgb = sigmoid(k0)     */
    int32_t rgbnet_width;
    int32_t rgbnet_direct;        /* DVGO only (lib/dvgo.py:382-386,409-412)                    */
    int32_t viewbase_pe;          /* number of view-direction frequencies                       */
    int32_t spatial_pe;           /* MPI only: number of positional frequencies                 */
    int32_t reserved0;
    const float* d_density;       /* [X*Y*Z]                                                    */
    const float* d_k0;            /* [C*X*Y*Z] planar, reference layout                         */
    const uint8_t* d_mask;        /* [mX*mY*mZ] bytes, non-zero = occupied                      */
    const float* d_act_shift_grid;/* MPI: [mpi_depth]; DVGO: NULL                               */
    const float* d_rgbnet_weight[K4_MAX_MLP_LAYERS]; /* [out,in] row major                      */
    const float* d_rgbnet_bias[K4_MAX_MLP_LAYERS];   /* [out]                                   */
    /* DirectContractedVoxGO only (lib/dcvgo.py:46-55,226-248): xyz_min/max above are -+(1+bg_len)  */
    float scene_center[3];        /* (xyz_min + xyz_max) / 2 of the foreground cube              */
    float scene_radius[3];        /* (xyz_max - xyz_min) / 2                                     */
    float bg_len;                 /* thickness of the contracted background shell                */
    int32_t world_len;            /* world_size[0]                                               */
} k4_scene_desc;

/* render_kwargs of the reference forward (run_sr.py:1311-1320; lib/dvgo.py:327) + kernel knobs. */
typedef struct k4_render_args {
    float near_;                  /* render_kwargs['near']                                      */
    float far_;                   /* render_kwargs['far'] (DVGO forces 1e9, lib/dvgo.py:307)    */
    float stepsize;               /* render_kwargs['stepsize']                                  */
    float bg;                     /* render_kwargs['bg']                                        */
    int32_t render_depth;         /* render_kwargs['render_depth']                              */
    int32_t mlp_mode;             /* K4_MLP_*                                                   */
    int32_t image_w;              /* >0: rays are a row-major HxW image -> 8x4 pixel tiles/warp */
    int32_t image_h;
    /* DirectContractedVoxGO only: the per-step ray parameters `t` of sample_ray (lib/dcvgo.py:241-248,
     * a function of world_len, bg_len and stepsize only; the host computes them with the reference's
     * own torch expressions) and the cumdist_thres threshold (lib/dcvgo.py:283) */
    const float* d_t_list;
    int32_t n_t;
    float dist_thres;
} k4_render_args;

typedef struct k4_render_out {
    float* d_rgb_marched;         /* [N,3]  == rgb_feature (aliased in the reference, dvgo.py:425-427) */
    float* d_depth;               /* [N] or NULL                                                */
    float* d_alphainv_last;       /* [N]                                                        */
    int32_t* d_ray_stats;         /* optional [N,4]: N_steps, S_m, S_d, S_c per ray (parity)    */
    float* d_t_minmax;            /* optional [N,2]: t_min, t_max per ray (parity)              */
    unsigned long long* d_counters; /* optional [4]: total S_m, S_d, S_c, MLP batches           */
} k4_render_out;

K4_API int k4_abi_version(void);
K4_API const char* k4_status_string(int status);
K4_API const char* k4_last_cuda_error(void);       /* thread-local text of the last CUDA failure */
K4_API int k4_device_check(void);                  /* K4_OK iff the current device is sm_100     */

/* Replaces: utils.load_model -> model.to(device) (lib/utils.py:62-66) for the marcher's tensors. */
K4_API int k4_scene_create(const k4_scene_desc* desc, k4_stream_t stream, k4_scene** out_scene);
K4_API int k4_scene_destroy(k4_scene* scene);
K4_API size_t k4_scene_device_bytes(const k4_scene* scene);
/* The fastest K4_MLP_* mode built for this scene's rgbnet shape: K4_MLP_TCGEN05_WS for every 3-layer MLP that one of
 * the instantiated tcgen05 configurations covers (csrc/k4_ws_cfgs.h: k0 <= 16 channels, <= 6 view / <= 5 position
 * frequencies, width <= 128, rgbnet_direct either way), else the mma.sync / fp32 kernels.  k4_scene_ws_config: the
 * id of that configuration, -1 if none. */
K4_API int k4_scene_best_mlp_mode(const k4_scene* scene);
K4_API int k4_scene_ws_config(const k4_scene* scene);

/* Scratch needed by k4_render_rays for n_rays (a few bytes of scheduler state). */
K4_API size_t k4_render_workspace_bytes(const k4_scene* scene, int64_t n_rays);

/*
 * Replaces: DirectVoxGO.forward (lib/dvgo.py:327-448) / DirectMPIGO.forward (lib/dmpigo.py:292-427)
 * at inference, i.e. the whole chain sample_pts_on_rays -> maskcache_lookup -> grid_sample ->
 * raw2alpha -> alpha2weight -> grid_sample -> rgbnet -> segment_coo in ONE launch.
 * d_rays_o/d_rays_d/d_viewdirs: [N,3] contiguous fp32 (the reference CHECK_CONTIGUOUS contract,
 * lib/cuda/render_utils.cpp:46-48).  Any N >= 0.
 */
K4_API int k4_render_rays(const k4_scene* scene, const k4_render_args* args,
                   const float* d_rays_o, const float* d_rays_d, const float* d_viewdirs,
                   int64_t n_rays, const k4_render_out* out,
                   void* d_workspace, size_t workspace_bytes, k4_stream_t stream);

/*
 * Multi-GPU frames (the reference has no multi-GPU path -- run_sr.py:99-128 renders every frame on
 * one device): the same fused launch for ONE RANK'S ROWS of a frame that is sharded in 8-row blocks dealt round-robin
 * over `world` ranks, with the exchange fused into the kernel: every ray's results are stored straight into the
 * image-order frame of EVERY rank -- d_frame[i] are the local frame and the peers' frames mapped with k4_peer_open
 * (NVLink P2P stores) -- instead of a local band buffer followed by an all-gather and a transpose.  The rays passed are
 * the rank's rows in order (k4_make_rays_rows), n_rays = rows * frame_w; local row r is image row
 * ((r / 8) * world + rank) * 8 + r % 8.  A frame is 5 * n_full floats: rgb [n_full,3] | depth [n_full] | alphainv
 * [n_full], n_full >= the padded image (world * ceil(ceil(H/8) / world) * 8 rows).  `out` may be NULL or carry the
 * optional parity outputs (d_ray_stats / d_t_minmax / d_counters, local ray order).  The stores are complete when the
 * launch has completed on this rank's stream; the caller orders the ranks (any barrier that follows it in stream order).
 */
typedef struct k4_frame_dst {
    int32_t n_dst;                /* frames to write: 1 .. K4_MAX_PEERS                                  */
    int32_t rank, world;          /* block-cyclic position of the rows of this call                      */
    int32_t frame_w;              /* W                                                                   */
    int64_t n_full;               /* rays per (padded) frame                                             */
    float* d_frame[K4_MAX_PEERS];
} k4_frame_dst;
K4_API int k4_render_rays_frames(const k4_scene* scene, const k4_render_args* args,
                   const float* d_rays_o, const float* d_rays_d, const float* d_viewdirs,
                   int64_t n_rays, const k4_frame_dst* dst, const k4_render_out* out,
                   void* d_workspace, size_t workspace_bytes, k4_stream_t stream);

/* Device memory other processes of the node can map (one process per GPU): k4_peer_alloc = cudaMalloc + zero fill
 * (pool / async allocations cannot be exported), k4_peer_export fills the 64-byte CUDA IPC handle to send to the other
 * ranks (any transport: torch.distributed.all_gather_object), k4_peer_open maps another PROCESS's allocation into the
 * current device's address space with peer access enabled, k4_peer_close unmaps it. */
K4_API int k4_peer_enable_all(void);                 /* peer access from the current device to every device that allows it */
K4_API int k4_peer_alloc(size_t bytes, void** d_ptr);
K4_API int k4_peer_free(void* d_ptr);
K4_API int k4_peer_export(void* d_ptr, unsigned char handle[64]);
K4_API int k4_peer_open(const unsigned char handle[64], void** d_ptr);
K4_API int k4_peer_close(void* d_ptr);

/*
 * Replaces: get_rays_of_a_view (lib/dvgo.py:516-582) feeding the chunk loop of render_viewpoints
 * (run_sr.py:99-128): pixel-centre rays of an HxW pinhole view generated on the device.
 * h_K: 9 floats row major, h_c2w: 12 floats (3x4) row major, both HOST pointers.
 * d_rays_o/d_rays_d/d_viewdirs: [H*W,3] outputs.
 */
K4_API int k4_make_rays(const float* h_K, const float* h_c2w, int32_t H, int32_t W, int32_t ndc,
                 int32_t inverse_y, int32_t flip_x, int32_t flip_y,
                 float* d_rays_o, float* d_rays_d, float* d_viewdirs, k4_stream_t stream);

/* Same for a subset of the image rows (multi-GPU: a rank generates only ITS rows of the frame, SURVEY.md
 * section 8e): d_rows = DEVICE array of n_rows image-row indices (NULL = all H rows in order);
 * outputs are [n_rows*W,3] in the order of d_rows.  Values are identical to the same pixels of k4_make_rays. */
K4_API int k4_make_rays_rows(const float* h_K, const float* h_c2w, int32_t H, int32_t W, int32_t ndc,
                 int32_t inverse_y, int32_t flip_x, int32_t flip_y, const int32_t* d_rows, int32_t n_rows,
                 float* d_rays_o, float* d_rays_d, float* d_viewdirs, k4_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * VC-Decoder: SFTNet (lib/sr_esrnet.py:400-465), the x4 SFT-RRDB upsampler that turns the marcher's
 * rgb_feature + depth images into the 4K frame (run_sr.py:1353-1387).
 *
 * h_params: HOST array of DEVICE pointers to the fp32 parameters (weight, bias alternating) of the
 * 229 convolutions in this fixed order -- the reference module's registration order:
 *   conv_first; CondNet.0, CondNet.2, CondNet.4, CondNet.6;
 *   for i in blocks: for j in rdb1..rdb3: conv1..conv5,
 *                                          sft0.{SFT_scale_conv0, SFT_scale_conv1, SFT_shift_conv0, SFT_shift_conv1},
 *                                          sft1.{same four};
 *                    body.i.sft0.{same four};
 *   sftbody.{same four}; conv_body; conv_up1; conv_up2; conv_hr; conv_last.
 * Only the shipped configuration SFTNet(n_in_colors=3, scale=4, num_feat=64, num_block<=32,
 * num_grow_ch=32, num_cond=1, dswise=False) (run_sr.py:1353) is built; others return K4_ERR_UNSUPPORTED.
 */
typedef struct k4_srnet_desc {
    int32_t n_in_colors, scale, num_feat, num_block, num_grow_ch, num_cond;
    int32_t n_params;                  /* = 2 * number of convolutions */
    int32_t reserved0;
    const float* const* h_params;
} k4_srnet_desc;

K4_API int k4_srnet_create(const k4_srnet_desc* desc, k4_stream_t stream, k4_srnet** out_net);
K4_API int k4_srnet_destroy(k4_srnet* net);
K4_API size_t k4_srnet_workspace_bytes(const k4_srnet* net, int32_t h, int32_t w);

/* Replaces: SFTNet.forward(x, cond) (lib/sr_esrnet.py:446-465) for one tile.
 * d_x [3,h,w], d_cond [1,h,w] planar fp32 (the reference's NCHW, batch 1); d_out [3,4h,4w] fp32.
 * Tiling (SFTNet.tile_process, lib/sr_esrnet.py:467-527) is host logic above this call. */
K4_API int k4_srnet_forward(const k4_srnet* net, const float* d_x, const float* d_cond, int32_t h, int32_t w,
                            float* d_out, void* d_workspace, size_t workspace_bytes, k4_stream_t stream);

/* The same forward when only a block of the tile's output is wanted -- what SFTNet.tile_process does with every tile
 * (lib/sr_esrnet.py:509-523 crops the 10-pixel pad away and copies the rest into the frame) and what the multi-GPU
 * decoder does with a tile's row parts.  keep_* select LR rows [keep_y0,keep_y1) x columns [keep_x0,keep_x1) of the
 * h x w input; their x4 pixels are written straight to d_out[c*out_plane_stride + Y*out_row_stride + X] with (Y, X)
 * relative to the block's first pixel (pass a pointer into the frame and the frame's strides).  Every layer computes only
 * the rows inside the remaining receptive field of the kept rows; the values are bit-identical to k4_srnet_forward's. */
K4_API int k4_srnet_forward_roi(const k4_srnet* net, const float* d_x, const float* d_cond, int32_t h, int32_t w,
                                int32_t keep_y0, int32_t keep_y1, int32_t keep_x0, int32_t keep_x1,
                                float* d_out, int64_t out_plane_stride, int64_t out_row_stride,
                                void* d_workspace, size_t workspace_bytes, k4_stream_t stream);

/* k4_srnet_forward_roi whose last convolution stores the kept block into n_extra MORE frames as well (h_extra: HOST array
 * of device pointers with d_out's meaning and strides -- the same window of the other ranks' frames, mapped with
 * k4_peer_open): the multi-GPU decoder's exchange fused into the kernel that produces the pixels, instead of an
 * all-gather of packed blocks and an assembly pass.  n_extra <= K4_MAX_PEERS - 1. */
K4_API int k4_srnet_forward_roi_peers(const k4_srnet* net, const float* d_x, const float* d_cond, int32_t h, int32_t w,
                                      int32_t keep_y0, int32_t keep_y1, int32_t keep_x0, int32_t keep_x1,
                                      float* d_out, int64_t out_plane_stride, int64_t out_row_stride,
                                      int32_t n_extra, float* const* h_extra,
                                      void* d_workspace, size_t workspace_bytes, k4_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Op-level surface: the 13 functions of the reference extension `render_utils_cuda`
 * (lib/cuda/render_utils.cpp:170-184), bit-identical results, caller-allocated outputs, no host sync.
 * Needed for training compatibility (lib/dvgo.py:453-511 autograd shims), not for inference.
 * All pointers are device pointers; masks are bytes (torch.bool); ids are int64.
 *   reference function                      entry point(s)
 *   infer_t_minmax            (.cpp:50)     k4_op_infer_t_minmax
 *   infer_n_samples           (.cpp:60)     k4_op_infer_n_samples
 *   infer_ray_start_dir       (.cpp:68)     k4_op_infer_ray_start_dir
 *   sample_pts_on_rays        (.cpp:76)     the three above + cumsum (host side) + k4_op_fill_ray_step_ids + k4_op_sample_pts
 *   sample_ndc_pts_on_rays    (.cpp:92)     k4_op_sample_ndc_pts
 *   sample_bg_pts_on_rays     (.cpp:104)    k4_op_sample_bg_pts
 *   maskcache_lookup          (.cpp:112)    k4_op_maskcache_lookup       (out must be pre-zeroed, as .cu:405)
 *   raw2alpha[_nonuni]        (.cpp:124,131) k4_op_raw2alpha             (interval_v = NULL | per-point intervals)
 *   raw2alpha[_nonuni]_backward (.cpp:138,146) k4_op_raw2alpha_backward
 *   alpha2weight              (.cpp:154)    k4_op_alpha2weight           (outputs pre-filled 0/1/1/0/0 as .cu:624-628)
 *   alpha2weight_backward     (.cpp:162)    k4_op_alpha2weight_backward  (grad pre-zeroed as .cu:684)
 */
K4_API int k4_op_infer_t_minmax(const float* d_rays_o, const float* d_rays_d, const float* d_xyz_min, const float* d_xyz_max,
                                float near_, float far_, int64_t n_rays, float* d_t_min, float* d_t_max, k4_stream_t stream);
K4_API int k4_op_infer_n_samples(const float* d_rays_d, const float* d_t_min, const float* d_t_max, float stepdist, int64_t n_rays,
                                 int64_t* d_n_samples, k4_stream_t stream);
K4_API int k4_op_infer_ray_start_dir(const float* d_rays_o, const float* d_rays_d, const float* d_t_min, int64_t n_rays,
                                     float* d_rays_start, float* d_rays_dir, k4_stream_t stream);
K4_API int k4_op_fill_ray_step_ids(const int64_t* d_n_steps_cumsum, int64_t n_rays, int64_t total, int64_t* d_ray_id,
                                   int64_t* d_step_id, k4_stream_t stream);
K4_API int k4_op_sample_pts(const float* d_rays_start, const float* d_rays_dir, const float* d_xyz_min, const float* d_xyz_max,
                            const int64_t* d_ray_id, const int64_t* d_step_id, float stepdist, int64_t total, float* d_pts,
                            uint8_t* d_mask_outbbox, k4_stream_t stream);
K4_API int k4_op_sample_ndc_pts(const float* d_rays_o, const float* d_rays_d, const float* d_xyz_min, const float* d_xyz_max,
                                int32_t N_samples, int64_t n_rays, float* d_pts, uint8_t* d_mask_outbbox, k4_stream_t stream);
K4_API int k4_op_sample_bg_pts(const float* d_rays_o, const float* d_rays_d, const float* d_t_max, float bg_preserve,
                               int32_t N_samples, int64_t n_rays, float* d_pts, k4_stream_t stream);
K4_API int k4_op_maskcache_lookup(const uint8_t* d_world, const float* d_xyz, uint8_t* d_out, const float* d_scale,
                                  const float* d_shift, int32_t sz_i, int32_t sz_j, int32_t sz_k, int64_t n_pts, k4_stream_t stream);
K4_API int k4_op_raw2alpha(const float* d_density, float shift, float interval, const float* d_interval_v, int64_t n_pts,
                           float* d_exp, float* d_alpha, k4_stream_t stream);
K4_API int k4_op_raw2alpha_backward(const float* d_exp, const float* d_grad_back, float interval, const float* d_interval_v,
                                    int64_t n_pts, float* d_grad, k4_stream_t stream);
K4_API int k4_op_alpha2weight(const float* d_alpha, const int64_t* d_ray_id, int64_t n_rays, int64_t n_pts, float* d_weight,
                              float* d_T, float* d_alphainv_last, int64_t* d_i_start, int64_t* d_i_end, k4_stream_t stream);
K4_API int k4_op_alpha2weight_backward(const float* d_alpha, const float* d_weight, const float* d_T, const float* d_alphainv_last,
                                       const int64_t* d_i_start, const int64_t* d_i_end, int64_t n_rays,
                                       const float* d_grad_weights, const float* d_grad_last, float* d_grad, k4_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Grid maintenance and optimiser steps either side of the render path during training
 * (SURVEY.md section 8 f-4).  Device pointers unless prefixed h_; caller's stream; no host sync.
 *   reference                                                       entry point
 *   total_variation_cuda.total_variation_add_grad (lib/cuda/total_variation.cpp:16-20; param
 *       [1,C,I,J,K] contiguous, grad updated in place; !dense_mode skips grad==0)   k4_op_total_variation_add_grad
 *   adam_upd_cuda.adam_upd / masked_adam_upd / adam_upd_with_perlr (lib/cuda/adam_upd.cpp:37-75;
 *       d_perlr != NULL selects the per-voxel-lr variant, else skip_zero_grad the masked one;
 *       `step` is the 1-based step count after the increment, lib/masked_adam.py:57)  k4_op_adam_upd
 *   update_occupancy_cache (lib/dvgo.py:224-233, lib/dmpigo.py:212-224): alpha of the trilinear density
 *       at the occupancy grid's points (d_lx/ly/lz = the three torch.linspace vectors), then
 *       mask &= max_pool3d(alpha, 3, stride 1, pad 1) > thres              k4_op_grid_alpha, k4_op_maxpool3_thres_and
 *   DenseGrid.scale_volume_grid (lib/grid.py:130-135: F.interpolate(mode='trilinear',
 *       align_corners=True) of a [1,C,X,Y,Z] grid to [1,C,X2,Y2,Z2])        k4_op_resample_trilinear
 */
K4_API int k4_op_total_variation_add_grad(const float* d_param, float* d_grad, float wx, float wy, float wz, int32_t dense_mode,
                                          int64_t n, int32_t sz_i, int32_t sz_j, int32_t sz_k, k4_stream_t stream);
K4_API int k4_op_adam_upd(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, const float* d_perlr,
                          int64_t n, int32_t step, float beta1, float beta2, float lr, float eps, int32_t skip_zero_grad,
                          k4_stream_t stream);
K4_API int k4_op_grid_alpha(const float* d_density, int32_t X, int32_t Y, int32_t Z, const float* h_xyz_min, const float* h_xyz_max,
                            const float* d_lx, const float* d_ly, const float* d_lz, int32_t mX, int32_t mY, int32_t mZ,
                            float shift, float interval, float* d_alpha, k4_stream_t stream);
K4_API int k4_op_maxpool3_thres_and(const float* d_alpha, int32_t mX, int32_t mY, int32_t mZ, float thres, uint8_t* d_mask,
                                    k4_stream_t stream);
/* ub360_utils_cuda.cumdist_thres (lib/cuda/ub360_utils.cpp:20-22; dist [n_rays, n_pts] -> bool mask, bit-identical) */
K4_API int k4_op_cumdist_thres(const float* d_dist, float thres, int64_t n_rays, int64_t n_pts, uint8_t* d_mask, k4_stream_t stream);
K4_API int k4_op_resample_trilinear(const float* d_src, int32_t C, int32_t X, int32_t Y, int32_t Z, float* d_dst, int32_t X2,
                                    int32_t Y2, int32_t Z2, k4_stream_t stream);

/* DenseGrid.forward with autograd (lib/grid.py:117-128): trilinear lookup of world points in a planar [1,C,X,Y,Z]
 * grid -- ((xyz-min)/(max-min)).flip(-1)*2-1, ATen grid_sampler_3d (bilinear, zeros, align_corners=True) and the
 * [C,M]->[M,C] transpose in one launch -- and its backward: the scatter of the output gradient [M,C] into the grid
 * gradient [C,X,Y,Z] (ACCUMULATED with atomics: the caller zero-fills), which the reference obtains from ATen's
 * grid_sampler_3d_backward.  h_xyz_min/max: 3 host floats each.  d_xyz [M,3], d_out / d_grad_out [M,C]. */
K4_API int k4_op_grid_sample(const float* d_grid, int32_t C, int32_t X, int32_t Y, int32_t Z, const float* h_xyz_min,
                             const float* h_xyz_max, const float* d_xyz, int64_t M, float* d_out, k4_stream_t stream);
K4_API int k4_op_grid_sample_backward(const float* d_grad_out, int32_t C, int32_t X, int32_t Y, int32_t Z, const float* h_xyz_min,
                                      const float* h_xyz_max, const float* d_xyz, int64_t M, float* d_grad_grid, k4_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* K4NERF_H_ */
