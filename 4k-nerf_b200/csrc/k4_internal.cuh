// k4_internal.cuh -- structures shared by the translation units of libk4nerf.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/k4nerf.h"

#define K4_SKIP_B 4               // occupancy voxels per skip cell edge
#define K4_SKIP_MAXD 48           // distance-field cap (cells); larger true distances are stored as the cap
#define K4_MAX_WIDTH 256          // widest hidden layer the fp32 path accepts
#define K4_MAX_DIM0 128           // widest MLP input the fp32 path accepts

// Device-side view of a scene; passed to kernels by value (__grid_constant__).
struct K4Dev {
    int kind;
    int X, Y, Z;                  // density / k0 grid size (z fastest)
    int C, Cpad;                  // k0 channels, padded to a multiple of 4 in the channel-last copy
    int mX, mY, mZ;               // occupancy mask size
    float xyz_min[3], xyz_max[3], xyz_len[3];
    float m_scale[3], m_shift[3];
    float act_shift, voxel_size, voxel_size_ratio, thres;
    int max_world_size, mpi_depth;
    int depth, width, direct, viewpe, spape;
    int dim0;                     // MLP input width (0 when depth == 0)
    int k0_view_off;              // first k0 channel fed to the MLP (3 when !direct, else 0)
    int n_in[K4_MAX_MLP_LAYERS], n_out[K4_MAX_MLP_LAYERS], ldw[K4_MAX_MLP_LAYERS];
    const float* density;         // [X,Y,Z]
    const float* k0cl;            // [X,Y,Z,Cpad] channel last
    const uint8_t* mask;          // [mX,mY,mZ]
    const float* act_grid;        // [mpi_depth] or nullptr
    const float* wT[K4_MAX_MLP_LAYERS];    // fp32 [n_in][ldw] (transposed, ldw = n_out padded to 4)
    const float* bias[K4_MAX_MLP_LAYERS];  // fp32 [ldw]
    // tensor-core packs (fp16 hi / lo parts, [n_out_pad][k_pad] row major = "B^T", K contiguous)
    const __half* wh[K4_MAX_MLP_LAYERS];
    const __half* wl[K4_MAX_MLP_LAYERS];
    int kpad[K4_MAX_MLP_LAYERS], npad[K4_MAX_MLP_LAYERS];
    // tcgen05 pack: one blob of canonical K-major (no swizzle) UMMA operand tiles, see tc_blob_layout()
    const unsigned char* tc_blob;
    int tc_kpad, tc_width;
    int tc_cfg;                   // id of the K4_WS_CFG_LIST entry (k4_ws_cfgs.h) the blob is packed for, -1: none
    int tc_exact;                 // the model's own shape equals that entry's (k4_march_tc.cu only runs exact shapes)
    // empty-space skipping (k4_march_common.cuh skip_steps): Chebyshev distance field over K4_SKIP_B^3-voxel cells of
    // the occupancy mask, in cells: 0 = the cell holds an occupied voxel, d = every cell within d-1 is empty
    const uint8_t* skip;          // [cX,cY,cZ] or nullptr (skipping off)
    int cX, cY, cZ;
    float m_iscale[3];            // 1 / m_scale
    // DirectContractedVoxGO (lib/dcvgo.py)
    float scene_center[3], scene_radius[3], bg_len, one_plus_bg;
    int world_len;
};

// ---- tcgen05 operand blob (built by k4_scene_create, staged by one TMA bulk copy) -----------------
// Canonical K-major SWIZZLE_NONE layout of a [rows][K] fp16 tile with K/8 16-byte chunks per row:
// byte offset of (row r, chunk kc) = (r/8)*(kchunks*128) + kc*128 + (r%8)*16, i.e. 8x16B "core
// matrices", LBO = 128 B between K chunks, SBO = kchunks*128 B between 8-row groups.
__host__ __device__ inline int tc_canon_off(int r, int kc, int kchunks) {
    return (r >> 3) * (kchunks * 128) + kc * 128 + (r & 7) * 16;
}
struct TcBlobLayout {
    int off_w1, off_w2, off_w3, off_b1, off_b2, off_b3, off_ones, total;
};
// W1 [W][kpad], W2 [W][W], W3 [16][W], bias tiles B1/B2 [W][16], B3 [16][16] (col 0 = fp16(b),
// col 1 = fp16(b - col0)), ONES [128][16] (cols 0,1 = 1): every layer's bias is one extra K=16 MMA.
__host__ __device__ inline TcBlobLayout tc_blob_layout(int kpad, int w) {
    TcBlobLayout L;
    int o = 0;
    L.off_w1 = o; o += w * kpad * 2;
    L.off_w2 = o; o += w * w * 2;
    L.off_w3 = o; o += 16 * w * 2;
    L.off_b1 = o; o += w * 16 * 2;
    L.off_b2 = o; o += w * 16 * 2;
    L.off_b3 = o; o += 16 * 16 * 2;
    L.off_ones = o; o += 128 * 16 * 2;
    L.total = o;
    return L;
}

struct k4_scene {
    K4Dev dev;
    int device;
    size_t bytes;
    void* allocs[64];
    int n_allocs;
};

struct K4RenderParams {
    float near_, far_, stepdist, interval, bg, inv_nsamples;
    const float* t_list;          // DCVGO: per-step ray parameter (device, n_samples entries)
    float dist_thres;             // DCVGO: cumdist_thres threshold (lib/dcvgo.py:283)
    int n_samples;                // MPI: samples per ray; DVGO: depth normaliser only
    int render_depth;
    int image_w, image_h;         // >0: 2-D 8x4 tiles
    long long n_rays;
    long long n_tiles;
    const float* rays_o;
    const float* rays_d;
    const float* viewdirs;
    float* rgb;
    float* depth;
    float* alphainv;
    int* ray_stats;
    float* t_minmax;
    unsigned long long* counters;
    unsigned int* tile_counter;
    // k4_render_rays_frames (multi-GPU frame output): n_dst > 0 -> rgb / depth / alphainv above are unused and the ray of
    // local row r, column c is stored at ray index g = (((r >> 3) * f_world + f_rank) * 8 + (r & 7)) * f_w + c of every
    // frame d_frame[0 .. n_dst) (the local one and the peer-mapped ones): rgb at 3g, depth at 3 f_nfull + g, alphainv
    // at 4 f_nfull + g
    int n_dst, f_rank, f_world, f_w;
    long long f_nfull;
    float* d_frame[K4_MAX_PEERS];
};

// error plumbing (k4_capi.cu)
void k4_set_cuda_error(cudaError_t e, const char* where);
#define K4_CUDA_TRY(expr)                                            \
    do {                                                             \
        cudaError_t _e = (expr);                                     \
        if (_e != cudaSuccess) { k4_set_cuda_error(_e, #expr); return K4_ERR_CUDA; } \
    } while (0)

// launchers
int k4_launch_march(const k4_scene* sc, const K4RenderParams& rp, int mlp_mode, cudaStream_t st);
int k4_launch_march_tc(const k4_scene* sc, K4RenderParams rp, cudaStream_t st);   // K4_ERR_UNSUPPORTED if the shape has no tcgen05 build
bool k4_tc_supported(const K4Dev& v);
bool k4_ws_supported(const K4Dev& v);
int k4_launch_march_ws(const k4_scene* sc, K4RenderParams rp, cudaStream_t st);   // warp-specialised variant (same shapes)
