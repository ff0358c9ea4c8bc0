// k4_internal.cuh -- structures shared by the translation units of libk4nerf.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/k4nerf.h"

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
};

struct k4_scene {
    K4Dev dev;
    int device;
    size_t bytes;
    void* allocs[64];
    int n_allocs;
};

struct K4RenderParams {
    float near_, far_, stepdist, interval, bg, inv_nsamples;
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
