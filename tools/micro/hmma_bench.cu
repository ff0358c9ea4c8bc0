// Microbenchmark: legacy warp-level tensor-core path (mma.sync.m16n8k16 f16->f32, SASS HMMA) on
// sm_100a -- peak issue rate, and the rate when every 4 MMAs need one ldmatrix.x4 of B from shared
// memory (the access pattern of the fused marcher's MLP, k4_march_mma.cuh).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int WITH_LDSM>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    __shared__ __align__(16) unsigned char sm[48 * 1024];
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u;
    __syncthreads();
    float acc[8][4];
    uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
    for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            uint32_t b0 = 0x3c003c00u, b1 = 0x3c003c00u, b2 = 0x3c003c00u, b3 = 0x3c003c00u;
            if (WITH_LDSM) {
                uint32_t addr = base + ((it * 4 + j) & 63) * 512;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                             : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
            }
            mma16816(acc[j], a, b0, b1);
            mma16816(acc[j + 1], a, b2, b3);
            mma16816(acc[(j + 4) & 7], a, b0, b1);
            mma16816(acc[(j + 5) & 7], a, b2, b3);
        }
    }
    float s = 0.f;
    for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) s += acc[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int variant = 0; variant < 2; ++variant) {
        for (int bps = 1; bps <= 4; bps *= 2) {
            const int iters = 20000;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (variant == 0) k<0><<<sms * bps, 256>>>(out, iters); else k<1><<<sms * bps, 256>>>(out, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                const double mmas = (double)sms * bps * 8 * iters * 16;
                const double tflops = mmas * 2.0 * 16 * 8 * 16 / (ms * 1e-3) / 1e12;
                if (rep == 2)
                    printf("{\"micro\": \"hmma_m16n8k16_f16\", \"ldmatrix_b\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.1f, \"mma_per_clk_per_sm_at_1.9GHz\": %.3f}\n",
                           variant, bps * 8, ms, tflops, mmas / sms / (ms * 1e-3) / 1.9e9);
            }
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
