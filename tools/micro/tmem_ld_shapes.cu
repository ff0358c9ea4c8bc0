// tmem_ld_shapes.cu -- which (row, column) of a tensor-memory tile each thread receives from tcgen05.ld.16x256b
// (the shape that matches the mma.sync m16n8 accumulator fragment).  The tile is written with tcgen05.st.32x32b
// (thread = row) as value = 1000 * row + column, read back with .16x256b.x1 / .x2 and decoded on the host.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(uint32_t* out) {
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(s32(&tslot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tb = tslot;
    const uint32_t tl = tb + ((uint32_t)(warp * 32) << 16);
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = 1000u * tid + c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                 :: "r"(tl), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    for (int h = 0; h < 2; ++h) {
        uint32_t r[8];
        const uint32_t ta = tb + ((uint32_t)(warp * 32 + 16 * h) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int i = 0; i < 8; ++i) out[((warp * 2 + h) * 32 + lane) * 8 + i] = r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tb), "r"(64u) : "memory");
}

int main() {
    uint32_t* d; cudaMalloc(&d, 4 * 2 * 32 * 8 * 4);
    probe<<<1, 128>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    static uint32_t h[4 * 2 * 32 * 8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    // expected (mma m16n8 accumulator layout per 8-column block j): reg 4j+0,1 -> row base+lane/4, cols 8j+2*(lane%4)+{0,1};
    //                                                               reg 4j+2,3 -> row base+lane/4+8, same cols
    int bad = 0;
    for (int w = 0; w < 4; ++w) for (int hh = 0; hh < 2; ++hh) for (int l = 0; l < 32; ++l) for (int i = 0; i < 8; ++i) {
        const uint32_t v = h[((w * 2 + hh) * 32 + l) * 8 + i];
        const int j = i / 4, k = i % 4;
        const int row = w * 32 + 16 * hh + l / 4 + (k >= 2 ? 8 : 0), col = 8 * j + 2 * (l % 4) + (k & 1);
        if (v != 1000u * row + col) { if (bad < 12) printf("warp %d half %d lane %d reg %d: got row %u col %u, expected row %d col %d\n", w, hh, l, i, v / 1000, v % 1000, row, col); ++bad; }
    }
    printf("{\"probe\": \"tcgen05.ld.16x256b.x2 register layout == mma.sync m16n8 accumulator fragment\", \"mismatches\": %d}\n", bad);
    return 0;
}
