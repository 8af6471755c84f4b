// Micro-probe: which (TMEM lane, column) does register i of thread t land in for tcgen05.st.16x256b (.x1 / .x2)?
// Development aid for the register -> TMEM gather path; not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_layout_probe tools/tmem_layout_probe.cu && ./tmem_layout_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out /* [128 lanes][16 cols] */) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s;
    const uint32_t lane_q = (uint32_t)(warp * 32) << 16;
    // clear 16 columns of this warp's 32 lanes
    {
        uint32_t z = 0xffffffffu;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(base + lane_q), "r"(z) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // 16x256b.x2 at lane offsets 0 and 16: register i of thread t encodes (half, t, i)
    for (int half = 0; half < 2; ++half) {
        uint32_t r[8];
        for (int i = 0; i < 8; ++i) r[i] = (uint32_t)(half << 16 | lane << 8 | i);
        const uint32_t taddr = base + lane_q + ((uint32_t)(half * 16) << 16);
        asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                     "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(base + lane_q) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 16; ++c) out[tid * 16 + c] = v[c];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32) : "memory");
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 128 * 16 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    static uint32_t h[128 * 16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("tcgen05.st.16x256b.x2: TMEM (lane, column) <- (half, thread, register); warp 0 (lanes 0-31), warp 1 lanes 32-35\n");
    for (int l = 0; l < 36; ++l) {
        printf("lane %3d:", l);
        for (int c = 0; c < 16; ++c) {
            uint32_t v = h[l * 16 + c];
            if (v == 0xffffffffu) printf("   ----  ");
            else printf(" h%u t%02u r%u", v >> 16, (v >> 8) & 0xff, v & 0xff);
        }
        printf("\n");
    }
    return 0;
}
