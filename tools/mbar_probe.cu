// Micro-probe: latency of an already-satisfied mbarrier.try_wait / test_wait, of elect.sync, of LDS, and of R2UR-fed
// uniform arithmetic, in a single otherwise idle warp.  Development aid.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mbar_probe tools/mbar_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(long long* out, int spinners) {
    __shared__ __align__(8) unsigned long long bar[4];
    __shared__ __align__(8) unsigned long long never;
    __shared__ int arr[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&never)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[i])) : "memory");
    }
    if (threadIdx.x < 64) arr[threadIdx.x] = threadIdx.x;
    __syncthreads();
    __shared__ volatile int stop;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (warp == 0) {
        long long t0, t1;
        uint32_t done, acc = 0;
        // 1: satisfied try_wait, 64 in a dependent chain (address depends on the previous result)
        t0 = clock64();
        uint32_t a = smem_u32(&bar[0]);
        for (int i = 0; i < 64; ++i) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(a), "r"(0u) : "memory");
            acc += done;
            a = smem_u32(&bar[0]) + ((done - 1u) & 8u);          // done == 1 -> same address
        }
        t1 = clock64();
        if (lane == 0) out[0] = (t1 - t0) / 64;
        // 2: satisfied test_wait
        t0 = clock64();
        for (int i = 0; i < 64; ++i) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(a), "r"(0u) : "memory");
            acc += done;
            a = smem_u32(&bar[0]) + ((done - 1u) & 8u);
        }
        t1 = clock64();
        if (lane == 0) out[1] = (t1 - t0) / 64;
        // 3: dependent LDS chain
        int idx = lane;
        t0 = clock64();
        for (int i = 0; i < 64; ++i) idx = arr[idx & 63];
        t1 = clock64();
        acc += idx;
        if (lane == 0) out[2] = (t1 - t0) / 64;
        // 4: elect.sync
        t0 = clock64();
        for (int i = 0; i < 64; ++i) {
            uint32_t pred;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
            acc += pred;
        }
        t1 = clock64();
        if (lane == 0) out[3] = (t1 - t0) / 64;
        // 5: unsatisfied try_wait (returns false at the time limit): how long does one attempt block
        t0 = clock64();
        for (int i = 0; i < 4; ++i) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&never)), "r"(0u) : "memory");
            acc += done;
        }
        t1 = clock64();
        if (lane == 0) out[4] = (t1 - t0) / 4;
        // 6: clock64 back to back
        t0 = clock64();
        long long s = 0;
        for (int i = 0; i < 64; ++i) s += clock64();
        t1 = clock64();
        acc += (uint32_t)s;
        if (lane == 0) out[5] = (t1 - t0) / 64;
        if (lane == 0) out[7] = acc;
        stop = 1;
    } else if (warp <= spinners) {
        // other warps blocked in try_wait on a barrier that never completes (what idle pipeline roles do)
        uint32_t done = 0;
        while (!stop && !done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&never)), "r"(0u) : "memory");
    }
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    for (int sp : {0, 4, 18}) {
        probe<<<1, 608>>>(d, sp);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("waiting warps %2d: try_wait(satisfied) %lld  test_wait(satisfied) %lld  LDS %lld  elect %lld  try_wait(unsatisfied, one attempt) %lld  clock64 %lld cycles\n",
               sp, h[0], h[1], h[2], h[3], h[4], h[5]);
    }
    return 0;
}
