// Micro-probe: cycles per tcgen05.mma (kind::tf32, M=128) as a function of N, of accumulator dependence
// and of where A lives (smem / TMEM).  Development aid, not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

#define MMA_F16_TS_COLL(NAME, QUAL) \
__device__ __forceinline__ void NAME(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) { \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16" QUAL " [%0], [%1], %2, %3, p;\n\t}\n" \
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory"); }
MMA_F16_TS_COLL(mma_f16_ts_fill, ".collector::a::fill")
MMA_F16_TS_COLL(mma_f16_ts_use, ".collector::a::use")
MMA_F16_TS_COLL(mma_f16_ts_last, ".collector::a::lastuse")
MMA_F16_TS_COLL(mma_f16_ts_discard, ".collector::a::discard")

// mode 11 = mode 10 taken apart: `var` bits switch single factors of the conv kernel's operand pattern off
//   1: one weight plane (b_lo := b_hi)          2: A confined to 32 columns, lo == hi (mode 9's A pattern)
//   4: one accumulator (no rotation)            8: A stage region fixed (lo = hi + 32 kept)
//  16: term order hi.Whi, hi.Wlo, lo.Whi        32: collector hints on the two MMAs that share A_hi (fill, lastuse)
//  64: A_lo directly behind A_hi inside a k16 block pair (lo = hi + 8, k16 stride 16)
// mode 9 = bf16 TS same accumulator (K=16), 10 = bf16 TS in the conv kernel's hi/lo pattern (3 MMAs per 16 channels, two
// weight planes, four accumulator tiles round robin every 12 MMAs)
// mode: 0 = SS same accumulator, 1 = SS alternating 2 accumulators, 2 = TS same, 3 = TS alternating,
//       4 = SS 4 accumulators round robin, 5 = bf16 SS same accumulator (K=16)
template <int mode>
__global__ void __launch_bounds__(256, 1) probe(int N, int M, int n_mma, long long* out, int st_iters, int var) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ __align__(8) unsigned long long bar2[4];
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    __shared__ volatile int stop_flag;
    if (threadIdx.x == 0) stop_flag = 0;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 256) reinterpret_cast<float*>(smem_raw)[i] = 0.001f * (i & 63);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t fmt = (mode == 5 || mode == 9 || mode == 10 || mode == 11 || mode == 12) ? 1u : 2u;      // bf16 : tf32
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (warp == 1) {
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 3; ++rep) {
            t0 = clock64();
            if (mode == 7 || mode == 8) {
                // stage loop: wait on the barrier committed 2 (mode 7) / 4 (mode 8) groups ago, then 16 MMAs + commit
                const int depth = mode == 7 ? 2 : 4;
                const int groups = n_mma / 16;
                for (int g = 0; g < groups; ++g) {
                    if (g >= depth) {
                        const int gg = g - depth;
                        uint32_t done;
                        do {
                            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                         : "=r"(done) : "r"(smem_u32(&bar2[gg % 4])), "r"((uint32_t)(((gg / 4) + rep * ((groups + 3) / 4)) & 1)) : "memory");
                        } while (!done);
                    }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one_sync()) {
                        const uint64_t db = desc_sw128(base + 16384);
                        for (int i = 0; i < 16; ++i)
                            mma_ts(tmem, tmem + 480 + (i & 3) * 8, db + (uint64_t)((i & 3) * 2), idesc, 1u);
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[g % 4])) : "memory");
                    }
                    __syncwarp();
                }
                if (elect_one_sync())
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                __syncwarp();
            } else if (mode == 12) {
                // mode 11's pattern (var 0) with a tcgen05.commit to a barrier nobody waits on after every `var` MMAs
                if (elect_one_sync()) {
                    const uint64_t bh = desc_sw128(base + 16384), bl = desc_sw128(base + 16384 + 8192);
                    const int stages = n_mma / 12;
                    int since = 0;
                    for (int s = 0; s < stages; ++s) {
                        const uint32_t a_base = tmem + 256 + (uint32_t)((s & 3) * 64);
                        const uint32_t dd = tmem + (uint32_t)((s & 3) * (N > 64 ? 0 : 64));
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16) {
                            const uint32_t a_hi = a_base + (uint32_t)k16 * 8u, a_lo = a_hi + 32u;
                            const uint64_t adv = (uint64_t)(k16 * 2);
                            mma_f16_ts(dd, a_lo, bh + adv, idesc, 1u); mma_f16_ts(dd, a_hi, bl + adv, idesc, 1u); mma_f16_ts(dd, a_hi, bh + adv, idesc, 1u);
                            since += 3;
                            if (var > 0 && since >= var) {
                                since = 0;
                                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[s & 3])) : "memory");
                            }
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                }
            } else if (mode == 11) {
                if (elect_one_sync()) {
                    const uint64_t bh = desc_sw128(base + 16384);
                    const uint64_t bl = (var & 1) ? bh : desc_sw128(base + 16384 + 8192);
                    const int stages = n_mma / 12;
                    for (int s = 0; s < stages; ++s) {
                        const uint32_t sa = (var & 8) ? 0u : (uint32_t)((s & 3) * 64);
                        const uint32_t a_base = (var & 2) ? tmem + 480 : tmem + 256 + sa;
                        const uint32_t lo_off = (var & 2) ? 0u : ((var & 64) ? 8u : 32u);
                        const uint32_t kstride = (var & 64) ? 16u : 8u;
                        const uint32_t dd = (var & 4) ? tmem : tmem + (uint32_t)((s & 3) * (N > 64 ? 0 : 64));
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16) {
                            const uint32_t a_hi = a_base + (uint32_t)k16 * kstride, a_lo = a_hi + lo_off;
                            const uint64_t adv = (uint64_t)(k16 * 2);
                            if (var & 32) {
                                if (var & 16) { mma_f16_ts_fill(dd, a_hi, bh + adv, idesc, 1u); mma_f16_ts_last(dd, a_hi, bl + adv, idesc, 1u); mma_f16_ts_discard(dd, a_lo, bh + adv, idesc, 1u); }
                                else { mma_f16_ts_discard(dd, a_lo, bh + adv, idesc, 1u); mma_f16_ts_fill(dd, a_hi, bl + adv, idesc, 1u); mma_f16_ts_last(dd, a_hi, bh + adv, idesc, 1u); }
                            } else if (var & 16) { mma_f16_ts(dd, a_hi, bh + adv, idesc, 1u); mma_f16_ts(dd, a_hi, bl + adv, idesc, 1u); mma_f16_ts(dd, a_lo, bh + adv, idesc, 1u); }
                            else { mma_f16_ts(dd, a_lo, bh + adv, idesc, 1u); mma_f16_ts(dd, a_hi, bl + adv, idesc, 1u); mma_f16_ts(dd, a_hi, bh + adv, idesc, 1u); }
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                }
            } else
            if (elect_one_sync()) {
                const uint64_t da = desc_sw128(base), db = desc_sw128(base + 16384);
#pragma unroll 12
                for (int i = 0; i < n_mma; ++i) {
                    const uint64_t adv = (uint64_t)((i & 3) * 2);
                    uint32_t d = tmem;
                    if (mode == 1 || mode == 3) d = tmem + (uint32_t)((i & 1) * 256);
                    if (mode == 4) d = tmem + (uint32_t)((i & 3) * 128);
                    if (mode == 9) mma_f16_ts(d, tmem + 480 + (i & 3) * 8, db + adv, idesc, 1u);
                    else if (mode == 10) {
                        const int k16 = (i / 3) & 3, term = i % 3, stage = i / 12;
                        const uint32_t a_hi = tmem + 256 + (uint32_t)((stage & 3) * 64 + k16 * 8), a_lo = a_hi + 32;
                        const uint64_t b_hi = db + (uint64_t)(k16 * 2), b_lo = desc_sw128(base + 16384 + 8192) + (uint64_t)(k16 * 2);
                        const uint32_t dd = tmem + (uint32_t)((stage & 3) * (N > 64 ? 0 : 64));
                        mma_f16_ts(dd, term == 0 ? a_lo : a_hi, term == 1 ? b_lo : b_hi, idesc, 1u);
                    }
                    else if (mode == 2 || mode == 3) mma_ts(d, tmem + 480 + (i & 3) * 8, db + adv, idesc, 1u);
                    else if (mode == 5) mma_f16_ss(d, da + adv, db + adv, idesc, 1u);
                    else mma_ss(d, da + adv, db + adv, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(smem_u32(&bar)), "r"((uint32_t)(rep & 1)) : "memory");
            } while (!done);
            t1 = clock64();
        }
        if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
        stop_flag = 1;
    } else if (warp >= 4 && st_iters > 0) {
        // concurrent TMEM writes: each warp stores 2 x 32 columns into its lane quadrant, columns [256, 480)
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t v = threadIdx.x;
        int it = 0;
        while (!stop_flag && it < st_iters) {
            const uint32_t taddr = tmem + lane_addr + 256 + (uint32_t)((it % 3) * 64);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + 32), "r"(v) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            ++it;
            // pace: roughly one 64-column store per `pace` cycles
            long long t = clock64();
            while (clock64() - t < 300) {}
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 148 * sizeof(long long));
    const char* names[] = {"SS same acc", "SS 2 accs", "TS same acc", "TS 2 accs", "SS 4 accs", "bf16 SS same acc", "?", "TS staged d2", "TS staged d4", "bf16 TS same acc", "bf16 TS hi/lo x3"};
    const int Ns[] = {16, 32, 64, 128, 256};
    for (int grid : {148})
        for (int mode : {2})
            for (int N : Ns) {
                if ((mode == 1 || mode == 3) && N > 256) continue;
                if (mode == 4 && N > 128) continue;
                if ((mode == 2 || mode == 3 || mode >= 7) && N > 224) continue;
                if (mode == 10 && N > 128) continue;   // acc must not overlap the A columns at 480
                for (int n_mma : {12, 48, 240}) {
                  for (int st_iters : {0, 100000}) {
                    if (st_iters && n_mma != 240) continue;
                    auto launch = [&](auto kern) {
                        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
                        kern<<<grid, 256, 64 * 1024>>>(N, 128, n_mma, d_out, st_iters, 0);
                    };
                    if (mode == 2) launch(probe<2>); else if (mode == 9) launch(probe<9>); else launch(probe<10>);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[148];
                    cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                    long long mx = 0;
                    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                    printf("grid %3d  %-18s N=%3d  n_mma=%3d  st=%d cycles=%7lld  per_mma=%6.1f\n", grid, names[mode], N, n_mma,
                           st_iters > 0, mx, (double)mx / n_mma);
                  }
                }
            }
    for (int N : {32, 64, 128})
        for (int mode : {7, 8})
            for (int n_mma : {64, 256, 1024}) {
                auto kern = mode == 7 ? probe<7> : probe<8>;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
                kern<<<148, 256, 64 * 1024>>>(N, 128, n_mma, d_out, 0, 0);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h[148];
                cudaMemcpy(h, d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("mode %d (tf32 TS, 16 MMAs + commit per group, wait on the group committed %d groups ago) N=%3d n_mma=%4d cycles=%7lld per_mma=%6.1f\n",
                       mode, mode == 7 ? 2 : 4, N, n_mma, mx, (double)mx / n_mma);
            }
    for (int N : {16, 32, 64, 128})
        for (int var : {0, 3, 6, 12, 24, 48}) {
            cudaFuncSetAttribute(probe<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            probe<12><<<148, 256, 64 * 1024>>>(N, 128, 480, d_out, 0, var);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("mode 12 commit every %2d MMAs N=%3d cycles=%7lld per_mma=%6.1f\n", var, N, mx, (double)mx / 480);
        }
    for (int N : {16, 32, 64, 128})
        for (int var : {0, 1, 2, 4, 8, 16, 32, 48, 64, 1 | 4, 1 | 8, 4 | 8, 1 | 4 | 8, 1 | 2 | 4, 16 | 64, 1 | 64, 1 | 4 | 64, 32 | 64}) {
            cudaFuncSetAttribute(probe<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            probe<11><<<148, 256, 64 * 1024>>>(N, 128, 240, d_out, 0, var);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("mode 11 var=%3d N=%3d cycles=%7lld per_mma=%6.1f\n", var, N, mx, (double)mx / 240);
        }
    return 0;
}
