// Deterministic cross-CTA column reduction (no floating-point atomics): used for BatchNorm statistics, where a sum whose
// rounding depends on the arrival order of CTAs makes two forward passes over the same input differ in the last bit -- and
// 21 train-mode BatchNorm layers amplify that to ~1e-3 at the BEV map.
//
// Every CTA contributes one partial per column.  CTAs are grouped by index (`group` consecutive CTAs); the last CTA of a
// group to arrive sums the group's partials IN INDEX ORDER, the last group to finish sums the group results in index
// order.  Which CTA does the summing varies from run to run, the order of every floating-point addition does not.
//
// Workspace (caller-provided, ud_det_reduce_bytes): [0, 4096) int32 arrival counters, zeroed by the launcher before the
// kernel ([0] groups done, [1 + g] CTAs of group g done); then n_groups x ncols doubles; then n_ctas x ncols doubles.
#pragma once
#include "common.cuh"

constexpr int UD_DET_COUNTER_BYTES = 4096;
constexpr int UD_DET_MAX_GROUPS = 1023;

static inline int ud_det_group(int ctas) {
    int g = 32;
    while ((ctas + g - 1) / g > UD_DET_MAX_GROUPS) g *= 2;
    return g;
}

static inline size_t ud_det_reduce_bytes(int ctas, int ncols) {
    const int g = ud_det_group(ctas);
    const int groups = (ctas + g - 1) / g;
    return (size_t)UD_DET_COUNTER_BYTES + ((size_t)groups + (size_t)ctas) * ncols * sizeof(double);
}

// Called by ALL threads of every CTA of a 1-D grid, after the CTA's partial for column `tid` (tid < ncols) is in `v`.
// `flag_s`: one int of shared memory.  On return of the last CTA, fin[0 .. ncols) holds the column sums.
__device__ __forceinline__ void ud_det_reduce(double v, int tid, int ncols, unsigned char* __restrict__ ws, int group,
                                              double* __restrict__ fin, int* flag_s) {
    int* counters = reinterpret_cast<int*>(ws);
    const int n_ctas = (int)gridDim.x;
    const int n_groups = (n_ctas + group - 1) / group;
    double* part2 = reinterpret_cast<double*>(ws + UD_DET_COUNTER_BYTES);
    double* part1 = part2 + (size_t)n_groups * ncols;
    const int grp = (int)blockIdx.x / group;
    const int first = grp * group;
    const int size = min(group, n_ctas - first);
    if (tid < ncols) part1[(size_t)blockIdx.x * ncols + tid] = v;
    __threadfence();
    __syncthreads();
    if (tid == 0) *flag_s = atomicAdd(&counters[1 + grp], 1) == size - 1;
    __syncthreads();
    if (!*flag_s) return;
    __threadfence();
    if (tid < ncols) {
        double acc = 0.0;
        for (int j = 0; j < size; ++j) acc += __ldcg(part1 + (size_t)(first + j) * ncols + tid);
        part2[(size_t)grp * ncols + tid] = acc;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) *flag_s = atomicAdd(&counters[0], 1) == n_groups - 1;
    __syncthreads();
    if (!*flag_s) return;
    __threadfence();
    if (tid < ncols) {
        double acc = 0.0;
        for (int g = 0; g < n_groups; ++g) acc += __ldcg(part2 + (size_t)g * ncols + tid);
        fin[tid] = acc;
    }
}
