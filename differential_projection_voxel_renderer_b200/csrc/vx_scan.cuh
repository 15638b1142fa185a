// vx_scan.cuh -- exclusive prefix sum of per-record task counts, one CTA (the secondary rasterizers' work lists:
// n is the number of quads / triangle slots of one call, not a per-frame hot loop).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int VX_SCAN_THREADS = 1024;

// base[i] = sum of count[0 .. i), *total = sum of all
static __global__ void __launch_bounds__(VX_SCAN_THREADS) vx_scan_counts_kernel(const uint32_t *__restrict__ count, int n,
                                                                                unsigned long long *__restrict__ base,
                                                                                unsigned long long *__restrict__ total) {
    __shared__ unsigned long long part[VX_SCAN_THREADS];
    const int tid = threadIdx.x;
    const int per = (n + VX_SCAN_THREADS - 1) / VX_SCAN_THREADS;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    unsigned long long s = 0;
    for (int i = lo; i < hi; ++i) s += count[i];
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < VX_SCAN_THREADS; o <<= 1) { // Hillis-Steele inclusive scan
        const unsigned long long v = tid >= o ? part[tid - o] : 0ull;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    unsigned long long run = part[tid] - s;
    for (int i = lo; i < hi; ++i) {
        base[i] = run;
        run += count[i];
    }
    if (tid == VX_SCAN_THREADS - 1) *total = part[tid];
}

// owner of task t = the right-most record with base <= t (a record without tasks shares the base of its successor, so
// it is never the right-most one among those <= t)
__device__ __forceinline__ int vx_task_owner(const unsigned long long *__restrict__ base, int n, unsigned long long t) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (base[mid] <= t) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
