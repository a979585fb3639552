// scan.cu -- single-CTA exclusive scan over int64 (row / pair offsets; tiny vs. the walks).
#include "gg_common.cuh"

namespace gg {
namespace {
__global__ void __launch_bounds__(1024) exclusive_scan_i64_kernel(long long *a, long long n, long long *total_out) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long v = (i < n) ? a[i] : 0;
        long long x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const long long y = __shfl_up_sync(FULL, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            long long t = warp_tot[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const long long y = __shfl_up_sync(FULL, t, off);
                if (lane >= off) t += y;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const long long before = carry_s + (wid ? warp_tot[wid - 1] : 0) + (x - v);
        if (i < n) a[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { a[n] = carry_s; if (total_out) *total_out = carry_s; }
}
}  // namespace

int launch_exclusive_scan_i64(long long *a, long long n, long long *total_out, cudaStream_t st) {
    exclusive_scan_i64_kernel<<<1, 1024, 0, st>>>(a, n, total_out);
    return check_cuda(cudaGetLastError(), "exclusive scan launch");
}
}  // namespace gg
