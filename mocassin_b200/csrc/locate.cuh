// locate.cuh -- the reference's table search, shared by the transport and the dust kernels.
#pragma once
#include <cuda_runtime.h>

// interpolation_mod.f90:48-81 for an ascending axis
__device__ __forceinline__ int locate_axis(const float *xa, int n, float x)
{
    if (x > __ldg(&xa[n - 1])) return n;
    if (x < __ldg(&xa[0])) return 0;
    int lo = 0, hi = n;                  // first 0-based index with xa > x
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(&xa[mid]) > x) hi = mid; else lo = mid + 1;
    }
    // no element > x  <=>  x == xa(n): minloc over an empty mask is 0 -> max(-1,1) = 1
    if (lo >= n) return 1;
    return lo > 1 ? lo : 1;              // (1-based first) - 1 = lo
}
