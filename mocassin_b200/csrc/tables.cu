// tables.cu -- table preparation, K4 fold epilogue and K1 opacity assembly kernels.
#include "types.h"

#include <cuda_runtime.h>
#include <cstdint>

namespace mcb {

// ---------------------------------------------------------------------------------------
// PDF upload: reference layout T(0:nCells,nbins) (cell fastest) -> [cell][nu] rows.
// 32x32 shared-memory tile transpose, both sides coalesced.  HBM-bound:
// 8 B per element (4 read + 4 write).
// ---------------------------------------------------------------------------------------
__global__ void transpose_pdf_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                     int nRows /*nCells+1*/, int nb)
{
    __shared__ float tile[32][33];
    int c0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int nu = n0 + j, cell = c0 + threadIdx.x;
        if (nu < nb && cell < nRows) tile[j][threadIdx.x] = src[(size_t)nu * nRows + cell];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int cell = c0 + j, nu = n0 + threadIdx.x;
        if (nu < nb && cell < nRows) dst[(size_t)cell * nb + nu] = tile[threadIdx.x][j];
    }
}

// getNu2's linear scan equals a binary search only on non-decreasing rows: verify.
__global__ void check_monotone_kernel(const float *__restrict__ pdfT, int nRows, int nb, int *bad)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)nRows * (size_t)(nb - 1);
    if (i >= total) return;
    size_t row = i / (nb - 1), k = i % (nb - 1);
    if (row == 0) return;                 // row 0 (inactive sink) is never sampled
    // uniforms are < 1: only min(cdf,1) matters to the scan (see mcb200_set_spectra)
    float a = fminf(pdfT[row * nb + k], 1.f), b = fminf(pdfT[row * nb + k + 1], 1.f);
    if (!(b >= a)) atomicExch(bad, 1);
}

cudaError_t launch_transpose_pdf(const float *src, float *dst, int nRows, int nb, cudaStream_t s)
{
    dim3 grid((nRows + 31) / 32, (nb + 31) / 32), block(32, 8);
    transpose_pdf_kernel<<<grid, block, 0, s>>>(src, dst, nRows, nb);
    return cudaGetLastError();
}

cudaError_t launch_check_monotone(const float *pdfT, int nRows, int nb, int *bad, cudaStream_t s)
{
    size_t total = (size_t)nRows * (size_t)(nb - 1);
    unsigned blocks = (unsigned)((total + 255) / 256);
    check_monotone_kernel<<<blocks, 256, 0, s>>>(pdfT, nRows, nb, bad);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// K4 fold epilogue: integer tallies of one transport call -> float32 estimators of the
// reference, then clear the integer tallies.
//   Jste(c,nu) += float(Q * lenUnit) * deltaE / dV(c)      (photon_mod.f90:1563-1574)
//   escapedPackets(c,nu,a) += float(count) * deltaE          (photon_mod.f90:414-462)
// One thread per element, each element touched by exactly one thread -> deterministic.
// HBM-bound: J: 8 B (Q) read + 4 B read; counts: 4 B (Q) read + 4 B read; writes only where Q != 0.
// ---------------------------------------------------------------------------------------
// Q, J point at element `first` of the (0:nCells, nbins) tables (any element: a rank's shard of
// a reduce-scattered range need not start on a plane boundary)
__global__ void fold_j_kernel(unsigned long long *__restrict__ Q, float *__restrict__ J,
                              const float *__restrict__ dV, int nRows, size_t first, size_t total,
                              double lenUnit, float deltaE)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        long long q = (long long)Q[i];
        if (q != 0) {
            int cell = (int)((first + i) % (size_t)nRows);
            float len = (float)((double)q * lenUnit);
            J[i] = J[i] + len * deltaE / dV[cell];
            Q[i] = 0ull;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Multi-GPU merge of the J tallies, fused over NVLink peer memory (replaces MPI_ALLREDUCE of Jste,
// iteration_mod.f90:627, plus the fold): one kernel that
//   1. reads this rank's share of the touched range from EVERY rank's JsteQ (peer loads through
//      NVLink/NVSwitch, 8 B per element and peer) and adds them as integers (exact, order free),
//   2. folds the sum once (same arithmetic as fold_j_kernel) on top of the local float32 Jste,
//   3. stores the new Jste element into every rank's Jste (peer stores, 4 B per element and peer),
//   4. clears this rank's own partial sum.
// Per element (N-1)/N * 12 B cross the links, against 16 B for an int64 all-reduce, and the fold,
// the clear of the share and the "all-gather" ride along in the same pass.  The other ranks'
// shares of the local JsteQ are cleared by the caller once every rank has finished reading.
// Q[r], J[r] = rank r's arrays (r == rank: local pointers).
// ---------------------------------------------------------------------------------------
struct P2PPeers {
    unsigned long long *Q[16];
    float *J[16];
    int nranks, rank;
};

// N = ranks (compile time), U = elements per thread and trip: U*(N-1) independent peer loads are in
// flight per thread before the first is used -- with few peers the links are latency bound otherwise
template <int N, int U>
__global__ void __launch_bounds__(256) p2p_reduce_fold_kernel(const __grid_constant__ P2PPeers P, const float *__restrict__ dV,
                                                              int nRows, size_t first, size_t total, double lenUnit, float deltaE)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
        unsigned long long q[U][N];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + (size_t)u * stride;
#pragma unroll
            for (int r = 0; r < N; ++r) q[u][r] = i < total ? __ldcv(&P.Q[r][first + i]) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + (size_t)u * stride;
            unsigned long long sum = 0;
#pragma unroll
            for (int r = 0; r < N; ++r) sum += q[u][r];
            if (sum != 0ull) {
                const size_t e = first + i;
                int cell = (int)(e % (size_t)nRows);
                float len = (float)((double)(long long)sum * lenUnit);
                float v = P.J[P.rank][e] + len * deltaE / dV[cell];
#pragma unroll
                for (int r = 0; r < N; ++r) P.J[r][e] = v;
                P.Q[P.rank][e] = 0ull;
            }
        }
    }
    __threadfence_system();              // peer stores performed before the kernel counts as complete
}

// any rank count up to 16 (run time)
__global__ void __launch_bounds__(256) p2p_reduce_fold_generic_kernel(const __grid_constant__ P2PPeers P, const float *__restrict__ dV,
                                                                      int nRows, size_t first, size_t total, double lenUnit, float deltaE)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t e = first + i;
        unsigned long long sum = 0;
        for (int r = 0; r < P.nranks; ++r) sum += __ldcv(&P.Q[r][e]);
        if (sum != 0ull) {
            int cell = (int)(e % (size_t)nRows);
            float len = (float)((double)(long long)sum * lenUnit);
            float v = P.J[P.rank][e] + len * deltaE / dV[cell];
            for (int r = 0; r < P.nranks; ++r) P.J[r][e] = v;
            P.Q[P.rank][e] = 0ull;
        }
    }
    __threadfence_system();
}

// Push variant of the merge, second half: the peers have already written their partial sums of this
// rank's share into its receive buffer (nSlots slots of `slotStride` elements, this range at `rOff`
// inside every slot; peer copies through the mapped buffers, posted writes instead of the round
// trips of peer loads).  Sum them with the local partial sum, fold once, store the float32 result
// into every rank's Jste (peer stores), clear the local partial sum.
template <int N>
__global__ void __launch_bounds__(256) p2p_sum_fold_kernel(const __grid_constant__ P2PPeers P, const unsigned long long *__restrict__ recv,
                                                           size_t slotStride, size_t rOff, const float *__restrict__ dV, int nRows,
                                                           size_t first, size_t total, double lenUnit, float deltaE)
{
    const int nr = N > 0 ? N : P.nranks;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long *Q = P.Q[P.rank];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t e = first + i;
        unsigned long long sum = Q[e];
#pragma unroll
        for (int sl = 0; sl < (N > 0 ? N - 1 : 15); ++sl)
            if (sl < nr - 1) sum += recv[(size_t)sl * slotStride + rOff + i];
        if (sum != 0ull) {
            int cell = (int)(e % (size_t)nRows);
            float len = (float)((double)(long long)sum * lenUnit);
            float v = P.J[P.rank][e] + len * deltaE / dV[cell];
#pragma unroll
            for (int r = 0; r < (N > 0 ? N : 16); ++r)
                if (r < nr) P.J[r][e] = v;
            Q[e] = 0ull;
        }
    }
    __threadfence_system();
}

cudaError_t launch_p2p_sum_fold(const P2PPeers &P, const unsigned long long *recv, size_t slotStride, size_t rOff, const float *dV,
                                int nRows, size_t first, size_t total, double lenUnit, float deltaE, int blocks, cudaStream_t s)
{
    if (total == 0) return cudaSuccess;
    switch (P.nranks) {
    case 2: p2p_sum_fold_kernel<2><<<blocks, 256, 0, s>>>(P, recv, slotStride, rOff, dV, nRows, first, total, lenUnit, deltaE); break;
    case 4: p2p_sum_fold_kernel<4><<<blocks, 256, 0, s>>>(P, recv, slotStride, rOff, dV, nRows, first, total, lenUnit, deltaE); break;
    case 8: p2p_sum_fold_kernel<8><<<blocks, 256, 0, s>>>(P, recv, slotStride, rOff, dV, nRows, first, total, lenUnit, deltaE); break;
    default: p2p_sum_fold_kernel<0><<<blocks, 256, 0, s>>>(P, recv, slotStride, rOff, dV, nRows, first, total, lenUnit, deltaE); break;
    }
    return cudaGetLastError();
}

// Push variant, first half as a kernel: every thread block streams the shares of ALL peers at once
// (element i of every peer's share in the same trip), so the stores of a rank are spread over all
// its links all the time instead of one destination after another.
// recv[r] = rank r's receive buffer; this rank writes slot `slot(r)` of it at rOff.
struct P2PPush {
    unsigned long long *recv[16];
    int nranks, rank;
};

template <int N>
__global__ void __launch_bounds__(256) p2p_push_kernel(const __grid_constant__ P2PPush P, const unsigned long long *__restrict__ Q,
                                                       size_t off, size_t count, size_t slotStride, size_t rOff)
{
    const int nr = N > 0 ? N : P.nranks;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        unsigned long long v[N > 0 ? N : 16];
#pragma unroll
        for (int r = 0; r < (N > 0 ? N : 16); ++r)
            if (r < nr && r != P.rank) v[r] = Q[off + (size_t)r * count + i];
#pragma unroll
        for (int r = 0; r < (N > 0 ? N : 16); ++r)
            if (r < nr && r != P.rank) {
                const int slot = P.rank < r ? P.rank : P.rank - 1;
                P.recv[r][(size_t)slot * slotStride + rOff + i] = v[r];
            }
    }
    __threadfence_system();
}

cudaError_t launch_p2p_push(const P2PPush &P, const unsigned long long *Q, size_t off, size_t count, size_t slotStride,
                            size_t rOff, int blocks, cudaStream_t s)
{
    if (count == 0) return cudaSuccess;
    switch (P.nranks) {
    case 2: p2p_push_kernel<2><<<blocks, 256, 0, s>>>(P, Q, off, count, slotStride, rOff); break;
    case 4: p2p_push_kernel<4><<<blocks, 256, 0, s>>>(P, Q, off, count, slotStride, rOff); break;
    case 8: p2p_push_kernel<8><<<blocks, 256, 0, s>>>(P, Q, off, count, slotStride, rOff); break;
    default: p2p_push_kernel<0><<<blocks, 256, 0, s>>>(P, Q, off, count, slotStride, rOff); break;
    }
    return cudaGetLastError();
}

// ---- packed push: 4 bytes per element on the links instead of 8 -----------------------------------
// A rank's partial sum of one (cell, nu) element rarely needs more than 32 bits (the unit is 2^-24 of
// the smallest cell half-width; only the cells next to a source collect more per call).  The packed
// push sends the low words of every element, and the high words only of the blocks of kPackBlk
// consecutive elements in which some high word is non-zero, with one flag byte per block (blocks
// that are zero altogether send the flag alone); the owner
// rebuilds the 64-bit values (low + (high << 32) where flagged) and sums them as before: exact.
// j = index inside this rank's share of ALL exchanged ranges (ranges back to back, rOff of each);
// receive buffer of a rank: low words [slot][j], high words [slot][j], flags [slot][j / kPackBlk].
constexpr int kPackBlk = 256;           // elements per flag (8 per lane of one warp)
struct P2PPackedPeers {
    unsigned int *lo[16];
    unsigned int *hi[16];
    unsigned char *flag[16];             // flags of the blocks sent to rank r, [rOff / kPackBlk + block]: a LOCAL staging
                                         // array, copied to the peer in one piece afterwards (2.3x10^6 one-byte peer
                                         // stores -- partial-sector writes across the link -- bounded the kernel at ~4 ms)
    int nranks, rank;
};

// 256-bit accesses (sm_100): one 32 B sector per lane
__device__ __forceinline__ void ld256_cs(const unsigned long long *p, unsigned long long (&v)[4])
{
    asm volatile("ld.global.cs.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st256_zero(unsigned long long *p)
{
    asm volatile("st.global.v4.u64 [%0], {%1,%1,%1,%1};" ::"l"(p), "l"(0ull) : "memory");
}

// Positions inside a receive slot: an element with ABSOLUTE index e of a share that starts at e0 sits at
// j = rOff + (e - (e0 & ~3)), so that groups of four elements aligned in Q and Jste are aligned in the
// slot too (rOff is a multiple of kPackBlk): every access below is one 16 B or 32 B vector per lane --
// peer stores of 4 B per lane ran at 315 GB/s on NVLink, under half of what 16 B per lane reaches.
// Q = base of the whole array; the share of rank r of this range = [off + r*count, off + (r+1)*count).
__global__ void __launch_bounds__(256) p2p_push_packed_kernel(const __grid_constant__ P2PPackedPeers P, const unsigned long long *__restrict__ Q,
                                                              size_t off, size_t count, size_t rOff, size_t slotStride, size_t flagStride)
{
    // one warp per flag block (kPackBlk = 2 groups of 4 elements per lane): no block-wide barrier
    const size_t nBlk = (count + 3 + kPackBlk - 1) / kPackBlk;
    const int lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nWarps = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t b = warp; b < nBlk; b += nWarps) {
        for (int d = 1; d < P.nranks; ++d) {
            const int r = (P.rank + d) % P.nranks;
            const int slot = P.rank < r ? P.rank : P.rank - 1;
            const size_t e0 = off + (size_t)r * count, end = e0 + count, A = e0 & ~(size_t)3;
            if (A + b * kPackBlk >= end) continue;           // (warp-uniform)
            unsigned long long v[2][4];
            size_t gs[2];
            bool full[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {                    // both loads in flight before the first store
                gs[k] = A + b * kPackBlk + (size_t)(k * 32 + lane) * 4;
                full[k] = gs[k] >= e0 && gs[k] + 4 <= end;
                if (full[k]) ld256_cs(Q + gs[k], v[k]);
                else
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[k][i] = (gs[k] + i >= e0 && gs[k] + i < end) ? __ldcs(Q + gs[k] + i) : 0ull;
            }
            unsigned int anyHi = 0u;
            unsigned int *lo = P.lo[r] + (size_t)slot * slotStride + rOff;
            const unsigned long long nz = v[0][0] | v[0][1] | v[0][2] | v[0][3] | v[1][0] | v[1][1] | v[1][2] | v[1][3];
            if (!__any_sync(0xffffffffu, nz != 0ull)) {       // nothing in this block (37 % of them at 128^3x600): flag only
                if (lane == 0) P.flag[r][rOff / kPackBlk + b] = (unsigned char)2;
                continue;
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const size_t j = gs[k] - A;
                if (full[k]) *reinterpret_cast<uint4 *>(lo + j) = make_uint4((unsigned int)v[k][0], (unsigned int)v[k][1], (unsigned int)v[k][2], (unsigned int)v[k][3]);
                else
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gs[k] + i >= e0 && gs[k] + i < end) lo[j + i] = (unsigned int)v[k][i];
                anyHi |= (unsigned int)((v[k][0] | v[k][1] | v[k][2] | v[k][3]) >> 32);
            }
            const int any = __any_sync(0xffffffffu, anyHi != 0u);
            if (lane == 0) P.flag[r][rOff / kPackBlk + b] = (unsigned char)(any ? 1 : 0);
            if (any) {
                unsigned int *hi = P.hi[r] + (size_t)slot * slotStride + rOff;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const size_t j = gs[k] - A;
                    if (full[k]) *reinterpret_cast<uint4 *>(hi + j) = make_uint4((unsigned int)(v[k][0] >> 32), (unsigned int)(v[k][1] >> 32), (unsigned int)(v[k][2] >> 32), (unsigned int)(v[k][3] >> 32));
                    else
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (gs[k] + i >= e0 && gs[k] + i < end) hi[j + i] = (unsigned int)(v[k][i] >> 32);
                }
            }
        }
    }
    __threadfence_system();
}

// Owner side: rebuild the peers' 64-bit partial sums of this rank's share [first, first+total), add the
// own one, fold (same float32 expression as fold_j_kernel) and store the result into every rank's Jste.
// One group of four elements per thread; a group with a non-zero sum rewrites all four Jste values
// (the unchanged ones are what every rank already holds: Jste of the shared ranges only ever changes here).
template <int N>
__global__ void __launch_bounds__(256) p2p_sum_fold_packed_kernel(const __grid_constant__ P2PPeers P, const unsigned int *__restrict__ lo,
                                                                  const unsigned int *__restrict__ hi, const unsigned char *__restrict__ flag,
                                                                  size_t slotStride, size_t flagStride, size_t rOff,
                                                                  const float *__restrict__ dV, int nRows, size_t first, size_t total,
                                                                  double lenUnit, float deltaE)
{
    const int nr = N > 0 ? N : P.nranks;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long *Q = P.Q[P.rank];
    float *Jown = P.J[P.rank];
    const size_t end = first + total, A = first & ~(size_t)3, nGroups = (end - A + 3) / 4;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < nGroups; g += stride) {
        const size_t gs = A + 4 * g, j = rOff + 4 * g;
        const bool full = gs >= first && gs + 4 <= end;
        unsigned long long sum[4];
        if (full) ld256_cs(Q + gs, sum);
        else
#pragma unroll
            for (int i = 0; i < 4; ++i) sum[i] = (gs + i >= first && gs + i < end) ? Q[gs + i] : 0ull;
#pragma unroll
        for (int sl = 0; sl < (N > 0 ? N - 1 : 15); ++sl)
            if (sl < nr - 1) {
                const unsigned char fb = flag[(size_t)sl * flagStride + j / kPackBlk];   // 0 low words, 1 low + high, 2 nothing
                if (fb == 2) continue;
                const bool f = fb == 1;
                if (full) {
                    const uint4 l = __ldcs(reinterpret_cast<const uint4 *>(lo + (size_t)sl * slotStride + j));
                    uint4 h = make_uint4(0u, 0u, 0u, 0u);
                    if (f) h = __ldcs(reinterpret_cast<const uint4 *>(hi + (size_t)sl * slotStride + j));
                    sum[0] += (unsigned long long)l.x | ((unsigned long long)h.x << 32);
                    sum[1] += (unsigned long long)l.y | ((unsigned long long)h.y << 32);
                    sum[2] += (unsigned long long)l.z | ((unsigned long long)h.z << 32);
                    sum[3] += (unsigned long long)l.w | ((unsigned long long)h.w << 32);
                } else
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gs + i >= first && gs + i < end) {
                            unsigned long long v = lo[(size_t)sl * slotStride + j + i];
                            if (f) v |= (unsigned long long)hi[(size_t)sl * slotStride + j + i] << 32;
                            sum[i] += v;
                        }
            }
        if ((sum[0] | sum[1] | sum[2] | sum[3]) == 0ull) continue;
        int cell = (int)(gs % (size_t)nRows);
        if (full) {
            const float4 j4 = *reinterpret_cast<const float4 *>(Jown + gs);
            float o[4] = {j4.x, j4.y, j4.z, j4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (sum[i] != 0ull) {
                    float len = (float)((double)(long long)sum[i] * lenUnit);
                    o[i] = o[i] + len * deltaE / dV[cell];
                }
                if (++cell == nRows) cell = 0;
            }
            const float4 w = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int r = 0; r < (N > 0 ? N : 16); ++r)
                if (r < nr) *reinterpret_cast<float4 *>(P.J[r] + gs) = w;
            st256_zero(Q + gs);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (sum[i] != 0ull) {                        // (only elements of the share can be non-zero here)
                    float len = (float)((double)(long long)sum[i] * lenUnit);
                    float v = Jown[gs + i] + len * deltaE / dV[cell];
#pragma unroll
                    for (int r = 0; r < (N > 0 ? N : 16); ++r)
                        if (r < nr) P.J[r][gs + i] = v;
                    Q[gs + i] = 0ull;
                }
                if (++cell == nRows) cell = 0;
            }
        }
    }
    __threadfence_system();
}

int p2p_pack_block() { return kPackBlk; }

cudaError_t launch_p2p_push_packed(const P2PPackedPeers &P, const unsigned long long *Q, size_t off, size_t count, size_t rOff,
                                   size_t slotStride, size_t flagStride, int blocks, cudaStream_t s)
{
    if (count == 0) return cudaSuccess;
    p2p_push_packed_kernel<<<blocks, 256, 0, s>>>(P, Q, off, count, rOff, slotStride, flagStride);
    return cudaGetLastError();
}

cudaError_t launch_p2p_sum_fold_packed(const P2PPeers &P, const unsigned int *lo, const unsigned int *hi, const unsigned char *flag,
                                       size_t slotStride, size_t flagStride, size_t rOff, const float *dV, int nRows, size_t first,
                                       size_t total, double lenUnit, float deltaE, int blocks, cudaStream_t s)
{
    if (total == 0) return cudaSuccess;
    switch (P.nranks) {
    case 2: p2p_sum_fold_packed_kernel<2><<<blocks, 256, 0, s>>>(P, lo, hi, flag, slotStride, flagStride, rOff, dV, nRows, first, total, lenUnit, deltaE); break;
    case 4: p2p_sum_fold_packed_kernel<4><<<blocks, 256, 0, s>>>(P, lo, hi, flag, slotStride, flagStride, rOff, dV, nRows, first, total, lenUnit, deltaE); break;
    case 8: p2p_sum_fold_packed_kernel<8><<<blocks, 256, 0, s>>>(P, lo, hi, flag, slotStride, flagStride, rOff, dV, nRows, first, total, lenUnit, deltaE); break;
    default: p2p_sum_fold_packed_kernel<0><<<blocks, 256, 0, s>>>(P, lo, hi, flag, slotStride, flagStride, rOff, dV, nRows, first, total, lenUnit, deltaE); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_p2p_reduce_fold(const P2PPeers &P, const float *dV, int nRows, size_t first, size_t total,
                                   double lenUnit, float deltaE, int blocks, cudaStream_t s)
{
    if (total == 0) return cudaSuccess;
    switch (P.nranks) {
    case 2: p2p_reduce_fold_kernel<2, 4><<<blocks, 256, 0, s>>>(P, dV, nRows, first, total, lenUnit, deltaE); break;
    case 4: p2p_reduce_fold_kernel<4, 2><<<blocks, 256, 0, s>>>(P, dV, nRows, first, total, lenUnit, deltaE); break;
    case 8: p2p_reduce_fold_kernel<8, 2><<<blocks, 256, 0, s>>>(P, dV, nRows, first, total, lenUnit, deltaE); break;
    default: p2p_reduce_fold_generic_kernel<<<blocks, 256, 0, s>>>(P, dV, nRows, first, total, lenUnit, deltaE); break;
    }
    return cudaGetLastError();
}

__global__ void fold_count_kernel(unsigned int *__restrict__ Q, float *__restrict__ E,
                                  size_t total, float deltaE)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        unsigned int q = Q[i];
        if (q != 0u) {
            E[i] = E[i] + (float)q * deltaE;
            Q[i] = 0u;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Sparse exchange of the escape counts (multi-rank): escapedQ is indexed by the cell a packet
// was last emitted or scattered in, so only a few per cent of its 1.26e9 entries are non-zero.
// compact: (index, count) pairs of the non-zero entries of a range, appended with one atomic
// per warp; scatter: add a list of pairs into the array (exact integer adds, any order).
// ---------------------------------------------------------------------------------------
__global__ void esc_compact_kernel(unsigned int *__restrict__ Q, size_t off, size_t len,
                                   unsigned long long *__restrict__ list, unsigned long long *count,
                                   unsigned long long capacity, int clear)
{
    // four independent 128 B loads per warp and trip (one was latency bound: 3 ms for 5 GB), one atomic per warp
    const unsigned int lane = threadIdx.x & 31u;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t base = warp * 128; base < len; base += nWarps * 128) {
        unsigned int q[4], m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t i = base + (size_t)k * 32 + lane;
            q[k] = i < len ? Q[off + i] : 0u;
        }
        unsigned int total = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { m[k] = __ballot_sync(0xffffffffu, q[k] != 0u); total += __popc(m[k]); }
        if (!total) continue;
        unsigned long long p = 0;
        if (lane == 0) p = atomicAdd(count, (unsigned long long)total);
        p = __shfl_sync(0xffffffffu, p, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (q[k] != 0u) {
                const size_t i = base + (size_t)k * 32 + lane;
                const unsigned long long at = p + __popc(m[k] & ((1u << lane) - 1u));
                if (list && at < capacity) { list[2 * at] = (unsigned long long)(off + i); list[2 * at + 1] = q[k]; }
                if (clear) Q[off + i] = 0u;
            }
            p += __popc(m[k]);
        }
    }
}

__global__ void esc_scatter_kernel(unsigned int *__restrict__ Q, size_t total, const unsigned long long *__restrict__ list,
                                   unsigned long long n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned long long idx = list[2 * i];
        unsigned int q = (unsigned int)list[2 * i + 1];
        if (q && idx < total) atomicAdd(&Q[idx], q);
    }
}

cudaError_t launch_esc_compact(unsigned int *Q, size_t off, size_t len, unsigned long long *list,
                               unsigned long long *count, unsigned long long capacity, int clear, int blocks,
                               cudaStream_t s)
{
    if (len == 0) return cudaSuccess;
    esc_compact_kernel<<<blocks, 256, 0, s>>>(Q, off, len, list, count, capacity, clear);
    return cudaGetLastError();
}

__global__ void esc_clear_list_kernel(unsigned int *__restrict__ Q, const unsigned long long *__restrict__ list, unsigned long long n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) Q[list[2 * i]] = 0u;
}

cudaError_t launch_esc_clear_list(unsigned int *Q, const unsigned long long *list, unsigned long long n, int blocks, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    esc_clear_list_kernel<<<blocks, 256, 0, s>>>(Q, list, n);
    return cudaGetLastError();
}

cudaError_t launch_esc_scatter(unsigned int *Q, size_t total, const unsigned long long *list, unsigned long long n,
                               int blocks, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    esc_scatter_kernel<<<blocks, 256, 0, s>>>(Q, total, list, n);
    return cudaGetLastError();
}

// merge the second tally set into the first (exact integer adds) and clear it
__global__ void merge_sets_kernel(unsigned long long *__restrict__ J0, unsigned long long *__restrict__ J1, size_t nJ,
                                  unsigned int *__restrict__ E0, unsigned int *__restrict__ E1, size_t nE,
                                  int *__restrict__ f0, int *__restrict__ f1, int nf)
{
    size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = i0; i < nJ; i += stride) { unsigned long long v = J1[i]; if (v) { J0[i] += v; J1[i] = 0ull; } }
    for (size_t i = i0; i < nE; i += stride) { unsigned int v = E1[i]; if (v) { E0[i] += v; E1[i] = 0u; } }
    for (size_t i = i0; i < (size_t)nf; i += stride) { if (f1[i]) { f0[i] = 1; f1[i] = 0; } }
}

// ---------------------------------------------------------------------------------------
// K7 SED reduction (writeSED, output_mod.f90:2561-2568): per (nu, viewing angle) plane, the
// number of packets that escaped, summed over the cells of origin.  Exact integer sum, so the
// 5 GB escapedPackets array never has to be exchanged or downloaded for the SED.
// grid = (slices of a plane, planes).  HBM-bound: 4 B per (cell, nu, angle) element.
// ---------------------------------------------------------------------------------------
constexpr int kSedPerThread = 16;

__global__ void __launch_bounds__(256) sed_sum_kernel(const unsigned int *__restrict__ escQ, size_t nR,
                                                      unsigned long long *__restrict__ sedQ)
{
    const unsigned int *plane = escQ + nR * (size_t)blockIdx.y;
    size_t base = (size_t)blockIdx.x * (256 * kSedPerThread);
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < kSedPerThread; ++k) {
        size_t i = base + (size_t)k * 256 + threadIdx.x;
        if (i < nR) s += plane[i];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ unsigned long long part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += part[w];
        if (t) atomicAdd(&sedQ[blockIdx.y], t);
    }
}

cudaError_t launch_sed_sum(const unsigned int *escQ, size_t nR, int firstPlane, int nPlanes,
                           unsigned long long *sedQ, cudaStream_t s)
{
    if (nPlanes <= 0) return cudaSuccess;
    dim3 grid((unsigned)((nR + 256 * kSedPerThread - 1) / (256 * kSedPerThread)), (unsigned)nPlanes);
    sed_sum_kernel<<<grid, 256, 0, s>>>(escQ + nR * (size_t)firstPlane, nR, sedQ + firstPlane);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Sublimation mask of the dust opacity terms from the device-resident dust state
// (iteration_mod.f90:189: Tdust(nS,ai,cell) < TdustSublime(dcp-1+nS)), so that K1 can rebuild
// scaOpac/absOpac/opacity after K5 (mcb200_dust_update) without Tdust leaving the device.
// Term t = (sg-1)*nSizes + (ai-1), sg = global species; one thread per (cell, term).
// ---------------------------------------------------------------------------------------
__global__ void dust_mask_kernel(const float *__restrict__ Tdust, const int *__restrict__ compOfCell,
                                 const int *__restrict__ dustComPoint, const int *__restrict__ nSpeciesPart,
                                 const float *__restrict__ Tsub, int nRows, int nSpeciesTot, int nSizes, int s0, int s1,
                                 unsigned char *__restrict__ on)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)nRows * (size_t)(nSpeciesTot * nSizes);
    if (i >= total) return;
    int cell = (int)(i % (size_t)nRows), t = (int)(i / (size_t)nRows);
    int sg = t / nSizes + 1, ai = t % nSizes + 1;
    unsigned char v = 0;
    int k = cell > 0 ? compOfCell[cell] : -1;
    if (k >= 0) {
        int nS = sg - dustComPoint[k] + 1;
        if (nS >= 1 && nS <= nSpeciesPart[k]) {
            float Td = Tdust[(size_t)nS + (size_t)s0 * ((size_t)ai + (size_t)s1 * (size_t)cell)];
            v = Td < Tsub[sg - 1] ? 1 : 0;
        }
    }
    on[i] = v;
}

cudaError_t launch_dust_mask(const float *Tdust, const int *compOfCell, const int *dustComPoint, const int *nSpeciesPart,
                             const float *Tsub, int nRows, int nSpeciesTot, int nSizes, int s0, int s1, unsigned char *on,
                             cudaStream_t s)
{
    size_t total = (size_t)nRows * (size_t)(nSpeciesTot * nSizes);
    if (!total) return cudaSuccess;
    dust_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(Tdust, compOfCell, dustComPoint, nSpeciesPart, Tsub, nRows,
                                                                     nSpeciesTot, nSizes, s0, s1, on);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Rows opacity(cell_r, 1:nbins) of a few cells out of the (0:nCells, nbins) table: what writeTauNu
// (output_mod.f90:2384-2505) reads along its three rays, without downloading the table.
// ---------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float *__restrict__ table, size_t nR, int nb, const int *__restrict__ cells,
                                   int nWanted, float *__restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nWanted * (size_t)nb) return;
    int r = (int)(i % (size_t)nWanted), nu = (int)(i / (size_t)nWanted);
    out[i] = table[(size_t)cells[r] + nR * (size_t)nu];          // out(r, nu), r fastest
}

cudaError_t launch_gather_rows(const float *table, size_t nR, int nb, const int *cells, int nWanted, float *out, cudaStream_t s)
{
    size_t total = (size_t)nWanted * (size_t)nb;
    if (!total) return cudaSuccess;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(table, nR, nb, cells, nWanted, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// K9 continuum-cube reduction (writeContCube, output_mod.f90:2762-2772): per cell and viewing
// angle, the folded escapedPackets summed over the frequency bins 1..nbins, in the reference's
// order (freq ascending, float32 running sum) so the result equals its loop bit for bit.
// One thread per cell (cell is the fast index of every (nu, angle) plane: coalesced), angle in
// blockIdx.y.  HBM-bound: each element of escapedPackets is read once, 4 B per (cell, nu, angle).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) contcube_kernel(const float *__restrict__ esc, size_t nR, int nb,
                                                       float *__restrict__ out)
{
    size_t cell = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (cell >= nR) return;
    const float *p = esc + nR * (size_t)(nb + 1) * blockIdx.y + cell;     // plane nu = 0 of this angle
    float s = 0.f;
    for (int nu = 1; nu <= nb; ++nu) s = s + __ldg(p + nR * (size_t)nu);
    out[nR * blockIdx.y + cell] = s;
}

cudaError_t launch_contcube(const float *esc, size_t nR, int nb, int nAngles, float *out, cudaStream_t s)
{
    dim3 grid((unsigned)((nR + 255) / 256), (unsigned)nAngles);
    contcube_kernel<<<grid, 256, 0, s>>>(esc, nR, nb, out);
    return cudaGetLastError();
}

cudaError_t launch_merge_sets(unsigned long long *J0, unsigned long long *J1, size_t nJ, unsigned int *E0,
                              unsigned int *E1, size_t nE, int *f0, int *f1, int nf, int blocks, cudaStream_t s)
{
    merge_sets_kernel<<<blocks, 256, 0, s>>>(J0, J1, nJ, E0, E1, nE, f0, f1, nf);
    return cudaGetLastError();
}

cudaError_t launch_fold_j(unsigned long long *Q, float *J, const float *dV, int nRows, size_t first, size_t total,
                          double lenUnit, float deltaE, int blocks, cudaStream_t s)
{
    if (total == 0) return cudaSuccess;
    fold_j_kernel<<<blocks, 256, 0, s>>>(Q, J, dV, nRows, first, total, lenUnit, deltaE);
    return cudaGetLastError();
}

cudaError_t launch_fold_count(unsigned int *Q, float *E, size_t total, float deltaE, int blocks,
                              cudaStream_t s)
{
    fold_count_kernel<<<blocks, 256, 0, s>>>(Q, E, total, deltaE);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// K1: opacity assembly (ionization_mod.f90:349-484 + iteration_mod.f90:166-227).
// One CTA per tile of kTile cells, all frequencies; the tile's species densities are
// staged once in shared memory ([species][cell], conflict-free), cross-sections are
// warp-uniform loads (same address for every lane, served by L1), the output rows are
// written as coalesced kTile*4-byte runs.  Bands covering a frequency are visited in the
// reference's order (CSR over nu built on the host) with separate mul and add
// (-fmad=false) so the sums are bit-identical to the reference's sequential loop.
// Bound: HBM writes, (nCells+1)*nbins*4 B per output table.
// ---------------------------------------------------------------------------------------
struct OpacityArgs {
    int nRows, nb;                        // nCells+1, nbins
    int nSpeciesDen;
    const float *den;                     // (0:nCells, nSpeciesDen) cell fastest
    const float *ff1;                     // (0:nCells) or NULL
    const float *xSec;                    // xSecArray, 1-based offsets
    const int *nuStart;                   // [nb+1] CSR offsets into nuBand*
    const int *nuBandSpecies;             // species column (0-based) per entry
    const int *nuBandXs;                  // 1-based xSecArray index for this (band, nu)
    // dust
    int nDustTerms;                       // flattened (species,size) terms for this call
    const float *Ndust;                   // (0:nCells)
    const unsigned char *dustOn;          // [term][cell]: Tdust(s,a,c) < TdustSublime(s)
    const float *dustCoef;                // [comp][term]: grainAbun(comp,s)*grainWeight(a)
    const int *dustCompOfCell;            // (0:nCells) 0-based component, or NULL (=0)
    const int *dustTermOn;                // [comp][term]: term belongs to component
    const int *dustScaP, *dustAbsP;       // [term] 1-based xSecArray offsets
    float *opacity, *scaOpac, *absOpac;   // outputs (0:nCells, nbins)
};

constexpr int kTile = 128;

__global__ void __launch_bounds__(kTile) opacity_kernel(const OpacityArgs A)
{
    extern __shared__ float sden[];       // [nSpeciesDen][kTile]
    int cell = blockIdx.x * kTile + threadIdx.x;
    bool ok = cell < A.nRows && cell >= 1;
    for (int s = 0; s < A.nSpeciesDen; ++s)
        sden[s * kTile + threadIdx.x] = ok ? A.den[(size_t)s * A.nRows + cell] : 0.f;
    float nd = 0.f;
    int comp = 0;
    if (A.nDustTerms > 0 && ok) {
        nd = A.Ndust[cell];
        if (A.dustCompOfCell) comp = A.dustCompOfCell[cell];
    }
    float ff = (A.ff1 && ok) ? A.ff1[cell] : 0.f;
    __syncthreads();
    if (cell >= A.nRows) return;
    for (int nu = 0; nu < A.nb; ++nu) {
        float op = 0.f;
        if (ok) {
            if (nu == 0) op = op + ff;
            int b0 = A.nuStart[nu], b1 = A.nuStart[nu + 1];
            for (int b = b0; b < b1; ++b) {
                float d = sden[A.nuBandSpecies[b] * kTile + threadIdx.x];
                if (d > 0.f) op = op + __ldg(&A.xSec[A.nuBandXs[b] - 1]) * d;
            }
        }
        size_t o = (size_t)nu * A.nRows + cell;
        if (A.nDustTerms > 0) {
            float sca = 0.f, ab = 0.f;
            if (ok && comp >= 0) {
                for (int t = 0; t < A.nDustTerms; ++t) {
                    if (!A.dustTermOn[comp * A.nDustTerms + t]) continue;
                    if (!A.dustOn[(size_t)t * A.nRows + cell]) continue;
                    float coef = A.dustCoef[comp * A.nDustTerms + t] * nd;
                    sca = sca + coef * __ldg(&A.xSec[A.dustScaP[t] + nu - 1]);
                    ab = ab + coef * __ldg(&A.xSec[A.dustAbsP[t] + nu - 1]);
                }
                op = op + (sca + ab);
            }
            A.scaOpac[o] = sca;
            A.absOpac[o] = ab;
        }
        A.opacity[o] = op;
    }
}

cudaError_t launch_opacity(const OpacityArgs &A, cudaStream_t s)
{
    int blocks = (A.nRows + kTile - 1) / kTile;
    size_t smem = (size_t)A.nSpeciesDen * kTile * sizeof(float);
    if (smem < 4) smem = 4;
    cudaFuncSetAttribute(opacity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    opacity_kernel<<<blocks, kTile, smem, s>>>(A);
    return cudaGetLastError();
}

// out(1:nMine, 1:nb) = table(first + (j-1)*stride, :) : the rows of the cells one rank owns under the
// reference's round-robin rule (iteration_mod.f90:832), compacted so that only they cross PCIe
__global__ void __launch_bounds__(256) gather_cells_kernel(const float *__restrict__ table, size_t nR, int first, int stride,
                                                           int nMine, float *__restrict__ out)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nMine) return;
    size_t nu = blockIdx.y;
    out[(size_t)j + (size_t)nMine * nu] = table[(size_t)first + (size_t)j * (size_t)stride + nR * nu];
}

cudaError_t launch_gather_cells(const float *table, size_t nR, int nb, int first, int stride, int nMine, float *out, cudaStream_t s)
{
    if (nMine <= 0 || nb <= 0) return cudaSuccess;
    dim3 grid((unsigned)((nMine + 255) / 256), (unsigned)nb);
    gather_cells_kernel<<<grid, 256, 0, s>>>(table, nR, first, stride, nMine, out);
    return cudaGetLastError();
}

}  // namespace mcb
