// transport.cu -- K2+K3: energy-packet emission and transport for sm_100a.
//
// Replaces energyPacketDriver / energyPacketRun / newPhotonPacket / initPhotonPacket /
// getNu2 / pathSegment / hg of the reference (source/photon_mod.f90:26-2974) plus
// randomUnitVector (vector_mod.f90:303-314) and locate (interpolation_mod.f90:48-81).
//
// Design (B200-first, not a translation):
//  * one energy packet per thread, persistent CTAs; a lane that finishes its packet
//    immediately pulls the next global packet index (warp-aggregated atomic), so the
//    32 lanes of a warp stay busy although packets live for 1..5000 generations;
//  * the three nested loops of the reference (packets / generations / cell crossings)
//    are flattened into one warp-level loop over a per-lane phase (NEED, EMIT, FLY,
//    SCATTER, ESCAPE): every trip executes exactly one cell crossing for the flying
//    lanes; the rare, expensive phases (emission = Philox + CDF search + sincos + log,
//    scattering = Henyey-Greenstein, escape = acos/atan binning) are *deferred* until
//    kBatch lanes of the warp wait for the same phase (or nothing is left to fly), so
//    they run with many active lanes instead of one (ncu: 1.2 -> see profiles/);
//  * counter-based Philox4x32-10 stream per packet (philox.cuh): results do not depend
//    on the thread/CTA/GPU that runs a packet;
//  * tallies are order-independent 64-bit integer reductions (RED.E.ADD.64 at L2):
//    path length in fixed point for Jste/Jdif, packet counts for escapedPackets /
//    linePackets; the float32 estimators of the reference are produced by the fold
//    epilogue (tables.cu) -> bitwise reproducible for a fixed seed on any GPU count;
//  * the inactive-cell sink row 0 of Jste (photon_mod.f90:1823 with active=0) is never
//    read by the reference and would be the hottest atomic address: skipped;
//  * mid-point cell walls come from precomputed per-axis tables (same float32
//    expression as photon_mod.f90:1266), the cell volume is not needed in the loop at
//    all (deltaE/dV is applied once per cell in the fold);
//  * re-emission CDF rows are stored transposed [cell][nu] and binary searched; the
//    result equals getNu2's linear scan for non-decreasing tables (checked on upload);
//  * float32 arithmetic with -fmad=false and the detmath.cuh transcendentals so that
//    every branch decision is bit-identical to the CPU oracle.
#include "transport_core.cuh"
#include "wf_rec.cuh"

namespace mcb {

// Work source of the persistent kernel: fresh packets 0..n-1, or (resume != NULL) the packets
// a wave-front run left alive, taken from its four event lists -- used to finish the long,
// thin tail of a batch (a few packets with thousands of generations) without paying a
// kernel-launch round trip per wave.
template <bool MULTI>
__device__ __forceinline__ void claim_packet(const TransportArgs &a, const WfArgs *resume, Transport<MULTI> &T,
                                             Lane &L, long long k)
{
    if (!resume) {
        T.start_packet(L, a.order ? (long long)__ldg(&a.order[k]) : k);
        return;
    }
    unsigned int j = (unsigned int)k;
    int ev = 0;
    for (; ev < EV_COUNT - 1; ++ev) {
        unsigned int c = resume->evCount[ev];
        if (j < c) break;
        j -= c;
    }
    rec_load<MULTI>(a, resume->recA, resume->recxA, L, resume->evList[ev][j]);
    L.phase = ev == EV_EMIT ? PH_EMIT : ev == EV_SCATTER ? PH_SCATTER : ev == EV_ESCAPE ? PH_ESCAPE : PH_FLY;
}

template <bool MULTI>
__global__ void __launch_bounds__(kThreads)
transport_kernel(const __grid_constant__ TransportArgs a, const WfArgs *resume)
{
    extern __shared__ unsigned int smem[];
    unsigned int *cnt = smem;                        // [C_COUNT][kThreads]
    unsigned int *qph = smem + C_COUNT * kThreads;   // [nbins]
    scratch_init(smem, a.P.nbins);

    Transport<MULTI> T(a, cnt, qph);
    Lane L;
    L.phase = PH_NEED;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int FULL = 0xffffffffu;

    for (;;) {
        // refill: lanes without a packet claim the next global packet index
        unsigned int need = __ballot_sync(FULL, L.phase == PH_NEED);
        if (need) {
            unsigned long long base = 0;
            int leader = __ffs(need) - 1;
            if ((int)lane == leader) base = atomicAdd(a.nextPacket, (unsigned long long)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (L.phase == PH_NEED) {
                long long k = (long long)(base + __popc(need & ((1u << lane) - 1u)));
                if (k < a.n) claim_packet<MULTI>(a, resume, T, L, k);
                else L.phase = PH_DONE;
            }
        }
        unsigned int fly = __ballot_sync(FULL, L.phase == PH_FLY);
        unsigned int pe = __ballot_sync(FULL, L.phase == PH_EMIT);
        unsigned int ps = __ballot_sync(FULL, L.phase == PH_SCATTER);
        unsigned int px = __ballot_sync(FULL, L.phase == PH_ESCAPE);
        if ((fly | pe | ps | px) == 0u) break;       // every lane is DONE
        // deferred rare phases: run when enough lanes wait, or nothing is left to fly
        bool flush = (fly == 0u) || (__popc(pe | ps | px) >= 2 * a.batch);
        if (px && (flush || __popc(px) >= a.batch)) { if (L.phase == PH_ESCAPE) T.do_escape(L); }
        if (ps && (flush || __popc(ps) >= a.batch)) { if (L.phase == PH_SCATTER) T.do_scatter(L); }
        if (pe && (flush || __popc(pe) >= a.batch)) { if (L.phase == PH_EMIT) T.do_emit(L); }
        // pre-reduce the tallies inside the warp while several lanes are in their first
        // cells (frequency-ordered processing: they share (cell, nu))
        bool agg = a.aggSteps > 0 &&
                   __popc(__ballot_sync(FULL, L.phase == PH_FLY && L.istep < a.aggSteps)) >= 4;
        if (L.phase == PH_FLY) T.step(L, agg);
    }

    scratch_flush(a, smem);
}

// ---------------------------------------------------------------------------------------
// Frequency ordering.  The first emission of packet k is a pure function of its Philox
// stream, so its frequency bin can be computed ahead of the transport kernel
// (first_nu_kernel replays exactly the draws do_emit will make for the first getNu2),
// and the packet indices counting-sorted by that bin.  Lanes then pull packets in
// frequency order: at any time the ~10^5 resident packets touch one or two nu-planes
// of opacity/Jste (8.4 + 16.8 MB at 128^3) that stay in the 126 MB L2 instead of
// streaming random 32 B sectors from HBM.  Results are unchanged (order independent).
// ---------------------------------------------------------------------------------------
__global__ void first_nu_kernel(const TransportArgs a, unsigned short *key, unsigned int *hist)
{
    extern __shared__ unsigned int sh[];             // [nbins+1]
    const int nb = a.P.nbins;
    for (int i = threadIdx.x; i <= nb; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const float *cdf = a.P.starCdf + (size_t)(a.iStar >= 1 ? a.iStar : 0) * nb;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < a.n; k += stride) {
        Rng rng;
        rng.init(a.seed, a.pidBase + (unsigned long long)(a.firstId + k), a.rngStream);
        int nuP = sample_cdf(rng, cdf, nb);
        if (nuP > nb) nuP = nb;
        key[k] = (unsigned short)nuP;
        atomicAdd(&sh[nuP], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= nb; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// exclusive scan of hist[0..nb] -> cursor[0..nb] (single block; nb <= a few thousand)
__global__ void scan_hist_kernel(const unsigned int *hist, unsigned int *cursor, int nb)
{
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int i = 0; i <= nb; ++i) { cursor[i] = run; run += hist[i]; }
    }
}

__global__ void scatter_order_kernel(const unsigned short *key, unsigned int *cursor, unsigned int *order,
                                     long long n, int nb)
{
    // each block owns a contiguous chunk; bins are reserved per block with one global
    // atomic per non-empty bin, positions inside the reservation come from shared atomics
    extern __shared__ unsigned int sh[];             // [2*(nb+1)]: counts, bases
    unsigned int *cntb = sh, *base = sh + (nb + 1);
    const long long chunk = 8192;
    for (long long c0 = (long long)blockIdx.x * chunk; c0 < n; c0 += (long long)gridDim.x * chunk) {
        long long c1 = c0 + chunk < n ? c0 + chunk : n;
        for (int i = threadIdx.x; i <= nb; i += blockDim.x) cntb[i] = 0u;
        __syncthreads();
        for (long long k = c0 + threadIdx.x; k < c1; k += blockDim.x) atomicAdd(&cntb[key[k]], 1u);
        __syncthreads();
        for (int i = threadIdx.x; i <= nb; i += blockDim.x) {
            base[i] = cntb[i] ? atomicAdd(&cursor[i], cntb[i]) : 0u;
            cntb[i] = 0u;
        }
        __syncthreads();
        for (long long k = c0 + threadIdx.x; k < c1; k += blockDim.x) {
            unsigned int b = key[k];
            order[base[b] + atomicAdd(&cntb[b], 1u)] = (unsigned int)k;
        }
        __syncthreads();
    }
}

cudaError_t launch_order(const TransportArgs &a, unsigned short *key, unsigned int *hist, unsigned int *cursor,
                         unsigned int *order, int numSMs, cudaStream_t stream)
{
    const int nb = a.P.nbins;
    cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(unsigned int) * (nb + 1), stream);
    if (e != cudaSuccess) return e;
    size_t sm1 = sizeof(unsigned int) * (nb + 1), sm2 = 2 * sm1;
    first_nu_kernel<<<numSMs * 8, 256, sm1, stream>>>(a, key, hist);
    scan_hist_kernel<<<1, 32, 0, stream>>>(hist, cursor, nb);
    scatter_order_kernel<<<numSMs * 4, 256, sm2, stream>>>(key, cursor, order, a.n, nb);
    return cudaGetLastError();
}

cudaError_t launch_transport(const TransportArgs &a, bool multi, int gridBlocks, cudaStream_t stream, const WfArgs *resume)
{
    size_t smem = (size_t)scratch_words(a.P.nbins) * sizeof(unsigned int);
    if (multi) {
        cudaFuncSetAttribute(transport_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        transport_kernel<true><<<gridBlocks, kThreads, smem, stream>>>(a, resume);
    } else {
        cudaFuncSetAttribute(transport_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        transport_kernel<false><<<gridBlocks, kThreads, smem, stream>>>(a, resume);
    }
    return cudaGetLastError();
}

int transport_blocks_per_sm(bool multi)
{
    int nb = 0;
    size_t smem = (size_t)scratch_words(1024) * sizeof(unsigned int);
    if (multi) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, transport_kernel<true>, kThreads, smem);
    else       cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, transport_kernel<false>, kThreads, smem);
    return nb;
}

}  // namespace mcb
