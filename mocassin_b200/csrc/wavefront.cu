// wavefront.cu -- wave-front variant of the transport (K2+K3) for large batches.
//
// Same physics and the same Transport<MULTI> phase functions as the persistent kernel
// (transport.cu), different schedule.  A packet history is a chain of *flights* (emission or
// scattering -> cell crossings -> next interaction / escape).  Instead of letting one
// thread carry a packet through all its phases, each wave runs every phase as its own
// fully converged kernel over compact lists:
//
//   wf_event_kernel<EMIT>     newPhotonPacket + head of pathSegment for every packet that
//                             was absorbed in the previous wave (wave 0: every packet)
//   wf_event_kernel<SCATTER>  Henyey-Greenstein / isotropic re-direction
//   wf_event_kernel<ESCAPE>   escape binning + tally
//   wf_sort_*                 counting sort of the ready flights by frequency bin
//   wf_fly_kernel             the hot loop: persistent lanes pull flights in frequency
//                             order and cross cells until the next event
//
// Why (ncu, profiles/): in the persistent kernel the rare phases ran with 1-8 of 32 lanes
// and lanes idled while waiting for them (19/32 active threads per instruction); here the
// event kernels run 32/32 and the FLY kernel contains only the cell-crossing code (small
// I-cache footprint, fewer registers, higher occupancy).  Re-sorting the flights by nu in
// EVERY wave keeps the opacity/Jste planes touched at any moment inside the 126 MB L2 for
// all generations, not only the first.  Packet state lives in one 64 B record per packet
// (two 32 B sectors per flight boundary).  Results are bit-identical to the persistent
// kernel: tallies are order-independent integers and every packet owns its Philox stream.
#include "transport_core.cuh"
#include "wf_rec.cuh"

namespace mcb {

// append the lanes with `pred` to a list (one atomic per warp)
__device__ __forceinline__ unsigned int warp_append(bool pred, unsigned int *count)
{
    const unsigned int FULL = 0xffffffffu;
    unsigned int m = __ballot_sync(FULL, pred);
    if (!m) return 0;
    unsigned int lane = threadIdx.x & 31u;
    int leader = __ffs(m) - 1;
    unsigned int base = 0;
    if ((int)lane == leader) base = atomicAdd(count, (unsigned int)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

// ---- event kernels: one thread per list entry, all lanes run the same phase ------------
template <bool MULTI, int EV>
__global__ void __launch_bounds__(kThreads) wf_event_kernel(const __grid_constant__ WfArgs w)
{
    extern __shared__ unsigned int smem[];
    scratch_init(smem, w.t.P.nbins);
    Transport<MULTI> T(w.t, smem, smem + C_COUNT * kThreads);
    const unsigned int total = w.inList ? *w.inCount : (unsigned int)w.t.n;
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int rounds = (total + stride - 1) / stride;
    for (unsigned int it = 0; it < rounds; ++it) {
        unsigned int i = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        bool valid = i < total;
        Lane L;
        L.phase = PH_DONE;
        if (EV == EV_CONT) {                          // flight continues: move its record on
            unsigned int pos = warp_append(valid, w.flyCount);
            if (valid) {
                unsigned int src = w.inList[i];
                const uint4 *a4 = reinterpret_cast<const uint4 *>(&w.recA[src]);
                uint4 v0 = a4[0], v1 = a4[1], v2 = a4[2], v3 = a4[3];
                uint4 *b4 = reinterpret_cast<uint4 *>(&w.recB[pos]);
                b4[0] = v0; b4[1] = v1; b4[2] = v2; b4[3] = v3;
                if (MULTI) *reinterpret_cast<uint4 *>(&w.recxB[pos]) = *reinterpret_cast<const uint4 *>(&w.recxA[src]);
                w.flyKey[pos] = w.recA[src].nuP;
            }
            continue;
        }
        if (valid) {
            if (EV == EV_EMIT && !w.inList) {
                // wave 0: packet i, or the i-th packet in order of its first frequency bin
                T.start_packet(L, w.t.order ? (long long)__ldg(&w.t.order[i]) : (long long)i);
            } else {
                rec_load<MULTI>(w.t, w.recA, w.recxA, L, w.inList[i]);
            }
            if (EV == EV_EMIT) {
                T.do_emit(L);
                if (L.phase == PH_ESCAPE) T.do_escape(L);     // packets below the ionisation edge
            } else if (EV == EV_SCATTER) {
                T.do_scatter(L);
            } else {
                T.do_escape(L, true);
            }
        }
        bool toFly = valid && L.phase == PH_FLY;
        if (EV == EV_EMIT && w.directA == 2) {
            // pre-ordered wave 0, exact variant: packet i of the frequency order goes to slot i of
            // the FLY array.  A packet that does not fly (escaped at emission, line packet, stop
            // condition) leaves a hole (nuP = 0) that the FLY kernel skips; the host sets
            // flyCount = n.  Measured slower than the appended variant below: with an exactly
            // sorted array every resident flight starts in the same frequency plane at the star
            // and the first crossings' reductions collide on a handful of addresses.
            if (valid) {
                if (toFly) {
                    L.vx = L.dx; L.vy = L.dy; L.vz = L.dz; L.absTau = 0.f;
                    rec_store<MULTI>(w.recA, w.recxA, L, i);
                } else {
                    reinterpret_cast<uint4 *>(&w.recA[i])[3] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            continue;
        }
        unsigned int pos = warp_append(toFly, w.flyCount);
        if (toFly) {
            L.vx = L.dx; L.vy = L.dy; L.vz = L.dz; L.absTau = 0.f;     // fresh flight
            if (EV == EV_EMIT && w.directA) {
                // pre-ordered wave 0: arrival order follows the frequency order to within one
                // grid-stride window (the launcher keeps the grid small: < 1 bin), so the record
                // goes straight to the FLY array and the counting sort is skipped
                rec_store<MULTI>(w.recA, w.recxA, L, pos);
            } else {
                rec_store<MULTI>(w.recB, w.recxB, L, pos);
                w.flyKey[pos] = (unsigned short)L.nuP;
            }
        }
    }
    scratch_flush(w.t, smem);
}

// ---- FLY: the hot loop ------------------------------------------------------------------
// Lanes claim flights in frequency order in chunks of kChunk consecutive records (one global
// atomic per chunk, records of the chunk prefetched into L2) and cross cells until the next
// event; ended flights are written back in place and their positions staged per warp in
// shared memory, flushed 32 at a time (one global atomic per 32 events, coalesced stores).
#ifndef MCB_ESC_COMPACT
#define MCB_ESC_COMPACT 1                // A/B build switch: compact escape entries (see wf_escape_compact_kernel)
#endif
#ifndef MCB_FLY_OCC
#define MCB_FLY_OCC 4                    // resident CTAs per SM the FLY kernel is compiled for (A/B: build.py --define)
#endif
#ifndef MCB_FLY_OCC_MULTI
#define MCB_FLY_OCC_MULTI 3              // multi-grid variant: more packet state (mother / sub-grid slots), 85 registers
#endif
template <bool MULTI, bool DENSE, int MODE>
__global__ void __launch_bounds__(kThreads, MULTI ? MCB_FLY_OCC_MULTI : MCB_FLY_OCC) wf_fly_kernel(const __grid_constant__ WfArgs w)
{
    extern __shared__ unsigned int smem[];
    scratch_init(smem, w.t.P.nbins);
    // single grid: the per-crossing iteration-limit test is folded into the step budget below
    Transport<MULTI, DENSE, MODE, MULTI, (MCB_LEANRNG != 0)> T(w.t, smem, smem + C_COUNT * kThreads);
    unsigned int *stage = smem + scratch_words(w.t.P.nbins) + (threadIdx.x >> 5) * (EV_COUNT * kStage);
    const unsigned int FULL = 0xffffffffu;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int total = *w.flyCount;
    unsigned int cur = 0, end = 0;                   // this warp's chunk [cur, end)
    bool exhausted = false;
    int nStaged[EV_COUNT] = {0, 0, 0, 0};
    int budget = 0;
    Lane L;
    L.phase = PH_NEED;
    unsigned int pos = 0;                            // this lane's flight: position in recA
    // Lanes whose flight ended (or that have none yet) wait until `batch` of the warp's lanes do,
    // then the warp writes the records back, stages the events and claims new flights in one
    // pass: that code runs for a few lanes at a time, so running it on every trip of the
    // crossing loop (41 % of the trips end a flight in some lane) cost more warp instructions
    // than the crossings themselves.
    const unsigned int batch = (unsigned int)(w.flyBatch < 1 ? 1 : w.flyBatch);
    unsigned int doneM = 0u;                         // lanes that found the work list empty
    for (;;) {
        unsigned int flyM = __ballot_sync(FULL, L.phase == PH_FLY);
        const unsigned int waitM = ~(flyM | doneM);
        if (waitM && ((unsigned int)__popc(waitM) >= batch || flyM == 0u)) {
            const bool ended = L.phase == PH_EMIT || L.phase == PH_SCATTER || L.phase == PH_ESCAPE || L.phase == PH_CONT;
#if MCB_ESC_COMPACT
            unsigned int escVal = 0u;
#endif
            // flights that ended: write the record back and stage its position
            if (__ballot_sync(FULL, ended)) {
                if (ended) {
                    unsigned int k;
#if MCB_ESC_COMPACT
                    if (!MULTI && w.escCompact && L.phase == PH_ESCAPE) {
                        // an escaping packet is finished: nothing of its record is needed again but the
                        // element of the escape tally it counts into, which goes on the event list in
                        // place of the record's position -- no write-back, and the ESCAPE pass of the
                        // next wave reads 4 contiguous bytes per packet instead of a 64-byte record
                        const uint4 *src = reinterpret_cast<const uint4 *>(&w.recA[pos]);
                        const uint4 c2 = src[2], c3 = src[3];
                        k = c2.z;
                        escVal = c2.w + (unsigned int)(w.t.g1.nCells + 1) * (c3.x & 0xffffu);
                        if ((c2.y >> 19) >= (unsigned int)kRecursionLimit) escVal |= 0x80000000u;
                    } else
#endif
                    rec_store_fly<MULTI>(w.recA, w.recxA, L, pos, k);
                    smem[C_SEGMENTS * kThreads + threadIdx.x] += L.segs;   // this flight's cell crossings
                    if (w.t.segsArr) w.t.segsArr[k] += L.segs;
                    L.segs = 0;
                }
#pragma unroll
                for (int ev = 0; ev < EV_COUNT; ++ev) {
                    const int ph = ev == EV_EMIT ? PH_EMIT : ev == EV_SCATTER ? PH_SCATTER : ev == EV_ESCAPE ? PH_ESCAPE : PH_CONT;
                    unsigned int m = __ballot_sync(FULL, L.phase == ph);
                    if (!m) continue;
#if MCB_ESC_COMPACT
                    if (L.phase == ph) stage[ev * kStage + nStaged[ev] + __popc(m & ((1u << lane) - 1u))] =
                        (ev == EV_ESCAPE && !MULTI && w.escCompact) ? escVal : pos;
#else
                    if (L.phase == ph) stage[ev * kStage + nStaged[ev] + __popc(m & ((1u << lane) - 1u))] = pos;
#endif
                    nStaged[ev] += __popc(m);
                    __syncwarp();
                    if (nStaged[ev] >= 32) {             // flush 32 staged positions
                        unsigned int base = 0;
                        if (lane == 0) base = atomicAdd(&w.evCount[ev], 32u);
                        base = __shfl_sync(FULL, base, 0);
                        w.evList[ev][base + lane] = stage[ev * kStage + lane];
                        unsigned int rest = lane + 32 < (unsigned int)nStaged[ev] ? stage[ev * kStage + 32 + lane] : 0u;
                        __syncwarp();
                        stage[ev * kStage + lane] = rest;
                        nStaged[ev] -= 32;
                        __syncwarp();
                    }
                }
                if (ended) L.phase = PH_NEED;
            }
            // claim new flights for every idle lane
            unsigned int need = __ballot_sync(FULL, L.phase == PH_NEED);
            while (need && !exhausted) {
                if (cur >= end) {                        // claim the next chunk
                    unsigned int base = 0;
                    if (lane == 0) base = (unsigned int)atomicAdd(w.nextFlight, (unsigned long long)kChunk);
                    base = __shfl_sync(FULL, base, 0);
                    if (base >= total) { exhausted = true; break; }
                    cur = base;
                    end = base + kChunk < total ? base + kChunk : total;
                    // pull the chunk's records towards L2 while the current flights finish
                    for (unsigned int p = cur + lane * 2; p < end; p += 64)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(&w.recA[p]));
                }
                unsigned int rank = __popc(need & ((1u << lane) - 1u));
                unsigned int avail = end - cur;
                bool take = (L.phase == PH_NEED) && rank < avail;
                if (take) {
                    pos = cur + rank;
                    rec_load<MULTI>(w.t, w.recA, w.recxA, L, pos);
                    L.phase = L.nuP ? PH_FLY : PH_NEED;      // nuP = 0: hole left by the pre-ordered wave 0
                    budget = w.stepBudget;
                    if (!MULTI) {                            // never run past the iteration limit (:2838-2846)
                        int left = w.t.P.safeLimit - L.istep;
                        budget = left < budget ? (left < 1 ? 1 : left) : budget;
                    }
                }
                unsigned int took = (unsigned int)__popc(need) < avail ? (unsigned int)__popc(need) : avail;
                cur += took;
                need = __ballot_sync(FULL, L.phase == PH_NEED);
            }
            if (exhausted && L.phase == PH_NEED) L.phase = PH_DONE;
            doneM = __ballot_sync(FULL, L.phase == PH_DONE);
            flyM = ~doneM;                           // every other lane holds a flight now
        }
        if (flyM == 0u) break;
        if (L.phase == PH_FLY) {
            T.step(L, false);                        // warp-aggregated tallies measured slower here too
            if (L.phase == PH_FLY && --budget <= 0) {
                // out of budget: continue in the next wave -- unless the budget was the iteration
                // limit itself, where the reference drops the packet
                if (!MULTI && L.istep >= w.t.P.safeLimit) T.finish(L, FATE_DROPPED);
                else L.phase = PH_CONT;
            }
        }
    }
#pragma unroll
    for (int ev = 0; ev < EV_COUNT; ++ev) {          // flush what is left in the staging buffers
        if (nStaged[ev] > 0) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&w.evCount[ev], (unsigned int)nStaged[ev]);
            base = __shfl_sync(FULL, base, 0);
            for (int i = lane; i < nStaged[ev]; i += 32) w.evList[ev][base + i] = stage[ev * kStage + i];
        }
    }
    scratch_flush(w.t, smem);
}

// ---- ESCAPE pass on compact entries (single grid, no viewing angles, no trace): entry = index of
// the escape-tally element (origin cell, nu) the packet counts into, bit 31 = the packet had reached
// the generation limit (energyPacketDriver counts it as trapped too).  Entries arrive in frequency
// order and most stellar packets share the star's cell: equal targets are counted once per warp.
__global__ void __launch_bounds__(256) wf_escape_compact_kernel(const __grid_constant__ WfArgs w)
{
    const unsigned int n = *w.inCount;
    unsigned int nEsc = 0, nTrap = 0;
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int rounds = (n + stride - 1) / stride;
    for (unsigned int it = 0; it < rounds; ++it) {
        unsigned int i = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = i < n;
        const unsigned int e = valid ? w.inList[i] : 0xffffffffu;
        const unsigned int peers = __match_any_sync(0xffffffffu, e & 0x7fffffffu);
        if (valid) {
            if ((int)(threadIdx.x & 31u) == __ffs(peers) - 1) atomicAdd(&w.t.g1.escQ[e & 0x7fffffffu], (unsigned int)__popc(peers));
            ++nEsc;
            nTrap += e >> 31;
        }
    }
    for (int d = 16; d > 0; d >>= 1) { nEsc += __shfl_xor_sync(0xffffffffu, nEsc, d); nTrap += __shfl_xor_sync(0xffffffffu, nTrap, d); }
    if ((threadIdx.x & 31u) == 0u) {
        if (nEsc) atomicAdd(&w.t.counters[C_ESCAPED], (unsigned long long)nEsc);
        if (nTrap) atomicAdd(&w.t.counters[C_TRAPPED], (unsigned long long)nTrap);
    }
}

bool wf_esc_compact_built() { return MCB_ESC_COMPACT != 0; }

cudaError_t wf_launch_escape_compact(const WfArgs &w, int blocks, cudaStream_t s)
{
    wf_escape_compact_kernel<<<blocks, 256, 0, s>>>(w);
    return cudaGetLastError();
}

// ---- counting sort of the ready flights by frequency bin: recB -> recA -------------------
__global__ void wf_hist_kernel(const unsigned short *key, const unsigned int *count, unsigned int *hist, int nb)
{
    extern __shared__ unsigned int sh[];
    for (int i = threadIdx.x; i <= nb; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const unsigned int n = *count;
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
        unsigned int i = i0 + threadIdx.x;
        bool valid = i < n;
        unsigned int b = valid ? key[i] : 0u;
        unsigned int peers = __match_any_sync(0xffffffffu, valid ? b : 0xffffffffu);
        if (valid && (int)(threadIdx.x & 31u) == __ffs(peers) - 1) atomicAdd(&sh[b], (unsigned int)__popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= nb; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

__global__ void wf_scan_kernel(unsigned int *hist, unsigned int *cursor, int nb)
{
    // nb <= a few thousand: one warp, shuffle scan in tiles of 32
    unsigned int lane = threadIdx.x, run = 0;
    for (int i0 = 0; i0 <= nb; i0 += 32) {
        int i = i0 + lane;
        unsigned int v = i <= nb ? hist[i] : 0u, x = v;
        for (int d = 1; d < 32; d <<= 1) { unsigned int y = __shfl_up_sync(0xffffffffu, x, d); if ((int)lane >= d) x += y; }
        if (i <= nb) { cursor[i] = run + x - v; hist[i] = 0u; }
        run += __shfl_sync(0xffffffffu, x, 31);
    }
}

// Keys arrive clustered (the event lists follow the previous wave's frequency order), so
// plain shared-memory atomics would serialise on a handful of bins: lanes with equal keys
// are aggregated first (MATCH.ANY), one shared atomic per distinct key per warp.
__device__ __forceinline__ unsigned int smem_count_aggregated(unsigned int *cntb, unsigned int b, bool valid)
{
    const unsigned int FULL = 0xffffffffu;
    unsigned int lane = threadIdx.x & 31u;
    unsigned int peers = __match_any_sync(FULL, valid ? b : 0xffffffffu);
    unsigned int rank = __popc(peers & ((1u << lane) - 1u));
    int leader = __ffs(peers) - 1;
    unsigned int base = 0;
    if (valid && (int)lane == leader) base = atomicAdd(&cntb[b], (unsigned int)__popc(peers));
    base = __shfl_sync(FULL, base, leader);
    return base + rank;
}

template <bool MULTI>
__global__ void wf_scatter_kernel(const WfArgs w, int nb)
{
    extern __shared__ unsigned int sh[];                 // [2*(nb+1)]
    unsigned int *cntb = sh, *base = sh + (nb + 1);
    const unsigned int n = *w.flyCount;
    const unsigned int chunk = 4096;
    for (unsigned int c0 = blockIdx.x * chunk; c0 < n; c0 += gridDim.x * chunk) {
        unsigned int c1 = c0 + chunk < n ? c0 + chunk : n;
        for (int i = threadIdx.x; i <= nb; i += blockDim.x) cntb[i] = 0u;
        __syncthreads();
        for (unsigned int i0 = c0; i0 < c1; i0 += blockDim.x) {
            unsigned int i = i0 + threadIdx.x;
            bool valid = i < c1;
            smem_count_aggregated(cntb, valid ? w.flyKey[i] : 0u, valid);
        }
        __syncthreads();
        for (int i = threadIdx.x; i <= nb; i += blockDim.x) {
            base[i] = cntb[i] ? atomicAdd(&w.cursor[i], cntb[i]) : 0u;
            cntb[i] = 0u;
        }
        __syncthreads();
        for (unsigned int i0 = c0; i0 < c1; i0 += blockDim.x) {
            unsigned int i = i0 + threadIdx.x;
            bool valid = i < c1;
            unsigned int b = valid ? w.flyKey[i] : 0u;
            unsigned int off = smem_count_aggregated(cntb, b, valid);
            if (valid) {
                unsigned int dst = base[b] + off;
                const uint4 *src = reinterpret_cast<const uint4 *>(&w.recB[i]);
                uint4 *d = reinterpret_cast<uint4 *>(&w.recA[dst]);
                uint4 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3];
                d[0] = v0; d[1] = v1; d[2] = v2; d[3] = v3;
                if (MULTI) *reinterpret_cast<uint4 *>(&w.recxA[dst]) = *reinterpret_cast<const uint4 *>(&w.recxB[i]);
            }
        }
        __syncthreads();
    }
}

// ---- host-side launchers -------------------------------------------------------------------
static size_t scratch_bytes(int nbins) { return (size_t)scratch_words(nbins) * sizeof(unsigned int); }

template <bool MULTI, int EV>
static cudaError_t launch_event_t(const WfArgs &w, int blocks, cudaStream_t s)
{
    size_t smem = scratch_bytes(w.t.P.nbins);
    cudaFuncSetAttribute(wf_event_kernel<MULTI, EV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wf_event_kernel<MULTI, EV><<<blocks, kThreads, smem, s>>>(w);
    return cudaGetLastError();
}

cudaError_t wf_launch_event(const WfArgs &w, bool multi, int ev, int blocks, cudaStream_t s)
{
    if (multi) {
        if (ev == EV_EMIT) return launch_event_t<true, EV_EMIT>(w, blocks, s);
        if (ev == EV_SCATTER) return launch_event_t<true, EV_SCATTER>(w, blocks, s);
        if (ev == EV_CONT) return launch_event_t<true, EV_CONT>(w, blocks, s);
        return launch_event_t<true, EV_ESCAPE>(w, blocks, s);
    }
    if (ev == EV_EMIT) return launch_event_t<false, EV_EMIT>(w, blocks, s);
    if (ev == EV_SCATTER) return launch_event_t<false, EV_SCATTER>(w, blocks, s);
    if (ev == EV_CONT) return launch_event_t<false, EV_CONT>(w, blocks, s);
    return launch_event_t<false, EV_ESCAPE>(w, blocks, s);
}

static size_t fly_smem(int nbins) { return scratch_bytes(nbins) + (size_t)(kThreads / 32) * EV_COUNT * kStage * sizeof(unsigned int); }

template <bool MULTI, bool DENSE, int MODE>
static cudaError_t launch_fly_t(const WfArgs &w, int blocks, cudaStream_t s)
{
    size_t smem = fly_smem(w.t.P.nbins);
    cudaFuncSetAttribute(wf_fly_kernel<MULTI, DENSE, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wf_fly_kernel<MULTI, DENSE, MODE><<<blocks, kThreads, smem, s>>>(w);
    return cudaGetLastError();
}

cudaError_t wf_launch_fly(const WfArgs &w, bool multi, int blocks, cudaStream_t s)
{
    const int mode = (w.t.P.lgDebug || w.t.P.lgPlane) ? 0 : (w.t.P.lgSym ? 1 : 2);
    if (multi) {                         // the same compile-time mode flags as the single-grid variants
        if (mode == 2) return launch_fly_t<true, false, 2>(w, blocks, s);
        return mode == 1 ? launch_fly_t<true, false, 1>(w, blocks, s) : launch_fly_t<true, false, 0>(w, blocks, s);
    }
    if (w.t.g1.dense) {
        if (mode == 2) return launch_fly_t<false, true, 2>(w, blocks, s);
        return mode == 1 ? launch_fly_t<false, true, 1>(w, blocks, s) : launch_fly_t<false, true, 0>(w, blocks, s);
    }
    if (mode == 2) return launch_fly_t<false, false, 2>(w, blocks, s);
    return mode == 1 ? launch_fly_t<false, false, 1>(w, blocks, s) : launch_fly_t<false, false, 0>(w, blocks, s);
}

int wf_fly_blocks_per_sm(bool multi)
{
    int nb = 0;
    size_t smem = fly_smem(1024);
    if (multi) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wf_fly_kernel<true, false, 0>, kThreads, smem);
    else       cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wf_fly_kernel<false, false, 0>, kThreads, smem);
    return nb;
}

cudaError_t wf_launch_sort(const WfArgs &w, bool multi, int numSMs, cudaStream_t s)
{
    const int nb = w.t.P.nbins;
    size_t sm1 = sizeof(unsigned int) * (nb + 1);
    wf_hist_kernel<<<numSMs * 4, 256, sm1, s>>>(w.flyKey, w.flyCount, w.hist, nb);
    wf_scan_kernel<<<1, 32, 0, s>>>(w.hist, w.cursor, nb);
    if (multi) wf_scatter_kernel<true><<<numSMs * 8, 256, 2 * sm1, s>>>(w, nb);
    else       wf_scatter_kernel<false><<<numSMs * 8, 256, 2 * sm1, s>>>(w, nb);
    return cudaGetLastError();
}

size_t wf_rec_bytes() { return sizeof(PacketRec); }
size_t wf_recx_bytes() { return sizeof(PacketRecX); }

}  // namespace mcb
