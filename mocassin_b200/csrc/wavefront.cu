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

namespace mcb {

enum { EV_EMIT = 0, EV_SCATTER = 1, EV_ESCAPE = 2, EV_CONT = 3, EV_COUNT = 4 };

constexpr int kChunk = 32;               // flights a warp claims at a time (one atomic per chunk); small, so
                                         // that all resident warps work inside a narrow window of the nu order
constexpr int kStage = 64;               // per-warp staging slots per event list

// One packet between two flights, 64 B = two 32 B sectors.
struct alignas(16) PacketRec {
    float rx, ry, rz, passProb;
    float dx, dy, dz;
    unsigned int rngn;
    float absTau;                        // optical depth so far (non-zero only for continued flights)
    unsigned int istepGen;               // istep (19 bits) | gen << 19 (13 bits)
    unsigned int k;                      // packet index within the call
    int orgC;
    unsigned short nuP, gP;
    unsigned short flagsLast;            // bits 0-1 chType, 2 lgStellar, 3 igpp, 4-6 vHat = -direction
                                         // on x,y,z (mirror reflections only flip signs)
    short xP, yP, zP;
    unsigned short orgG, pad;
};
static_assert(sizeof(PacketRec) == 64, "PacketRec must be 64 bytes");
struct alignas(16) PacketRecX {          // 16 B, multi-grid only: enPacket%xP(1:2) slots
    short mx, my, mz, sx, sy, sz;
    unsigned int pad;
};

// recB: flights in arrival order (written by the event kernels); recA: the same flights
// moved into frequency order (read and updated in place by the FLY kernel).
struct WfArgs {
    TransportArgs t;
    PacketRec *recA, *recB;
    PacketRecX *recxA, *recxB;
    const unsigned int *inList;          // event kernels: positions in recA (NULL: wave 0, packets 0..n-1)
    const unsigned int *inCount;
    unsigned short *flyKey;              // nu key of recB[i]
    unsigned int *flyCount;              // entries in recB / recA
    unsigned int *evList[EV_COUNT];      // positions in recA of flights that ended, per event
    unsigned int *evCount;               // [EV_COUNT]
    int stepBudget;                      // cell crossings per flight per wave (longer flights continue
                                         // in the next wave, so one straggler cannot hold a wave open)
    unsigned int *hist, *cursor;         // [nbins+1]
    unsigned long long *nextFlight;      // FLY work counter
};

template <bool MULTI>
__device__ __forceinline__ void rec_store(PacketRec *rec, PacketRecX *recx, const Lane &L, unsigned int pos)
{
    PacketRec r;
    r.rx = L.rx; r.ry = L.ry; r.rz = L.rz; r.passProb = L.passProb;
    r.dx = L.dx; r.dy = L.dy; r.dz = L.dz;
    r.rngn = L.rng.n;
    r.absTau = L.absTau;
    r.istepGen = ((unsigned int)L.istep & 0x7ffffu) | ((unsigned int)L.gen << 19);
    r.k = (unsigned int)L.k;
    r.orgC = L.orgC;
    r.nuP = (unsigned short)L.nuP; r.gP = (unsigned short)L.gP;
    r.flagsLast = (unsigned short)((L.chType & 3) | (L.lgStellar ? 4 : 0) | (L.igpp ? 8 : 0) |
                                   (L.vx != L.dx ? 16 : 0) | (L.vy != L.dy ? 32 : 0) | (L.vz != L.dz ? 64 : 0));
    r.xP = (short)L.xP; r.yP = (short)L.yP; r.zP = (short)L.zP;
    r.orgG = (unsigned short)L.orgG; r.pad = 0;
    const uint4 *src = reinterpret_cast<const uint4 *>(&r);
    uint4 *dst = reinterpret_cast<uint4 *>(&rec[pos]);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
    if (MULTI) {
        PacketRecX x;
        x.mx = (short)L.mx; x.my = (short)L.my; x.mz = (short)L.mz;
        x.sx = (short)L.sx; x.sy = (short)L.sy; x.sz = (short)L.sz; x.pad = 0;
        *reinterpret_cast<uint4 *>(&recx[pos]) = *reinterpret_cast<const uint4 *>(&x);
    }
}

template <bool MULTI>
__device__ __forceinline__ void rec_load(const TransportArgs &t, const PacketRec *rec, const PacketRecX *recx,
                                         Lane &L, unsigned int pos)
{
    PacketRec r;
    const uint4 *src = reinterpret_cast<const uint4 *>(&rec[pos]);
    uint4 *dst = reinterpret_cast<uint4 *>(&r);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
    L.k = (long long)r.k;
    L.rng.init(t.seed, (unsigned long long)(t.firstId + (long long)r.k), (uint32_t)t.iStar, r.rngn);
    L.rx = r.rx; L.ry = r.ry; L.rz = r.rz; L.passProb = r.passProb;
    L.dx = r.dx; L.dy = r.dy; L.dz = r.dz;
    // vHat = direction, except for the signs a mirror reflection flipped (continued flights)
    L.vx = (r.flagsLast & 16) ? -r.dx : r.dx;
    L.vy = (r.flagsLast & 32) ? -r.dy : r.dy;
    L.vz = (r.flagsLast & 64) ? -r.dz : r.dz;
    L.absTau = r.absTau;
    L.segs = 0; L.istep = (int)(r.istepGen & 0x7ffffu); L.gen = (int)(r.istepGen >> 19);
    L.nuP = r.nuP; L.gP = r.gP;
    L.chType = r.flagsLast & 3; L.lgStellar = (r.flagsLast >> 2) & 1; L.igpp = (r.flagsLast >> 3) & 1;
    L.lastNuP = r.nuP;                       // a stored packet's last emission is its current nu
    L.xP = r.xP; L.yP = r.yP; L.zP = r.zP;
    L.orgG = r.orgG; L.orgC = r.orgC;
    L.fate = 0; L.pendFate = FATE_ESCAPED; L.planeG = 0;
    if (MULTI) {
        PacketRecX x;
        *reinterpret_cast<uint4 *>(&x) = *reinterpret_cast<const uint4 *>(&recx[pos]);
        L.mx = x.mx; L.my = x.my; L.mz = x.mz; L.sx = x.sx; L.sy = x.sy; L.sz = x.sz;
    } else {
        L.mx = L.xP; L.my = L.yP; L.mz = L.zP;       // single grid: the mother slot is the cell
        L.sx = L.sy = L.sz = -1;
    }
}

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
                T.start_packet(L, (long long)i);
            } else {
                rec_load<MULTI>(w.t, w.recA, w.recxA, L, w.inList[i]);
            }
            if (EV == EV_EMIT) {
                T.do_emit(L);
                if (L.phase == PH_ESCAPE) T.do_escape(L);     // packets below the ionisation edge
            } else if (EV == EV_SCATTER) {
                T.do_scatter(L);
            } else {
                T.do_escape(L);
            }
        }
        bool toFly = valid && L.phase == PH_FLY;
        unsigned int pos = warp_append(toFly, w.flyCount);
        if (toFly) {
            L.vx = L.dx; L.vy = L.dy; L.vz = L.dz; L.absTau = 0.f;     // fresh flight
            rec_store<MULTI>(w.recB, w.recxB, L, pos);
            w.flyKey[pos] = (unsigned short)L.nuP;
        }
    }
    scratch_flush(w.t, smem);
}

// ---- FLY: the hot loop ------------------------------------------------------------------
// Lanes claim flights in frequency order in chunks of kChunk consecutive records (one global
// atomic per chunk, records of the chunk prefetched into L2) and cross cells until the next
// event; ended flights are written back in place and their positions staged per warp in
// shared memory, flushed 32 at a time (one global atomic per 32 events, coalesced stores).
template <bool MULTI, bool DENSE>
__global__ void __launch_bounds__(kThreads, 4) wf_fly_kernel(const __grid_constant__ WfArgs w)
{
    extern __shared__ unsigned int smem[];
    scratch_init(smem, w.t.P.nbins);
    Transport<MULTI, DENSE> T(w.t, smem, smem + C_COUNT * kThreads);
    unsigned int *stage = smem + C_COUNT * kThreads + w.t.P.nbins + (threadIdx.x >> 5) * (EV_COUNT * kStage);
    const unsigned int FULL = 0xffffffffu;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int total = *w.flyCount;
    unsigned int cur = 0, end = 0;                   // this warp's chunk [cur, end)
    bool exhausted = false;
    int nStaged[EV_COUNT] = {0, 0, 0, 0};
    int budget = 0;
    Lane L;
    L.phase = PH_NEED;
    for (;;) {
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
                rec_load<MULTI>(w.t, w.recA, w.recxA, L, cur + rank);
                L.phase = PH_FLY;
                budget = w.stepBudget;
                L.fate = (int)(cur + rank);          // position in recA (fate is unused while flying)
            }
            unsigned int took = (unsigned int)__popc(need) < avail ? (unsigned int)__popc(need) : avail;
            cur += took;
            need = __ballot_sync(FULL, L.phase == PH_NEED);
        }
        if (exhausted && L.phase == PH_NEED) L.phase = PH_DONE;
        if (__ballot_sync(FULL, L.phase == PH_FLY) == 0u) break;
        const unsigned int pos = (unsigned int)L.fate;
        if (L.phase == PH_FLY) {
            T.step(L, false);
            if (L.phase == PH_FLY) {
                L.fate = (int)pos;
                if (--budget <= 0) L.phase = PH_CONT;           // out of budget: continue next wave
            }
        }
        // flights that just ended: write the record back and stage its position
        bool ended = L.phase == PH_EMIT || L.phase == PH_SCATTER || L.phase == PH_ESCAPE || L.phase == PH_CONT;
        if (__ballot_sync(FULL, ended)) {
            if (ended) {
                rec_store<MULTI>(w.recA, w.recxA, L, pos);
                smem[C_SEGMENTS * kThreads + threadIdx.x] += L.segs;   // this flight's cell crossings
                if (w.t.segsArr) w.t.segsArr[L.k] += L.segs;
                L.segs = 0;
            }
#pragma unroll
            for (int ev = 0; ev < EV_COUNT; ++ev) {
                const int ph = ev == EV_EMIT ? PH_EMIT : ev == EV_SCATTER ? PH_SCATTER : ev == EV_ESCAPE ? PH_ESCAPE : PH_CONT;
                unsigned int m = __ballot_sync(FULL, L.phase == ph);
                if (!m) continue;
                if (L.phase == ph) stage[ev * kStage + nStaged[ev] + __popc(m & ((1u << lane) - 1u))] = pos;
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

// ---- counting sort of the ready flights by frequency bin: recB -> recA -------------------
__global__ void wf_hist_kernel(const unsigned short *key, const unsigned int *count, unsigned int *hist, int nb)
{
    extern __shared__ unsigned int sh[];
    for (int i = threadIdx.x; i <= nb; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const unsigned int n = *count;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&sh[key[i]], 1u);
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
        for (unsigned int i = c0 + threadIdx.x; i < c1; i += blockDim.x) atomicAdd(&cntb[w.flyKey[i]], 1u);
        __syncthreads();
        for (int i = threadIdx.x; i <= nb; i += blockDim.x) {
            base[i] = cntb[i] ? atomicAdd(&w.cursor[i], cntb[i]) : 0u;
            cntb[i] = 0u;
        }
        __syncthreads();
        for (unsigned int i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
            unsigned int b = w.flyKey[i];
            unsigned int dst = base[b] + atomicAdd(&cntb[b], 1u);
            const uint4 *src = reinterpret_cast<const uint4 *>(&w.recB[i]);
            uint4 *d = reinterpret_cast<uint4 *>(&w.recA[dst]);
            uint4 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3];
            d[0] = v0; d[1] = v1; d[2] = v2; d[3] = v3;
            if (MULTI) *reinterpret_cast<uint4 *>(&w.recxA[dst]) = *reinterpret_cast<const uint4 *>(&w.recxB[i]);
        }
        __syncthreads();
    }
}

// ---- host-side launchers -------------------------------------------------------------------
static size_t scratch_bytes(int nbins) { return (size_t)(C_COUNT * kThreads + nbins) * sizeof(unsigned int); }

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

template <bool MULTI, bool DENSE>
static cudaError_t launch_fly_t(const WfArgs &w, int blocks, cudaStream_t s)
{
    size_t smem = fly_smem(w.t.P.nbins);
    cudaFuncSetAttribute(wf_fly_kernel<MULTI, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wf_fly_kernel<MULTI, DENSE><<<blocks, kThreads, smem, s>>>(w);
    return cudaGetLastError();
}

cudaError_t wf_launch_fly(const WfArgs &w, bool multi, int blocks, cudaStream_t s)
{
    if (multi) return launch_fly_t<true, false>(w, blocks, s);
    if (w.t.g1.dense) return launch_fly_t<false, true>(w, blocks, s);
    return launch_fly_t<false, false>(w, blocks, s);
}

int wf_fly_blocks_per_sm(bool multi)
{
    int nb = 0;
    size_t smem = fly_smem(1024);
    if (multi) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wf_fly_kernel<true, false>, kThreads, smem);
    else       cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wf_fly_kernel<false, false>, kThreads, smem);
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
