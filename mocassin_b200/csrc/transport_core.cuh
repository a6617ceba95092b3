// transport_core.cuh -- device code shared by the persistent transport kernel
// (transport.cu) and the wave-front kernels (wavefront.cu): packet state, sampling
// primitives and the Transport<MULTI> phase functions.  See transport.cu for the design.
#pragma once
#include "types.h"
#include "detmath.cuh"
#include "locate.cuh"
#include "philox.cuh"

#include <cuda_runtime.h>

#ifndef MCB_FASTWALL
#define MCB_FASTWALL 1                   // A/B build switches (build.py --variant), see profiles/README.md
#endif
#ifndef MCB_LEANRNG
#define MCB_LEANRNG 0
#endif

namespace mcb {

enum { CH_STELLAR = 0, CH_DIFFEXT = 1, CH_DIFFUSE = 2, CH_DUSTEMI = 3 };
enum { PH_NEED = 0, PH_EMIT = 1, PH_FLY = 2, PH_SCATTER = 3, PH_ESCAPE = 4, PH_DONE = 5, PH_CONT = 6 };
enum { FATE_ESCAPED = 1, FATE_LINE = 2, FATE_DROPPED = 3, FATE_TRAPPED = 4, FATE_EARLY = 5 };

constexpr int kThreads = 256;
constexpr int kRecursionLimit = 5000;    // constants_mod.f90:56
constexpr int kBatch = 8;                // lanes that must wait for a rare phase before it runs
constexpr int kAggSteps = 6;             // first steps of a flight whose tallies are warp-aggregated

struct Lane {
    Rng rng;
    float rx, ry, rz;        // rVec
    float vx, vy, vz;        // vHat
    float iax, iay, iaz;     // RN(1/|vHat|) per axis: magnitudes change only when a flight starts
    float dx, dy, dz;        // enPacket%direction (direction at last emission / scatter)
    float absTau, passProb;
    int xP, yP, zP, gP;      // local indices in the current grid
    int mx, my, mz;          // enPacket%xP(1),yP(1),zP(1): mother-grid slot
    int sx, sy, sz;          // enPacket%xP(2),yP(2),zP(2): sub-grid slot
    int igpp;                // 0 = mother slot, 1 = sub slot (may go stale like the reference's)
    int nuP, lgStellar;
    int orgG, orgC;          // enPacket%origin
    int chType;
    int istep, gen;
    unsigned int segs;
    int lastNuP, fate, pendFate;
    unsigned long long planeBase;   // (nuP-1)*(nCells+1) of grid planeG: index of the packet's nu-plane
    int planeG;                     // grid the cached planeBase belongs to (0 = none)
    long long k;
    int phase;
};

// reciprocals for div_rn() below; call whenever vHat gets new magnitudes (emission, scattering,
// record load).  Mirror reflections only flip signs.
__device__ __forceinline__ void set_inv(Lane &L)
{
    L.iax = 1.f / fabsf(L.vx); L.iay = 1.f / fabsf(L.vy); L.iaz = 1.f / fabsf(L.vz);
}

// Correctly rounded num / v from the correctly rounded reciprocal ia = RN(1/|v|):
// q = RN(num*y), r = num - v*q (exact in an FMA), RN(q + r*y) = RN(num/v) (Markstein's
// correction step; scripts/div_check.c, tests: test_device_division_is_ieee).  Three
// instructions instead of the ~11 + range check of the IEEE divider, and no slow path, so a
// zero numerator needs no special care.  Valid while no intermediate underflows or overflows:
// |v| in (1e-10, 1] (the caller's `moving` test) and num = 0 or 2e-31 < |num| < 3e28; num is a
// difference of grid coordinates, and mcb200_set_grid rejects axes outside 1e-10..1e27 cm.
__device__ __forceinline__ float div_rn(float num, float v, float ia)
{
    float y = copysignf(ia, v);
    float q = num * y;
    float r = __fmaf_rn(-v, q, num);
    return __fmaf_rn(r, y, q);
}

template <bool DENSE = false>
__device__ __forceinline__ int active_at(const DevGrid &g, int x, int y, int z)
{
    if (DENSE || g.dense) return 1 + (z - 1) + g.nz * ((y - 1) + g.ny * (x - 1));
    // nx*ny*nz < 2^31 (checked at upload): 32-bit index arithmetic
    return __ldg(&g.active[(x - 1) + g.nx * ((y - 1) + g.ny * (z - 1))]);
}

// getNu2 (photon_mod.f90:720-764) on a contiguous non-decreasing CDF row
__device__ __forceinline__ int sample_cdf(Rng &rng, const float *cdf, int nb)
{
    float u = rng.uniform();
    for (int i = 1; i <= 10000; ++i) {
        if (u == 0.f || u == 1.f || u == 0.9999999f) u = rng.uniform(); else break;
    }
    int lo = 0, hi = nb;                 // number of leading entries with u >= cdf
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (u >= __ldg(&cdf[mid])) lo = mid + 1; else hi = mid;
    }
    int nuP = lo < 1 ? 1 : lo;
    if (nuP < nb - 1) nuP = nuP + 1;
    return nuP;
}

// getNu2 on a strided row (linePDF, debug only): faithful linear scan.  getNu2 takes the scan
// bound and the "+1" rule from the module variable nbins whatever the length of the row
// (photon_mod.f90:747-757): nscan = entries that exist (nLines), nb = nbins.
__device__ __forceinline__ int sample_cdf_strided(Rng &rng, const float *cdf, size_t stride, int nscan, int nb)
{
    float u = rng.uniform();
    for (int i = 1; i <= 10000; ++i) {
        if (u == 0.f || u == 1.f || u == 0.9999999f) u = rng.uniform(); else break;
    }
    int nuP = 1;
    for (int is = 1; is <= nscan && is <= nb; ++is) {
        if (u >= __ldg(&cdf[(size_t)(is - 1) * stride])) nuP = is; else break;
    }
    if (nuP < nb - 1) nuP = nuP + 1;
    if (nuP > nscan) nuP = nscan;        // the reference would index past the row here
    return nuP;
}

// vector_mod.f90:303-314
__device__ __forceinline__ void random_unit_vector(Rng &rng, float &u, float &v, float &w)
{
    float r1 = rng.uniform();
    w = 2.f * r1 - 1.f;
    float t = sqrtf(1.f - w * w);
    float r2 = rng.uniform();
    float ang = 3.141592654f * (2.f * r2 - 1.f);
    float sn, cs;
    dm_sincosf(ang, sn, cs);
    u = t * cs;
    v = t * sn;
}

// isotropic direction of initPhotonPacket (photon_mod.f90:650-675)
__device__ __forceinline__ void new_direction(const DevParams &P, Lane &L, bool stellar)
{
    for (int irepeat = 1; irepeat <= 1000000; ++irepeat) {
        random_unit_vector(L.rng, L.dx, L.dy, L.dz);
        if (L.dx != 0.f && L.dy != 0.f && L.dz != 0.f) break;
    }
    if (P.lgSym && stellar && !P.lgMultistars) {
        if (L.dx < 0.f) L.dx = -L.dx;
        if (L.dy < 0.f) L.dy = -L.dy;
        if (L.dz < 0.f) L.dz = -L.dz;
    }
}

// photon_mod.f90:2875-2974; returns ierr
__device__ __forceinline__ int hg(const DevParams &P, Lane &L)
{
    float v0 = L.dx, v1 = L.dy, v2 = L.dz;
    float hgg = __ldg(&P.gSca[L.nuP - 1]);
    float random0 = L.rng.uniform();
    float s = 2.f * random0 - 1.f;
    float cost, sint;
    if (hgg >= 0.0001f) {
        float q = (1.f - hgg * hgg) / (1.f + hgg * s);
        cost = 0.5f / hgg * (1.f + hgg * hgg - q * q);
    } else {
        cost = s;
    }
    if (cost >= 1.0f) { cost = 1.0f; sint = 0.f; }
    else if (cost < -1.0f) { cost = -1.0f; sint = 0.f; }
    else sint = sqrtf(1.f - cost * cost);
    float random = L.rng.uniform();
    float phi = (2.f * 3.141592654f) * random;
    float cosp, sinp;
    dm_sincosf(phi, sinp, cosp);
    float denom = sqrtf(1.f - v2 * v2);
    float o0, o1, o2;
    if (denom > 0.001f) {
        o0 = sint / denom * (v0 * v2 * cosp - v1 * sinp) + v0 * cost;
        o1 = sint / denom * (v1 * v2 * cosp + v0 * sinp) + v1 * cost;
        o2 = -sint * cosp * denom + v2 * cost;
    } else {
        o0 = sint * cosp;
        o1 = sint * sinp;
        o2 = (v2 >= 0.f) ? cost : -cost;
    }
    bool f0 = (o0 >= 0.f || o0 < 0.f), f1 = (o1 >= 0.f || o1 < 0.f), f2 = (o2 >= 0.f || o2 < 0.f);
    if ((fabsf(o0) <= 1.f && fabsf(o1) <= 1.f && fabsf(o2) <= 1.f) && f0 && f1 && f2) {
        L.dx = o0; L.dy = o1; L.dz = o2;
        return 0;
    } else if ((fabsf(o0) >= 1.f) || (fabsf(o1) >= 1.f) || ((fabsf(o2) >= 1.f) && f0 && f1 && f2)) {
        return 0;                        // reference renormalises a local copy only (:2951-2952)
    }
    return 1;
}

// Philox stream id of packet k of this call.  Stellar / diffuse-source packets: global packet
// index.  Resonance-line packets: 2^40 + global enumeration index (cell looked up in the
// rank-local prefix table).
__device__ __forceinline__ int res_cell_of(const TransportArgs &a, long long k)
{
    int lo = 0, hi = a.nResCells;        // last c with resPrefix[c] <= k
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if ((long long)__ldg(&a.resPrefix[mid]) <= k) lo = mid; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ unsigned long long packet_pid(const TransportArgs &a, long long k)
{
    if (!a.resCells) return a.pidBase + (unsigned long long)(a.firstId + k);
    int c = res_cell_of(a, k);
    return (1ull << 40) + a.resCells[c].gid + (unsigned long long)(k - (long long)a.resPrefix[c]);
}

// MODE (chosen by the launcher from the call's flags, so their tests leave the crossing loop):
// 0 generic; 1 neither debug tallies nor plane-parallel illumination; 2 = 1 and not symmetricXYZ.
// LIMIT = false: the caller (FLY kernel) folds the iteration limit of :2838-2846 into its own
// per-flight step budget instead of testing it on every crossing.
// LEANRNG: step() draws through Rng::uniform_at (no cached Philox block, packet id rebuilt from the
// packet index): frees the registers of the cache in the crossing loop of the FLY kernel.
template <bool MULTI, bool DENSE = false, int MODE = 0, bool LIMIT = true, bool LEANRNG = false>
struct Transport {
    // single dense grid: the table index of the packet's cell is carried along and advanced with
    // the cell indices instead of being rebuilt from (x,y,z) on every crossing
    static constexpr bool kInc = DENSE && !MULTI;
    const TransportArgs &a;
    unsigned int *cnt;                   // per-thread event counters in shared memory
    unsigned int *qph;                   // per-CTA Qphot histogram in shared memory

    __device__ __forceinline__ Transport(const TransportArgs &a_, unsigned int *cnt_, unsigned int *qph_)
        : a(a_), cnt(cnt_), qph(qph_) {}

    __device__ __forceinline__ const DevGrid &G(int gP) const
    {
        if (MULTI) return a.grids[gP - 1];
        return a.g1;
    }
    __device__ __forceinline__ float uni(Lane &L) const
    {
        if (LEANRNG) return L.rng.uniform_at(a.seed, packet_pid(a, L.k), a.rngStream);
        return L.rng.uniform();
    }
    __device__ __forceinline__ bool debug() const { return MODE >= 1 ? false : (bool)a.P.lgDebug; }
    __device__ __forceinline__ bool plane() const { return MODE >= 1 ? false : (bool)a.P.lgPlane; }
    __device__ __forceinline__ bool sym() const { return MODE == 2 ? false : (bool)a.P.lgSym; }
    __device__ __forceinline__ void count(int which) { cnt[which * kThreads + threadIdx.x]++; }
    __device__ __forceinline__ void fail(Lane &L, int code)
    {
        atomicMax(a.errFlag, code);
        finish(L, FATE_DROPPED);
    }

    __device__ __forceinline__ void finish(Lane &L, int fate)
    {
        L.fate = fate;
        // energyPacketDriver bookkeeping (photon_mod.f90:120-126)
        if (L.gen >= kRecursionLimit) count(C_TRAPPED);
        if (fate == FATE_DROPPED) count(C_DROPPED);
        cnt[C_SEGMENTS * kThreads + threadIdx.x] += L.segs;
        if (a.fates) {
            unsigned int tot = L.segs + (a.segsArr ? a.segsArr[L.k] : 0u);
            int4 f = make_int4((int)tot, L.gen, L.lastNuP, fate);
            reinterpret_cast<int4 *>(a.fates)[L.k] = f;
        }
        L.phase = PH_NEED;
    }
    // leave through the (deferred) escape tally
    __device__ __forceinline__ void escape(Lane &L, int fate)
    {
        L.pendFate = fate;
        L.phase = PH_ESCAPE;
    }

    // ---- PH_ESCAPE: escape tally (photon_mod.f90:373-462 and the six copies in pathSegment)
    // agg: called by the converged lanes of an event kernel -- packets that escape together mostly
    // left the same cell at the same frequency (every stellar packet of a wave counts into the star's
    // cell), so equal targets are counted once per warp (MATCH.ANY) instead of lane by lane on a few
    // hundred hot addresses
    __device__ __forceinline__ void do_escape(Lane &L, bool agg = false)
    {
        const DevParams &P = a.P;
        // The direction bins select the viewing-angle planes; without viewing angles (nAngleBins = 0)
        // their only other trace is the reference's idirT/idirP >= 1 sanity stop, which a finite
        // direction cannot trip (acos >= 0; a negative atan bin is wrapped), so acos/atan are skipped.
        int idirT = 1, idirP = 1;
        if (P.nAngleBins > 0) {
            if (P.lgSym) idirT = (int)(dm_acosf(fabsf(L.dz)) / P.dTheta) + 1;
            else         idirT = (int)(dm_acosf(L.dz) / P.dTheta) + 1;
            if (idirT > P.totT) idirT = P.totT;
            if (fabsf(L.dx) < 1.e-35f) idirP = 0;
            else if (P.lgSym) idirP = (int)(dm_atanf(fabsf(L.dy) / fabsf(L.dx)) / P.dPhi);
            else              idirP = (int)(dm_atanf(L.dy / L.dx) / P.dPhi);
            if (idirP < 0) idirP = P.totP + idirP;
            idirP = idirP + 1;
            if (idirP > P.totP) idirP = P.totP;
            if (idirT < 1 || idirP < 1) { fail(L, 10); return; }
        }
        if (L.orgG < 1 || L.orgG > P.nGrids) { fail(L, 11); return; }
        if (L.orgC < 0) { fail(L, 12); return; }
        const DevGrid &g = G(L.orgG);
        size_t plane = (size_t)(g.nCells + 1) * (size_t)(P.nbins + 1);
        size_t base = (size_t)L.orgC + (size_t)(g.nCells + 1) * (size_t)L.nuP;
        if (P.nAngleBins > 0) {
            int vt = __ldg(&P.vpPtheta[idirT]), vp = __ldg(&P.vpPphi[idirP]);
            if (vt > 0 && __ldg(&P.vpPhi[vt]) < 0.f) {
                atomicAdd(&g.escQ[base + plane * vt], 1u);
                atomicAdd(&g.escQ[base], 1u);
            } else if (vt == vp || __ldg(&P.vpTheta[vp]) == __ldg(&P.vpTheta[vt]) ||
                       __ldg(&P.vpPhi[vt]) == __ldg(&P.vpPhi[vp])) {
                atomicAdd(&g.escQ[base + plane * vt], 1u);
                if (vt != 0) atomicAdd(&g.escQ[base], 1u);
            } else {
                atomicAdd(&g.escQ[base], 1u);
            }
        } else if (agg) {
            const unsigned int act = __activemask();
            const unsigned long long key = (unsigned long long)base | ((unsigned long long)(unsigned int)L.orgG << 48);
            const unsigned int peers = __match_any_sync(act, key);
            if ((int)(threadIdx.x & 31u) == __ffs(peers) - 1) atomicAdd(&g.escQ[base], (unsigned int)__popc(peers));
        } else {
            atomicAdd(&g.escQ[base], 1u);
        }
        count(C_ESCAPED);
        if (L.pendFate == FATE_EARLY) count(C_EARLY);
        finish(L, L.pendFate);
    }

    // J estimator add (photon_mod.f90:1563-1574, 1822-1833) in fixed point.  Called by all
    // flying lanes of the warp at one converged site.  When packets are processed in
    // frequency order many lanes of a warp sit in the same (cell, nu) during their first
    // steps (everything starts in the star's cell): those adds are pre-reduced inside the
    // warp (MATCH.ANY + REDUX) so the L2 atomic unit of that one address is not hammered
    // by 32 separate reductions per warp.  Integer sums are exact -> same result.
    __device__ __forceinline__ void j_add(const DevGrid &g, const Lane &L, int cell, size_t tix, float len, bool aggregate)
    {
        long long q = __float2ll_rn(len * g.invLenUnit);
        unsigned long long *Q = (debug() && !L.lgStellar) ? g.JdifQ : g.JsteQ;
        unsigned long long *addr = &Q[tix];
        bool live = kInc || cell > 0;    // sink row 0: never read by the reference
        if (aggregate) {
            // one round: the lanes that share the address of the first live lane are summed
            // (REDUX) and added once; everybody else falls through to a plain reduction
            unsigned int act = __activemask();
            bool small = live && q >= 0 && q < (1ll << 26);
            unsigned int cand = __ballot_sync(act, small);
            if (cand) {
                int lead = __ffs(cand) - 1;
                unsigned long long la = __shfl_sync(act, (unsigned long long)addr, lead);
                unsigned int same = __ballot_sync(act, small && (unsigned long long)addr == la);
                if (small && (unsigned long long)addr == la) {
                    unsigned int sum = __reduce_add_sync(same, (unsigned int)q);
                    if ((int)(threadIdx.x & 31u) == lead) atomicAdd(addr, (unsigned long long)sum);
                    return;
                }
            }
        }
        if (live) atomicAdd(addr, (unsigned long long)q);
    }

    // ---- first packet of a history (photon_mod.f90:93-170) ---------------------------
    __device__ __forceinline__ void start_packet(Lane &L, long long k)
    {
        const DevParams &P = a.P;
        L.k = k;
        L.segs = 0; L.gen = 0; L.fate = 0; L.lastNuP = 0; L.planeG = 0;
        L.mx = L.my = L.mz = -1;
        L.sx = L.sy = L.sz = -1;
        if (a.resCells) {                // resonance-line packet: "diffuse" from a cell centre
            int c = res_cell_of(a, k);
            const ResCell rc = a.resCells[c];
            L.rng.init(a.seed, (1ull << 40) + rc.gid + (unsigned long long)(k - (long long)a.resPrefix[c]), a.rngStream);
            L.chType = CH_DIFFUSE;
            L.gP = rc.grid;
            L.rx = rc.px; L.ry = rc.py; L.rz = rc.pz;
            if (L.gP == 1) { L.mx = rc.x; L.my = rc.y; L.mz = rc.z; }
            else { L.sx = rc.x; L.sy = rc.y; L.sz = rc.z; L.mx = rc.mx; L.my = rc.my; L.mz = rc.mz; }
            L.phase = PH_EMIT;
            return;
        }
        L.rng.init(a.seed, a.pidBase + (unsigned long long)(a.firstId + k), a.rngStream);
        if (a.iStar >= 1) {
            const int *si = &P.starIdx[4 * (a.iStar - 1)];
            L.chType = CH_STELLAR;
            L.gP = __ldg(&si[3]);
            L.rx = __ldg(&P.starPos[3 * (a.iStar - 1) + 0]);
            L.ry = __ldg(&P.starPos[3 * (a.iStar - 1) + 1]);
            L.rz = __ldg(&P.starPos[3 * (a.iStar - 1) + 2]);
            int ix = __ldg(&si[0]), iy = __ldg(&si[1]), iz = __ldg(&si[2]);
            if (L.gP == 1) { L.mx = ix; L.my = iy; L.mz = iz; } else { L.sx = ix; L.sy = iy; L.sz = iz; }
        } else {
            const DevGrid &g = G(a.difGrid);
            L.chType = CH_DIFFEXT;
            L.gP = a.difGrid;
            // energyPacketDriver :148-150 builds posDiff with zAxis(cellLoc(2)) but
            // newPhotonPacket('diffExt') :882-884 overwrites the position from difSource
            L.rx = __ldg(&g.xAxis[a.difX - 1]);
            L.ry = __ldg(&g.yAxis[a.difY - 1]);
            L.rz = __ldg(&g.zAxis[a.difZ - 1]);
            if (L.gP == 1) { L.mx = a.difX; L.my = a.difY; L.mz = a.difZ; } else { L.sx = a.difX; L.sy = a.difY; L.sz = a.difZ; }
        }
        L.phase = PH_EMIT;
    }

    // ---- PH_EMIT: energyPacketRun up to the head of pathSegment (photon_mod.f90:289-487,
    //      768-1059, 1098-1192) ---------------------------------------------------------
    __device__ __forceinline__ void do_emit(Lane &L)
    {
        const DevParams &P = a.P;
        L.gen++;
        int igp = (L.gP == 1) ? 0 : 1;
        int ex = igp ? L.sx : L.mx, ey = igp ? L.sy : L.my, ez = igp ? L.sz : L.mz;
        const float *cdf;                // contiguous CDF row to sample
        int cell = 0;
        bool stellar = false;
        if (L.chType == CH_STELLAR) {
            cdf = P.starCdf + (size_t)a.iStar * P.nbins;
            cell = __ldg(&P.starCell[a.iStar - 1]);
            if (cell < 0) { fail(L, 35); return; }
            stellar = true;
        } else if (L.chType == CH_DIFFEXT) {
            cdf = P.starCdf;
            cell = active_at(G(L.gP), ex, ey, ez);
            if (cell < 0) { fail(L, 38); return; }
        } else {                         // CH_DIFFUSE / CH_DUSTEMI: re-emission in the absorbing cell
            const DevGrid &g = G(L.gP);
            cell = active_at(g, ex, ey, ez);
            if (cell <= 0) { fail(L, 40); return; }
            if (L.chType == CH_DIFFUSE) {
                float random = 1.f - L.rng.uniform();
                if (random <= __ldg(&g.totalLines[cell])) {
                    // non-ionising line packet: leaves silently (:923-957, :475-485)
                    int nuL = 0;
                    if (P.lgDebug) {
                        nuL = sample_cdf_strided(L.rng, g.linePDF + cell, (size_t)(g.nCells + 1), P.nLines, P.nbins);
                        atomicAdd(&g.lineQ[(size_t)(nuL - 1) * (size_t)(g.nCells + 1) + (size_t)cell], 1u);
                    }
                    L.lastNuP = nuL;
                    count(C_LINE);
                    finish(L, FATE_LINE);
                    return;
                }
            }
            cdf = g.pdfT + (size_t)cell * P.nbins;
        }
        int nuP = sample_cdf(L.rng, cdf, P.nbins);
        L.lastNuP = nuP;
        if (nuP >= P.nbins) { fail(L, 33); return; }   // fatal in the reference (:826,868,962,1018)
        L.lgStellar = stellar ? 1 : 0;
        if (stellar && P.lgPlane) {
            // plane-parallel ionisation: enter through the y=0 face along +y (:561-646)
            const DevGrid &g = G(L.gP);
            const float *xa = g.xAxis, *za = g.zAxis;
            float random = 1.f - L.rng.uniform();
            float x1 = __ldg(&xa[0]), x2 = __ldg(&xa[1]), xm = __ldg(&xa[g.nx - 2]), xn = __ldg(&xa[g.nx - 1]);
            L.rx = -(x2 - x1) / 2.f + random * ((x2 - x1) / 2.f + (xn - xm) / 2.f + xn);
            if (L.rx < x1) L.rx = x1;
            if (L.rx > xn) L.rx = xn;
            ex = locate_axis(xa, g.nx, L.rx);
            if (ex < g.nx) { if (ex >= 1 && L.rx >= (__ldg(&xa[ex - 1]) + __ldg(&xa[ex])) / 2.f) ex = ex + 1; }
            L.ry = 0.f;
            ey = 1;
            random = 1.f - L.rng.uniform();
            float z1 = __ldg(&za[0]), z2 = __ldg(&za[1]), zm = __ldg(&za[g.nz - 2]), zn = __ldg(&za[g.nz - 1]);
            L.rz = -(z2 - z1) / 2.f + random * ((z2 - z1) / 2.f + (zn - zm) / 2.f + zn);
            if (L.rz < z1) L.rz = z1;
            if (L.rz > zn) L.rz = zn;
            ez = locate_axis(za, g.nz, L.rz);
            if (ez < g.nz) {             // sic: xAxis(zP) in the z test (:612)
                if (ez >= 1 && ez <= g.nx && L.rz >= (__ldg(&xa[ez - 1]) + __ldg(&za[ez])) / 2.f) ez = ez + 1;
            }
            if (ex < 1) ex = 1;
            if (ez < 1) ez = 1;
            L.dx = 0.f; L.dy = 1.f; L.dz = 0.f;
            if (igp) { L.sx = ex; L.sy = ey; L.sz = ez; } else { L.mx = ex; L.my = ey; L.mz = ez; }
            if (P.planeDist) atomicAdd(&P.planeDist[(ex - 1) + G(1).nx * (ez - 1)], 1);
            if (ex > g.nx || ez > g.nz) { fail(L, 24); return; }
            cell = active_at(g, ex, ey, ez);
        } else {
            new_direction(P, L, stellar);
        }
        L.orgG = L.gP; L.orgC = cell;
        L.nuP = nuP;
        L.planeG = 0;
        atomicOr(&qph[a.P.nbins + (nuP >> 5)], 1u << (nuP & 31));   // nu-plane touched (flushed per CTA)
        float nu = __ldg(&P.nuArray[nuP - 1]);
        if (stellar && nu > 1.f) atomicAdd(&qph[nuP - 1], 1u);     // Qphot, :859-861
        if (!P.lgDust && nu < P.ionEdge1) {                         // :370-465
            escape(L, FATE_EARLY);
            return;
        }
        // head of pathSegment (:1098-1192)
        count(C_FLIGHTS);
        L.igpp = igp;
        L.xP = ex; L.yP = ey; L.zP = ez;
        {
            const DevGrid &g = G(L.gP);
            if (L.xP <= 0 || L.xP > g.nx || L.yP <= 0 || L.yP > g.ny || L.zP <= 0 || L.zP > g.nz) { fail(L, 52); return; }
        }
        L.vx = L.dx; L.vy = L.dy; L.vz = L.dz;
        set_inv(L);
        L.absTau = 0.f;
        L.passProb = -dm_logf(1.f - L.rng.uniform());
        L.istep = 0;
        L.phase = PH_FLY;
    }

    // ---- PH_SCATTER: new direction after a dust scattering (photon_mod.f90:1750-1799) ----
    __device__ __forceinline__ void do_scatter(Lane &L)
    {
        const DevParams &P = a.P;
        const DevGrid &g = G(L.gP);
        // initPhotonPacket(..., lgHG=.true.) (:1753-1766): slots, origin, non-stellar
        if (L.igpp) { L.sx = L.xP; L.sy = L.yP; L.sz = L.zP; } else { L.mx = L.xP; L.my = L.yP; L.mz = L.zP; }
        L.lgStellar = 0;
        if (P.lgIso) new_direction(P, L, false);
        L.orgG = L.gP;
        {
            int igpi = (L.gP == 1) ? 0 : 1;
            int ox = igpi ? L.sx : L.mx, oy = igpi ? L.sy : L.my, oz = igpi ? L.sz : L.mz;
            if (ox < 1 || ox > g.nx || oy < 1 || oy > g.ny || oz < 1 || oz > g.nz) { fail(L, 24); return; }
            L.orgC = active_at(g, ox, oy, oz);
        }
        if (!P.lgIso) {
            for (int ihg = 1; ihg <= 10; ++ihg) { if (hg(P, L) == 0) break; }
        }
        L.vx = L.dx; L.vy = L.dy; L.vz = L.dz;
        set_inv(L);
        if (!(L.dx >= 0.f || L.dx < 0.f)) { fail(L, 71); return; }
        L.absTau = 0.f;
        L.passProb = -dm_logf(1.f - L.rng.uniform());
        L.phase = PH_FLY;
    }

    // distance to the wall ahead on one axis (photon_mod.f90:1263-1295), completely branch
    // free: W[iP] is the wall ahead for v>0, W[iP-1] for v<0.  A packet sitting exactly on
    // the wall (|dS|<1e-10, in float32 at these magnitudes: w==r) is snapped and stepped
    // over; on the outermost wall of the mother grid the reference `return`s without a
    // tally -> `drop`.  The zero numerator is kept away from the divider (it would take the
    // IEEE slow path and, like any rare branch here, split one lane off the warp for the
    // rest of the step: ncu showed 58 % of the warp trips running a 1-lane straggler).
    __device__ __forceinline__ void wall(const float *W, int n, float v, float ia, float &r, int &iP, bool outer, float &dS,
                                         bool &drop, bool &pos, bool &snapped)
    {
        pos = v > 1.e-10f;
        bool neg = v < -1.e-10f;
        bool moving = pos || neg;
        float w = __ldg(&W[pos ? iP : iP - 1]);
        float num = w - r;
        float d = div_rn(num, v, ia);                // == num / v, bit for bit
        dS = moving ? d : 1.e35f;
        if (__builtin_expect(moving && fabsf(d) < 1.e-10f, 0)) {   // sitting on the wall: snap and step over (rare,
            r = w;                                   // tiny body: reconverges immediately)
            snapped = true;
            if (pos) { if (iP < n) iP = iP + 1; else drop = drop || outer; }
            else     { if (iP > 1) iP = iP - 1; }
        }
    }

    // absorbed: hand over to the next generation (photon_mod.f90:2848-2870)
    __device__ __forceinline__ void absorb(Lane &L, int packetType)
    {
        count(C_ABS);
        if (L.istep >= a.P.safeLimit) { finish(L, FATE_DROPPED); return; }   // :2838-2846
        L.igpp = (L.gP == 1) ? 0 : 1;
        if (L.igpp) { L.sx = L.xP; L.sy = L.yP; L.sz = L.zP; } else { L.mx = L.xP; L.my = L.yP; L.mz = L.zP; }
        L.chType = packetType;
        if (L.gen >= kRecursionLimit) { finish(L, FATE_TRAPPED); return; }
        L.phase = PH_EMIT;
    }

    // ---- PH_FLY: one trip of the cell-crossing loop (photon_mod.f90:1194-2836) ----------
    __device__ __forceinline__ void step(Lane &L, bool aggEarly)
    {
        const DevParams &P = a.P;
        L.istep++;
        L.segs++;
        float dSx = 0.f, dSy = 0.f, dSz = 0.f;
        bool posx = false, posy = false, posz = false;
        int cell;
        // carried table index: the opacity of the cell is requested before the wall arithmetic
        // so that its L2 latency overlaps it
        const bool early = kInc && L.planeG == L.gP;
        float opacEarly = 0.f;
        if (early) opacEarly = __ldg(&a.g1.opacity[L.planeBase]);
        bool snapped = false;
        // Fast path of the wall stage: the three wall distances with none of the
        // special cases.  Those -- an axis sitting on its wall (snap, drop on the outermost wall, the
        // zero distance replaced by the axis end, :1263-1432) or a non-finite distance -- all begin
        // with a moving axis whose |distance| is not >= 1e-10; when no axis is in that state the
        // generic stage below reduces to exactly these lines, and it is skipped.  Otherwise nothing
        // has been modified yet and the generic stage runs from scratch.
        bool fast = false;
        if (MCB_FASTWALL) {
            const DevGrid &g = G(L.gP);
            // multi-grid: only inside a cell of the current grid (an index outside it, or a mother cell
            // that holds a sub-grid, takes the generic stage with its sub-grid entry and its stops)
            bool inside = true;
            int c0 = 1;
            if (MULTI) {
                inside = !(L.xP > g.nx || L.xP < 1 || L.yP > g.ny || L.yP < 1 || L.zP > g.nz || L.zP < 1);
                if (inside) { c0 = active_at(g, L.xP, L.yP, L.zP); inside = c0 >= 0; }
            }
            if (inside) {
                if (sym()) {             // :1248-1261, as at the head of the generic stage (idempotent)
                    const DevGrid &m = G(1);
                    if (L.rx <= m.x1) { L.vx = fabsf(L.vx); L.rx = m.x1; }
                    if (L.ry <= m.y1) { L.vy = fabsf(L.vy); L.ry = m.y1; }
                    if (L.rz <= m.z1) { L.vz = fabsf(L.vz); L.rz = m.z1; }
                }
                const bool px = L.vx > 1.e-10f, py = L.vy > 1.e-10f, pz = L.vz > 1.e-10f;
                const bool mx = px || L.vx < -1.e-10f, my = py || L.vy < -1.e-10f, mz = pz || L.vz < -1.e-10f;
                const float dx = div_rn(__ldg(&g.xWall[px ? L.xP : L.xP - 1]) - L.rx, L.vx, L.iax);
                const float dy = div_rn(__ldg(&g.yWall[py ? L.yP : L.yP - 1]) - L.ry, L.vy, L.iay);
                const float dz = div_rn(__ldg(&g.zWall[pz ? L.zP : L.zP - 1]) - L.rz, L.vz, L.iaz);
                const bool special = (mx && !(fabsf(dx) >= 1.e-10f)) || (my && !(fabsf(dy) >= 1.e-10f)) ||
                                     (mz && !(fabsf(dz) >= 1.e-10f));
                if (__builtin_expect(!special, 1)) {
                    fast = true;
                    posx = px; posy = py; posz = pz;
                    dSx = mx ? dx : 1.e35f; dSy = my ? dy : 1.e35f; dSz = mz ? dz : 1.e35f;
                    cell = MULTI ? c0 : (kInc ? 1 : active_at<DENSE>(g, L.xP, L.yP, L.zP));
                }
            }
        }
        for (int j = 1; !fast; ++j) {
            if (MULTI) {
                const DevGrid &g0 = G(L.gP);
                if (L.xP > g0.nx || L.xP < 1 || L.yP > g0.ny || L.yP < 1 || L.zP > g0.nz || L.zP < 1) { fail(L, 58); return; }
                int c0 = active_at(g0, L.xP, L.yP, L.zP);
                if (c0 < 0) {            // entering a sub-grid (:1206-1243)
                    L.mx = L.xP; L.my = L.yP; L.mz = L.zP;
                    L.gP = -c0;
                    if (L.gP > P.nGrids) { fail(L, 59); return; }
                    const DevGrid &s = G(L.gP);
                    L.xP = locate_axis(s.xAxis, s.nx, L.rx);
                    if (L.xP == 0) L.xP = 1;
                    if (L.xP < s.nx) { if (L.rx > __ldg(&s.xWall[L.xP])) L.xP = L.xP + 1; }
                    L.yP = locate_axis(s.yAxis, s.ny, L.ry);
                    if (L.yP == 0) L.yP = 1;
                    if (L.yP < s.ny) { if (L.ry > __ldg(&s.yWall[L.yP])) L.yP = L.yP + 1; }
                    L.zP = locate_axis(s.zAxis, s.nz, L.rz);
                    if (L.zP == 0) L.zP = 1;
                    if (L.zP < s.nz) { if (L.rz > __ldg(&s.zWall[L.zP])) L.zP = L.zP + 1; }
                    L.igpp = 1;
                }
            }
            const DevGrid &g = G(L.gP);
            if (sym()) {                 // :1248-1261 (always against the mother grid's first point)
                const DevGrid &m = G(1);
                if (L.rx <= m.x1) { L.vx = fabsf(L.vx); L.rx = m.x1; }
                if (L.ry <= m.y1) { L.vy = fabsf(L.vy); L.ry = m.y1; }
                if (L.rz <= m.z1) { L.vz = fabsf(L.vz); L.rz = m.z1; }
            }
            // the reference returns at the first axis found on the outer wall, i.e. before
            // the later axes are looked at; nothing after a `return` is observable
            bool drop = false;
            // outermost wall of the mother grid: `return` (x,z not in plane mode, :1276,1315,1354)
            const bool outerY = L.gP == 1, outerXZ = outerY && !plane();
            wall(g.xWall, g.nx, L.vx, L.iax, L.rx, L.xP, outerXZ, dSx, drop, posx, snapped);
            wall(g.yWall, g.ny, L.vy, L.iay, L.ry, L.yP, outerY, dSy, drop, posy, snapped);
            wall(g.zWall, g.nz, L.vz, L.iaz, L.rz, L.zP, outerXZ, dSz, drop, posz, snapped);
            if (__builtin_expect(drop, 0)) { finish(L, FATE_DROPPED); return; }
            if (__builtin_expect((dSx != dSx) | (dSy != dSy) | (dSz != dSz), 0)) { fail(L, 60); return; }
            if (kInc) {
                if (snapped) L.planeG = 0;           // indices moved: rebuild the carried table index
                cell = 1;                            // every cell of a dense grid is active; the id is
                break;                               // recomputed where it is needed (scattering)
            }
            cell = active_at<DENSE>(g, L.xP, L.yP, L.zP);
            if (!MULTI || cell >= 0) break;
            if (j >= a.P.safeLimit) { fail(L, 63); return; }
        }
        const DevGrid &g = G(L.gP);
        // index of (cell, nuP) in the (0:nCells, nbins) tables: the nu-plane offset is constant
        // during a flight inside one grid and cached in the lane
        if (L.planeG != L.gP) {
            L.planeBase = (unsigned long long)(unsigned int)(L.nuP - 1) * (unsigned int)(g.nCells + 1);
            if (kInc) L.planeBase += (unsigned int)active_at<DENSE>(g, L.xP, L.yP, L.zP);
            L.planeG = L.gP;
        }
        size_t tix = kInc ? (size_t)L.planeBase : (size_t)L.planeBase + (size_t)(unsigned int)cell;
        float opac = (early && !snapped) ? opacEarly : __ldg(&g.opacity[tix]);

        // cells on a wall (:1395-1397): the axis end coordinate replaces a zero distance
        if (!fast) {                                 // (the fast path has no distance below 1e-10)
            if (fabsf(dSx) < 1.e-10f) dSx = g.xN;
            if (fabsf(dSy) < 1.e-10f) dSy = g.yN;
            if (fabsf(dSz) < 1.e-10f) dSz = g.zN;
        }
        dSx = fabsf(dSx); dSy = fabsf(dSy); dSz = fabsf(dSz);
        float dS = fminf(fminf(dSx, dSy), dSz);
        if (__builtin_expect(!fast && dS <= 0.f, 0)) {   // :1404-1432, only if an axis end coordinate is <= 0
            if (dSx <= 0.f)      dS = fminf(dSy, dSz);
            else if (dSy <= 0.f) dS = fminf(dSx, dSz);
            else                 dS = fminf(dSx, dSy);
            if (dS <= 0.f) { fail(L, 64); return; }
        }

        float tauCell = dS * opac;
        const bool interacts = (L.absTau + tauCell > L.passProb) && (cell > 0);
        // path length inside this cell: up to the interaction point or to the wall
        float dlLoc = dS;
        if (interacts) dlLoc = (L.passProb - L.absTau) / opac;
        j_add(g, L, cell, tix, dlLoc, aggEarly);

        if (interacts) {
            // ---- interaction (:1517-1814) ----
            L.rx = L.rx + dlLoc * L.vx;
            L.ry = L.ry + dlLoc * L.vy;
            L.rz = L.rz + dlLoc * L.vz;
            if (!(L.rx >= 0.f || L.rx < 0.f) || !(L.ry >= 0.f || L.ry < 0.f) || !(L.rz >= 0.f || L.rz < 0.f)) { fail(L, 65); return; }
            if (sym() && L.gP == 1) {
                if (L.rx <= g.x1) { L.vx = fabsf(L.vx); L.rx = g.x1; }
                if (L.ry <= g.y1) { L.vy = fabsf(L.vy); L.ry = g.y1; }
                if (L.rz <= g.z1) { L.vz = fabsf(L.vz); L.rz = g.z1; }
            }
            if (P.R_out > 0.f) {                          // :1577
                float tx = L.rx / 1.e10f, ty = L.ry / 1.e10f, tz = L.rz / 1.e10f;
                float rr = sqrtf(tx * tx + ty * ty + tz * tz) * 1.e10f;
                if (rr >= P.R_out) { escape(L, FATE_ESCAPED); return; }
            }
            if (P.lgDust) {
                float probSca = __ldg(&g.scaOpac[tix]) / opac;
                float random = 1.f - uni(L);
                bool scattered;
                if (random > probSca) scattered = false;
                else if (random <= probSca) scattered = true;
                else { fail(L, 68); return; }
                if (!scattered) {
                    absorb(L, P.lgGas ? CH_DIFFUSE : CH_DUSTEMI);
                    return;
                }
                count(C_SCA);
                if (kInc) cell = active_at<DENSE>(g, L.xP, L.yP, L.zP);
                if (!__ldg(&g.canScatter[cell])) { fail(L, 69); return; }
                if (L.istep >= a.P.safeLimit) { finish(L, FATE_DROPPED); return; }   // loop ends: :2838
                L.phase = PH_SCATTER;
                return;
            }
            if (!P.lgGas) { fail(L, 72); return; }
            absorb(L, CH_DIFFUSE);
            return;
        }

        // ---- no interaction in this cell (:1817-2731) ----
        L.absTau = L.absTau + tauCell;
        L.rx = L.rx + dS * L.vx;
        L.ry = L.ry + dS * L.vy;
        L.rz = L.rz + dS * L.vz;

        if (MULTI && L.gP > 1) track_mother(L);

        // :1961-1976
        // The reference's if / else-if chain on (dS == dSa .and. vHat_a > 0 / < 0).  An axis whose
        // distance is the (finite) minimum is moving, so v_a > 0 <=> pos_a from the wall stage;
        // only a packet with no moving axis at all (dS = 1e35, impossible for a unit vector)
        // needs the literal chain.
        if (dS < 1.e35f) {
            bool ex = dS == dSx, ey = (dS == dSy) & !ex, ez = (dS == dSz) & !ex & !ey;
            int ix = ex ? (posx ? 1 : -1) : 0, iy = ey ? (posy ? 1 : -1) : 0, iz = ez ? (posz ? 1 : -1) : 0;
            L.xP += ix; L.yP += iy; L.zP += iz;
            if (kInc) L.planeBase += (unsigned long long)(long long)(ix * (g.ny * g.nz) + iy * g.nz + iz);
        } else {
            if (kInc) L.planeG = 0;
            if (dS == dSx && L.vx > 0.f) L.xP = L.xP + 1;
            else if (dS == dSx && L.vx < 0.f) L.xP = L.xP - 1;
            else if (dS == dSy && L.vy > 0.f) L.yP = L.yP + 1;
            else if (dS == dSy && L.vy < 0.f) L.yP = L.yP - 1;
            else if (dS == dSz && L.vz > 0.f) L.zP = L.zP + 1;
            else if (dS == dSz && L.vz < 0.f) L.zP = L.zP - 1;
        }

        if (plane()) {
            if (kInc) L.planeG = 0;
            if (!step_tail_plane(L)) return;
        } else if (!MULTI) {
            // single grid: every test of :1986-2194, :2417-2540 and :2733-2834 that is true
            // ends in the same escape tally, so they collapse into one predicate
            bool out = (L.rx >= g.xHi) | (L.ry >= g.yHi) | (L.rz >= g.zHi) |
                       (L.xP > g.nx) | (L.yP > g.ny) | (L.zP > g.nz);
            bool low = (L.rx <= g.xLo) | (L.ry <= g.yLo) | (L.rz <= g.zLo) |
                       (L.xP < 1) | (L.yP < 1) | (L.zP < 1);
            out = out | (low & !sym());
            if (out) { escape(L, FATE_ESCAPED); return; }
            if (sym()) {                 // :2674-2699
                if (L.rx <= g.x1 || L.xP < 1) { L.vx = fabsf(L.vx); L.xP = 1; L.rx = g.x1; if (kInc) L.planeG = 0; }
                if (L.ry <= g.y1 || L.yP < 1) { L.vy = fabsf(L.vy); L.yP = 1; L.ry = g.y1; if (kInc) L.planeG = 0; }
                if (L.rz <= g.z1 || L.zP < 1) { L.vz = fabsf(L.vz); L.zP = 1; L.rz = g.z1; if (kInc) L.planeG = 0; }
            }
        } else {
            if (!step_tail_multi(L)) return;
        }
        if (LIMIT && __builtin_expect(L.istep >= a.P.safeLimit, 0)) finish(L, FATE_DROPPED);   // :2838-2846
    }

    // plane-parallel tail of a non-interacting step (:2199-2414, then :2703-2726): mirror at
    // the x and z faces, escape through the y faces; false = packet left FLY
    __device__ __forceinline__ bool step_tail_plane(Lane &L)
    {
        bool lgReturn = false;
        const DevGrid &m = G(1);
        {
            const DevGrid &c = G(L.gP);
            if (L.ry <= c.yLo || L.yP < 1) {
                if (L.gP == 1) { L.yP = 1; lgReturn = true; }
                else to_mother(L);
            }
        }
        {
            const DevGrid &c = G(L.gP);
            if (L.ry > c.yHi || L.yP > c.ny) {
                if (L.gP == 1) { L.yP = c.ny; lgReturn = true; }
                else to_mother(L);
            }
        }
        if (L.rx <= m.x1 || L.xP < 1) { L.xP = 1; L.rx = G(L.gP).x1; L.vx = -L.vx; }
        { const DevGrid &c = G(L.gP); if ((L.rx <= c.xLo || L.xP < 1) && L.gP > 1) to_mother(L); }
        {
            const DevGrid &c = G(L.gP);
            int im = c.nx <= m.nx ? c.nx : m.nx;                 // grid(1)%xAxis(grid(gP)%nx)
            if (L.rx >= __ldg(&m.xAxis[im - 1]) || L.xP > c.nx) { L.xP = c.nx; L.rx = c.xN; L.vx = -L.vx; }
        }
        { const DevGrid &c = G(L.gP); if ((L.rx >= c.xHi || L.xP > c.nx) && L.gP > 1) to_mother(L); }
        if (L.rz <= m.z1 || L.zP < 1) { L.zP = 1; L.rz = G(L.gP).y1; L.vz = -L.vz; }   // sic: yAxis(1), :2278
        { const DevGrid &c = G(L.gP); if ((L.rz <= c.zLo || L.zP < 1) && L.gP > 1) to_mother(L); }
        {
            const DevGrid &c = G(L.gP);
            int im = c.nz <= m.nz ? c.nz : m.nz;
            if (L.rz >= __ldg(&m.zAxis[im - 1]) || L.zP > c.nz) { L.zP = c.nz; L.rz = c.zN; L.vz = -L.vz; }
        }
        { const DevGrid &c = G(L.gP); if ((L.rz >= c.zHi || L.zP > c.nz) && L.gP > 1) to_mother(L); }
        if (lgReturn) { escape(L, FATE_ESCAPED); return false; }
        if (MULTI && L.gP > 1) {         // leaving the sub-grid (:2703-2726)
            const DevGrid &s = G(L.gP);
            if (((L.rx <= s.x1 || L.xP < 1) && L.vx <= 0.f) || ((L.ry <= s.y1 || L.yP < 1) && L.vy <= 0.f) ||
                ((L.rz <= s.z1 || L.zP < 1) && L.vz <= 0.f) || ((L.rx >= s.xN || L.xP > s.nx) && L.vx >= 0.f) ||
                ((L.ry >= s.yN || L.yP > s.ny) && L.vy >= 0.f) || ((L.rz >= s.zN || L.zP > s.nz) && L.vz >= 0.f)) to_mother(L);
        }
        return true;
    }
    __device__ __forceinline__ void to_mother(Lane &L)
    {
        L.xP = L.mx; L.yP = L.my; L.zP = L.mz;
        L.gP = 1;
        L.igpp = 0;
    }

    // multi-grid tail of a non-interacting step (:1986-2834); false = packet left FLY
    __device__ __forceinline__ bool step_tail_multi(Lane &L)
    {
        const DevParams &P = a.P;
        if (!sym()) {                  // "be 6/6/06" block (:1986-2194)
            bool lgReturn = false;
            {
                const DevGrid &c = G(L.gP);
                if (L.ry <= c.yLo || L.yP < 1) {
                    if (L.gP == 1) { L.yP = 1; lgReturn = true; }
                    else to_mother_stale(L);
                }
            }
            {
                const DevGrid &c = G(L.gP);
                if (L.ry > c.yHi || L.yP > c.ny) {
                    if (L.gP == 1) { L.yP = c.ny; lgReturn = true; }
                    else to_mother_stale(L);
                }
            }
            {
                const DevGrid &c = G(L.gP);
                if ((L.rx <= c.xLo || L.xP < 1) && L.gP == 1) { L.xP = 1; lgReturn = true; }
                if ((L.rx <= c.xLo || L.xP < 1) && L.gP > 1) to_mother_stale(L);
            }
            {
                const DevGrid &c = G(L.gP);
                if ((L.rx >= c.xHi || L.xP > c.nx) && L.gP == 1) { L.xP = c.nx; lgReturn = true; }
                if ((L.rx >= c.xHi || L.xP > c.nx) && L.gP > 1) to_mother_stale(L);
            }
            {
                const DevGrid &c = G(L.gP);
                if ((L.rz <= c.zLo || L.zP < 1) && L.gP == 1) { L.zP = 1; lgReturn = true; }
                if ((L.rz <= c.zLo || L.zP < 1) && L.gP > 1) to_mother_stale(L);
            }
            {
                const DevGrid &c = G(L.gP);
                if ((L.rz >= c.zHi || L.zP > c.nz) && L.gP == 1) { L.zP = c.nz; lgReturn = true; }
                if ((L.rz >= c.zHi || L.zP > c.nz) && L.gP > 1) to_mother_stale(L);
            }
            if (lgReturn) { escape(L, FATE_ESCAPED); return false; }
        }

        // still inside the simulation region? (:2417-2671)
        {
            const DevGrid &m = G(1);
            const DevGrid &c = G(L.gP);
            bool lowOut = !sym() && (L.rx <= m.xLo || L.ry <= m.yLo || L.rz <= m.zLo);
            if (lowOut || (L.rx >= c.xHi) || (L.ry >= c.yHi) || (L.rz >= c.zHi) ||
                L.xP > c.nx || L.yP > c.ny || L.zP > c.nz) {
                if (L.gP == 1) { escape(L, FATE_ESCAPED); return false; }
                L.xP = L.mx; L.yP = L.my; L.zP = L.mz;
                L.gP = 1;
                L.igpp = 0;
                float tx = L.rx / 1.e10f, ty = L.ry / 1.e10f, tz = L.rz / 1.e10f;
                float radius = 1.e10f * sqrtf(tx * tx + ty * ty + tz * tz);
                if ((radius >= P.R_out && P.R_out >= 0.f) ||
                    (L.rx >= m.xHi) || (L.ry >= m.yHi) || (L.rz >= m.zHi) || lowOut) {
                    escape(L, FATE_ESCAPED);
                    return false;
                }
            }
        }

        if (sym()) {                   // :2674-2699
            const DevGrid &m = G(1);
            const DevGrid &c = G(L.gP);
            if (L.rx <= m.x1 || (L.gP == 1 && L.xP < 1)) { L.vx = fabsf(L.vx); L.mx = 1; L.xP = 1; L.rx = c.x1; }
            if (L.ry <= m.y1 || (L.gP == 1 && L.yP < 1)) { L.vy = fabsf(L.vy); L.my = 1; L.yP = 1; L.ry = c.y1; }
            if (L.rz <= m.z1 || (L.gP == 1 && L.zP < 1)) { L.vz = fabsf(L.vz); L.mz = 1; L.zP = 1; L.rz = m.z1; }
        }

        if (L.gP > 1) {                  // leaving the sub-grid (:2703-2726)
            const DevGrid &s = G(L.gP);
            if (((L.rx <= s.x1 || L.xP < 1) && L.vx <= 0.f) ||
                ((L.ry <= s.y1 || L.yP < 1) && L.vy <= 0.f) ||
                ((L.rz <= s.z1 || L.zP < 1) && L.vz <= 0.f) ||
                ((L.rx >= s.xN || L.xP > s.nx) && L.vx >= 0.f) ||
                ((L.ry >= s.yN || L.yP > s.ny) && L.vy >= 0.f) ||
                ((L.rz >= s.zN || L.zP > s.nz) && L.vz >= 0.f)) {
                L.xP = L.mx; L.yP = L.my; L.zP = L.mz;
                L.gP = 1;
                L.igpp = 0;
            }
        }
        // :2733-2834
        if (L.gP == 1) {
            const DevGrid &m = G(1);
            if (L.xP > m.nx || L.yP > m.ny || L.zP > m.nz) { escape(L, FATE_ESCAPED); return false; }
        }
        return true;
    }

    // gP>1 branch of the 6/6/06 block: back to the mother grid *without* resetting igpp
    // (photon_mod.f90:1995-1999 etc.; motherP must be 1)
    __device__ __forceinline__ void to_mother_stale(Lane &L)
    {
        L.xP = L.mx; L.yP = L.my; L.zP = L.mz;
        L.gP = G(L.gP).motherP;
    }

    // keep track of the mother-grid cell while inside a sub-grid (:1842-1958)
    __device__ __forceinline__ void track_mother(Lane &L)
    {
        const DevGrid &m = G(G(L.gP).motherP);
        if (L.mx <= 0 || L.my <= 0 || L.mz <= 0 || L.mx > m.nx || L.my > m.ny || L.mz > m.nz) {
            // sic: a + b/2 instead of (a+b)/2 (:1854-1855)
            L.mx = locate_axis(m.xAxis, m.nx, L.rx);
            if (L.mx < m.nx && L.mx >= 1) { if (L.rx > (__ldg(&m.xAxis[L.mx - 1]) + __ldg(&m.xAxis[L.mx]) / 2.f)) L.mx = L.mx + 1; }
            L.my = locate_axis(m.yAxis, m.ny, L.ry);
            if (L.my < m.ny && L.my >= 1) { if (L.ry > (__ldg(&m.yAxis[L.my - 1]) + __ldg(&m.yAxis[L.my]) / 2.f)) L.my = L.my + 1; }
            L.mz = locate_axis(m.zAxis, m.nz, L.rz);
            if (L.mz < m.nz && L.mz >= 1) { if (L.rz > (__ldg(&m.zAxis[L.mz - 1]) + __ldg(&m.zAxis[L.mz]) / 2.f)) L.mz = L.mz + 1; }
        } else {
            if (L.vx > 0.f) { if (L.mx < m.nx) { if (L.rx > __ldg(&m.xWall[L.mx])) L.mx = L.mx + 1; } }
            else            { if (L.mx > 1)    { if (L.rx <= __ldg(&m.xWall[L.mx - 1])) L.mx = L.mx - 1; } }
            if (L.vy > 0.f) { if (L.my < m.ny) { if (L.ry > __ldg(&m.yWall[L.my])) L.my = L.my + 1; } }
            else            { if (L.my > 1)    { if (L.ry <= __ldg(&m.yWall[L.my - 1])) L.my = L.my - 1; } }
            if (L.vz > 0.f) { if (L.mz < m.nz) { if (L.rz > __ldg(&m.zWall[L.mz])) L.mz = L.mz + 1; } }
            else            { if (L.mz > 1)    { if (L.rz <= __ldg(&m.zWall[L.mz - 1])) L.mz = L.mz - 1; } }
        }
    }
};


// per-CTA shared-memory scratch: [C_COUNT][kThreads] event counters + [nbins] Qphot histogram
// + [(nbins+32)/32] bitmap of the frequency bins packets were emitted in
__host__ __device__ __forceinline__ int scratch_words(int nbins) { return C_COUNT * kThreads + nbins + (nbins + 32) / 32; }
__device__ __forceinline__ void scratch_init(unsigned int *smem, int nbins)
{
    for (int i = threadIdx.x; i < scratch_words(nbins); i += kThreads) smem[i] = 0u;
    __syncthreads();
}
__device__ __forceinline__ void scratch_flush(const TransportArgs &a, unsigned int *smem)
{
    unsigned int *cnt = smem, *qph = smem + C_COUNT * kThreads;
    __syncthreads();
    // per counter: warp shuffle reduction of the 256 per-thread slots, one atomic per warp
#pragma unroll
    for (int c = 0; c < C_COUNT; ++c) {
        unsigned int v = cnt[c * kThreads + threadIdx.x];
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31u) == 0u && v) atomicAdd(&a.counters[c], (unsigned long long)v);
    }
    for (int i = threadIdx.x; i < a.P.nbins; i += kThreads)
        if (qph[i]) atomicAdd(&a.qphotCounts[i], (unsigned long long)qph[i]);
    // a packet tallies in whatever grid it flies through: flag the bin in every grid
    const unsigned int *bits = qph + a.P.nbins;
    for (int i = threadIdx.x; i <= a.P.nbins; i += kThreads) {
        if (bits[i >> 5] & (1u << (i & 31))) {
            if (a.P.nGrids == 1) a.g1.nuTouched[i] = 1;
            else for (int g = 0; g < a.P.nGrids; ++g) a.grids[g].nuTouched[i] = 1;
        }
    }
}

}  // namespace mcb
