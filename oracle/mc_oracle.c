/*
 * oracle/mc_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the reference's energy-packet transport.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  Pinned against the reference's own code run through oracle/f90ref (see
 * mc_oracle.h, tests/test_reference_pin.py).
 *
 * Follows, routine by routine (all citations into /root/reference/source/):
 *   energyPacketDriver   photon_mod.f90:26-286    -> run_packet / oracle_transport
 *   energyPacketRun      photon_mod.f90:289-487   -> energy_packet_run
 *   initPhotonPacket     photon_mod.f90:491-684   -> init_photon_packet
 *   getNu2               photon_mod.f90:720-764   -> get_nu2
 *   newPhotonPacket      photon_mod.f90:768-1059  -> new_photon_packet
 *   pathSegment          photon_mod.f90:1061-2872 -> path_segment
 *   hg                   photon_mod.f90:2875-2974 -> hg
 *   randomUnitVector     vector_mod.f90:303-314   -> random_unit_vector
 *   locate               interpolation_mod.f90:48-81 -> locate
 * All arithmetic is IEEE float32 as in the reference (default REAL, Makefile has no
 * -fdefault-real-8); compile with -ffp-contract=off.
 *
 * Deliberate, documented differences from the reference:
 *  (1) RNG: the reference reseeds the Fortran intrinsic generator from the wall clock
 *      (photon_mod.f90:68-87), so it is irreproducible.  Here every packet owns a
 *      Philox4x32-10 stream keyed by (seed, global packet id, source index); a uniform
 *      is (word >> 8) * 2^-24 in [0,1) exactly like a 24-bit real(4) random_number.
 *  (2) log/sin/cos/acos/atan come from detmath.h (see there).
 *  (3) In newPhotonPacket('stellar') / ('diffExt') the unused slot of the local
 *      index arrays orX/orY/orZ is uninitialised stack memory in the reference
 *      (photon_mod.f90:781,841-843); here it is -1 (the value energyPacketDriver gives
 *      the unused slot of inX/inY/inZ, :107-113), which makes the mother-index
 *      tracker re-locate the packet (:1844-1871) instead of using garbage.
 *  (4) lg1D (dead: set_input_mod.f90:291-297 aborts) is not restated.
 *  (5) Besides the faithful float32 tallies the oracle also keeps order-independent
 *      integer tallies (path length in fixed point, packet counts); these are what the
 *      CUDA path is compared with bit for bit.
 */
#include "mc_oracle.h"
#include "detmath.h"

#include <stdlib.h>
#include <stdio.h>
#include <pthread.h>

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al. 2011), counter-based                          */
/* ------------------------------------------------------------------------- */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

typedef struct Rng {
    uint32_t k0, k1;        /* seed */
    uint32_t p0, p1;        /* packet id */
    uint32_t stream;        /* source index */
    uint32_t n;             /* draws so far */
    uint32_t buf[4];
} Rng;

static inline void rng_init(Rng *r, uint64_t seed, uint64_t pid, uint32_t stream)
{
    r->k0 = (uint32_t)seed; r->k1 = (uint32_t)(seed >> 32);
    r->p0 = (uint32_t)pid;  r->p1 = (uint32_t)(pid >> 32);
    r->stream = stream; r->n = 0;
}

static inline float rng_uniform(Rng *r)
{
    uint32_t lane = r->n & 3u;
    if (lane == 0) {
        r->buf[0] = r->p0; r->buf[1] = r->p1; r->buf[2] = r->n >> 2; r->buf[3] = r->stream;
        philox4x32_10(r->buf, r->k0, r->k1);
    }
    r->n++;
    return (float)(r->buf[lane] >> 8) * 5.9604644775390625e-08f; /* 2^-24 */
}

/* ------------------------------------------------------------------------- */
typedef struct { float x, y, z; } vec3;

typedef struct Packet {          /* type photon_packet, common_mod.f90:305-321 */
    int32_t nuP;
    int32_t xP[2], yP[2], zP[2]; /* [0]=mother, [1]=sub  (Fortran 1,2) */
    int32_t origin[2];           /* grid, cell */
    int32_t iG;
    float nu;
    int lgStellar, lgLine;
    vec3 position, direction;
} Packet;

enum { CH_STELLAR = 0, CH_DIFFEXT = 1, CH_DIFFUSE = 2, CH_DUSTEMI = 3 };

typedef struct Ctx {
    const OrParams *P;
    OrGrid *grids;
    int32_t iStar;
    float deltaE;
    OrCounters *C;
    int64_t *qphotCounts;
    Rng rng;
    int atomicMode;              /* 1: integer tallies with atomic adds, no fp32 tallies */
    int64_t segs;                /* segments of the current packet */
    int fateCode;
} Ctx;

#define ERR_STOP(code) do { return -(code); } while (0)

static inline int32_t ACTIVE(const OrGrid *g, int x, int y, int z)
{
    return g->active[(size_t)(x - 1) + (size_t)g->nx * ((size_t)(y - 1) + (size_t)g->ny * (size_t)(z - 1))];
}
static inline size_t T2(const OrGrid *g, int cell, int nu)
{
    return (size_t)(nu - 1) * (size_t)(g->nCells + 1) + (size_t)cell;
}

static inline void addq(Ctx *c, int64_t *p, int64_t v)
{
    if (c->atomicMode) __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
    else *p += v;
}

/* interpolation_mod.f90:48-81 */
static int32_t locate(const float *xa, int32_t n, float x)
{
    if (x > xa[n - 1]) return n;
    if (x < xa[0]) return 0;
    /* minloc((xa-x),1,(xa-x).gt.0): first location of the smallest positive xa-x */
    int32_t best = 0;
    float bestv = 0.f;
    for (int32_t i = 1; i <= n; ++i) {
        float d = xa[i - 1] - x;
        if (d > 0.f && (best == 0 || d < bestv)) { best = i; bestv = d; }
    }
    int32_t ns = best - 1;
    return ns > 1 ? ns : 1;
}

/* vector_mod.f90:303-314 */
static vec3 random_unit_vector(Rng *r)
{
    float r1 = rng_uniform(r);
    float w = 2.f * r1 - 1.f;
    float t = sqrtf(1.f - w * w);
    float r2 = rng_uniform(r);
    float ang = 3.141592654f * (2.f * r2 - 1.f);
    float sn, cs;
    dm_sincosf(ang, &sn, &cs);
    vec3 v;
    v.x = t * cs;
    v.y = t * sn;
    v.z = w;
    return v;
}

/* photon_mod.f90:720-764.  getNu2 takes the scan bound and the "+1" rule from the MODULE
 * variable nbins whatever the length of the row it is given; for linePDF(cell,:) (nLines
 * entries, :931) the scan can only stop inside the row because the row ends with 1 > random,
 * but the rule stays "nuP < nbins-1" (pinned against the translated reference,
 * tests/test_reference_pin.py).  nscan = entries that exist, nbins = the module's nbins. */
static int32_t get_nu2_row(Rng *r, const float *probDen, size_t stride, int32_t nscan, int32_t nbins)
{
    float random = rng_uniform(r);
    int i;
    for (i = 1; i <= 10000; ++i) {
        if (random == 0.f || random == 1.f || random == 0.9999999f) random = rng_uniform(r);
        else break;
    }
    int32_t nuP = 1;
    for (int32_t is = 1; is <= nscan && is <= nbins; ++is) {
        if (random >= probDen[(size_t)(is - 1) * stride]) nuP = is;
        else break;
    }
    if (nuP < nbins - 1) nuP = nuP + 1;
    if (nuP > nscan) nuP = nscan;      /* the reference would index past the row here */
    return nuP;
}

static int32_t get_nu2(Rng *r, const float *probDen, size_t stride, int32_t nbins)
{
    float random = rng_uniform(r);
    int i;
    for (i = 1; i <= 10000; ++i) {
        if (random == 0.f || random == 1.f || random == 0.9999999f) random = rng_uniform(r);
        else break;
    }
    int32_t nuP = 1;
    for (int32_t is = 1; is <= nbins; ++is) {
        if (random >= probDen[(size_t)(is - 1) * stride]) nuP = is;
        else break;
    }
    if (nuP < nbins - 1) nuP = nuP + 1;
    return nuP;
}

/* escape-direction bins, the block repeated at photon_mod.f90:373-412 (and 6 more) */
static int escape_bins(const OrParams *P, vec3 d, int32_t *pT, int32_t *pP)
{
    int32_t idirT, idirP;
    if (P->lgSymmetricXYZ) idirT = (int32_t)(dm_acosf(fabsf(d.z)) / P->dTheta) + 1;
    else                   idirT = (int32_t)(dm_acosf(d.z) / P->dTheta) + 1;
    if (idirT > P->totAngleBinsTheta) idirT = P->totAngleBinsTheta;
    if (idirT < 1 || idirT > P->totAngleBinsTheta) return -1;
    if (fabsf(d.x) < 1.e-35f) idirP = 0;
    else if (P->lgSymmetricXYZ) idirP = (int32_t)(dm_atanf(fabsf(d.y) / fabsf(d.x)) / P->dPhi);
    else                        idirP = (int32_t)(dm_atanf(d.y / d.x) / P->dPhi);
    if (idirP < 0) idirP = P->totAngleBinsPhi + idirP;
    idirP = idirP + 1;
    if (idirP > P->totAngleBinsPhi) idirP = P->totAngleBinsPhi;
    if (idirP < 1 || idirP > P->totAngleBinsPhi) return -2;
    *pT = idirT; *pP = idirP;
    return 0;
}

static inline void esc_add(Ctx *c, OrGrid *g, int32_t cell, int32_t nuP, int32_t ang)
{
    size_t idx = (size_t)cell + (size_t)(g->nCells + 1) * ((size_t)nuP + (size_t)(c->P->nbins + 1) * (size_t)ang);
    if (g->escapedPackets && !c->atomicMode) g->escapedPackets[idx] = g->escapedPackets[idx] + c->deltaE;
    if (g->escapedQ) addq(c, &g->escapedQ[idx], 1);
}

/* the escape tally repeated at photon_mod.f90:414-462,1633-1666,2152-2189,2372-2409,
 * 2498-2537,2625-2660,2793-2828 */
static int escape_tally(Ctx *c, const Packet *p)
{
    const OrParams *P = c->P;
    int32_t idirT, idirP;
    if (escape_bins(P, p->direction, &idirT, &idirP)) ERR_STOP(10);
    if (p->origin[0] < 1 || p->origin[0] > P->nGrids) ERR_STOP(11);
    if (p->origin[1] < 0) ERR_STOP(12);
    OrGrid *g = &c->grids[p->origin[0] - 1];
    int32_t cell = p->origin[1], nuP = p->nuP;
    if (P->nAngleBins > 0) {
        int32_t vt = P->viewPointPtheta[idirT], vp = P->viewPointPphi[idirP];
        if (vt > 0 && P->viewPointPhi[vt] < 0.f) {
            esc_add(c, g, cell, nuP, vt);
            esc_add(c, g, cell, nuP, 0);
        } else if (vt == vp || P->viewPointTheta[vp] == P->viewPointTheta[vt] ||
                   P->viewPointPhi[vt] == P->viewPointPhi[vp]) {
            esc_add(c, g, cell, nuP, vt);
            if (vt != 0) esc_add(c, g, cell, nuP, 0);
        } else {
            esc_add(c, g, cell, nuP, 0);
        }
    } else {
        esc_add(c, g, cell, nuP, 0);
    }
    c->C->nEscaped++;
    return 0;
}

/* photon_mod.f90:491-684 */
static int init_photon_packet(Ctx *c, Packet *pk, int32_t nuP, vec3 position, vec3 direction,
                              int lgLine, int lgStellar, const int32_t xP[2], const int32_t yP[2],
                              const int32_t zP[2], int32_t gP, int lgHG)
{
    const OrParams *P = c->P;
    pk->direction = direction;
    if (!(direction.x >= 0.f || direction.x < 0.f)) ERR_STOP(20);
    pk->position = position;
    pk->iG = gP;
    int igpi;
    if (gP == 1) igpi = 0; else if (gP > 1) igpi = 1; else ERR_STOP(21);
    pk->nuP = nuP;
    pk->lgStellar = lgStellar;
    if (lgLine) { pk->nu = 0.f; pk->lgLine = 1; }
    else        { pk->nu = P->nuArray[nuP - 1]; pk->lgLine = 0; }
    for (int k = 0; k < 2; ++k) { pk->xP[k] = xP[k]; pk->yP[k] = yP[k]; pk->zP[k] = zP[k]; }

    if (!lgHG || P->lgIsotropic || pk->lgStellar) {
        if (pk->lgStellar && P->lgPlaneIonization) {
            /* plane-parallel ionisation: emission from the y=0 face, :561-646 */
            const OrGrid *g = &c->grids[gP - 1];
            const float *xa = g->xAxis - 1, *za = g->zAxis - 1;
            float random = rng_uniform(&c->rng);
            random = 1.f - random;
            pk->position.x = -(xa[2] - xa[1]) / 2.f + random * ((xa[2] - xa[1]) / 2.f + (xa[g->nx] - xa[g->nx - 1]) / 2.f + xa[g->nx]);
            if (pk->position.x < xa[1]) pk->position.x = xa[1];
            if (pk->position.x > xa[g->nx]) pk->position.x = xa[g->nx];
            pk->xP[igpi] = locate(g->xAxis, g->nx, pk->position.x);
            if (pk->xP[igpi] < g->nx) {
                if (pk->xP[igpi] >= 1 && pk->position.x >= (xa[pk->xP[igpi]] + xa[pk->xP[igpi] + 1]) / 2.f) pk->xP[igpi] = pk->xP[igpi] + 1;
            }
            pk->position.y = 0.f;
            pk->yP[igpi] = 1;
            random = rng_uniform(&c->rng);
            random = 1.f - random;
            pk->position.z = -(za[2] - za[1]) / 2.f + random * ((za[2] - za[1]) / 2.f + (za[g->nz] - za[g->nz - 1]) / 2.f + za[g->nz]);
            if (pk->position.z < za[1]) pk->position.z = za[1];
            if (pk->position.z > za[g->nz]) pk->position.z = za[g->nz];
            pk->zP[igpi] = locate(g->zAxis, g->nz, pk->position.z);
            if (pk->zP[igpi] < g->nz) {
                /* sic: xAxis(zP) in the z test, :612 (guarded against indexing outside xAxis) */
                int zi = pk->zP[igpi];
                if (zi >= 1 && zi <= g->nx && pk->position.z >= (xa[zi] + za[zi + 1]) / 2.f) pk->zP[igpi] = zi + 1;
            }
            if (pk->xP[igpi] < 1) pk->xP[igpi] = 1;
            if (pk->zP[igpi] < 1) pk->zP[igpi] = 1;
            pk->direction.x = 0.f; pk->direction.y = 1.f; pk->direction.z = 0.f;
            if (P->planeIonDistribution) {
                int32_t *d = &P->planeIonDistribution[(pk->xP[igpi] - 1) + (size_t)c->grids[0].nx * (pk->zP[igpi] - 1)];
                if (c->atomicMode) __atomic_fetch_add(d, 1, __ATOMIC_RELAXED); else *d += 1;
            }
        } else {
            int irepeat;
            for (irepeat = 1; irepeat <= 1000000; ++irepeat) {
                pk->direction = random_unit_vector(&c->rng);
                if (!(pk->direction.x >= 0.f || pk->direction.x < 0.f)) ERR_STOP(23);
                if (pk->direction.x != 0.f && pk->direction.y != 0.f && pk->direction.z != 0.f) break;
            }
        }
        if (P->lgSymmetricXYZ && pk->lgStellar && !P->lgMultistars) {
            if (pk->direction.x < 0.f) pk->direction.x = -pk->direction.x;
            if (pk->direction.y < 0.f) pk->direction.y = -pk->direction.y;
            if (pk->direction.z < 0.f) pk->direction.z = -pk->direction.z;
        }
    }
    pk->origin[0] = gP;
    {
        const OrGrid *g = &c->grids[gP - 1];
        int x = pk->xP[igpi], y = pk->yP[igpi], z = pk->zP[igpi];
        if (x < 1 || x > g->nx || y < 1 || y > g->ny || z < 1 || z > g->nz) ERR_STOP(24);
        pk->origin[1] = ACTIVE(g, x, y, z);
    }
    return 0;
}

/* photon_mod.f90:768-1059 */
static int new_photon_packet(Ctx *c, Packet *pk, int chType, vec3 position, const int32_t xP[2],
                             const int32_t yP[2], const int32_t zP[2], int32_t *gP,
                             const int32_t difSource[3])
{
    const OrParams *P = c->P;
    const vec3 nullUnitVector = { 1.f, 0.f, 0.f };   /* grid_mod.f90:51-53 */
    int igpn;
    int32_t nuP;
    int32_t orX[2] = { -1, -1 }, orY[2] = { -1, -1 }, orZ[2] = { -1, -1 };
    int rc;
    if (*gP == 1) igpn = 0; else if (*gP > 1) igpn = 1; else ERR_STOP(30);

    switch (chType) {
    case CH_STELLAR: {
        int iStar = c->iStar;
        const int32_t *si = &P->starIndeces[4 * (iStar - 1)];
        const float *sp = &P->starPosition[3 * (iStar - 1)];
        if (position.x != sp[0] || position.y != sp[1] || position.z != sp[2]) ERR_STOP(31);
        if (si[3] == 1) igpn = 0; else if (si[3] > 1) igpn = 1; else ERR_STOP(32);
        nuP = get_nu2(&c->rng, &P->inSpectrumProbDen[(size_t)iStar * P->nbins], 1, P->nbins);
        if (nuP >= P->nbins) ERR_STOP(33);
        if (nuP < 1) ERR_STOP(34);
        orX[igpn] = si[0]; orY[igpn] = si[1]; orZ[igpn] = si[2];
        if (ACTIVE(&c->grids[si[3] - 1], orX[igpn], orY[igpn], orZ[igpn]) < 0) ERR_STOP(35);
        vec3 spv = { sp[0], sp[1], sp[2] };
        rc = init_photon_packet(c, pk, nuP, spv, nullUnitVector, 0, 1, orX, orY, orZ, si[3], 0);
        if (rc) return rc;
        if (pk->nu > 1.f) {
            c->C->Qphot = c->C->Qphot + c->deltaE / (2.1799153e-11f * pk->nu);
            if (c->qphotCounts) {
                if (c->atomicMode) __atomic_fetch_add(&c->qphotCounts[nuP - 1], 1, __ATOMIC_RELAXED);
                else c->qphotCounts[nuP - 1]++;
            }
        }
        break;
    }
    case CH_DIFFEXT: {
        nuP = get_nu2(&c->rng, &P->inSpectrumProbDen[0], 1, P->nbins);
        if (nuP >= P->nbins) ERR_STOP(36);
        if (nuP < 1) ERR_STOP(37);
        const OrGrid *g = &c->grids[*gP - 1];
        vec3 positionLoc;
        positionLoc.x = g->xAxis[difSource[0] - 1];
        positionLoc.y = g->yAxis[difSource[1] - 1];
        positionLoc.z = g->zAxis[difSource[2] - 1];
        orX[igpn] = difSource[0]; orY[igpn] = difSource[1]; orZ[igpn] = difSource[2];
        if (ACTIVE(g, orX[igpn], orY[igpn], orZ[igpn]) < 0) ERR_STOP(38);
        rc = init_photon_packet(c, pk, nuP, positionLoc, nullUnitVector, 0, 0, orX, orY, orZ, *gP, 0);
        if (rc) return rc;
        break;
    }
    case CH_DIFFUSE: {
        if (!P->lgGas) ERR_STOP(39);
        const OrGrid *g = &c->grids[*gP - 1];
        int32_t cell = ACTIVE(g, xP[igpn], yP[igpn], zP[igpn]);
        if (cell <= 0) ERR_STOP(40);
        float random = rng_uniform(&c->rng);
        random = 1.f - random;
        if (random <= g->totalLines[cell]) {
            if (P->lgDebug) {
                /* the reference passes linePDF(cell,:) to getNu2, which scans 1..nbins and
                 * applies "nuP<nbins-1 -> nuP+1" with the module's nbins (photon_mod.f90:747-757) */
                nuP = get_nu2_row(&c->rng, &g->linePDF[cell], (size_t)(g->nCells + 1), P->nLines, P->nbins);
                if (nuP < 1) ERR_STOP(41);
            } else {
                nuP = 0;
            }
            rc = init_photon_packet(c, pk, nuP, position, nullUnitVector, 1, 0, xP, yP, zP, *gP, 0);
            if (rc) return rc;
        } else {
            nuP = get_nu2(&c->rng, &g->recPDF[cell], (size_t)(g->nCells + 1), P->nbins);
            if (nuP >= P->nbins) ERR_STOP(42);
            if (nuP < 1) ERR_STOP(43);
            rc = init_photon_packet(c, pk, nuP, position, nullUnitVector, 0, 0, xP, yP, zP, *gP, 0);
            if (rc) return rc;
        }
        break;
    }
    case CH_DUSTEMI: {
        if (!P->lgDust) ERR_STOP(44);
        if (P->lgGas) ERR_STOP(45);
        const OrGrid *g = &c->grids[*gP - 1];
        int32_t cell = ACTIVE(g, xP[igpn], yP[igpn], zP[igpn]);
        if (cell <= 0) ERR_STOP(46);
        nuP = get_nu2(&c->rng, &g->dustPDF[cell], (size_t)(g->nCells + 1), P->nbins);
        if (nuP >= P->nbins) ERR_STOP(47);
        if (nuP < 1) ERR_STOP(48);
        rc = init_photon_packet(c, pk, nuP, position, nullUnitVector, 0, 0, xP, yP, zP, *gP, 0);
        if (rc) return rc;
        break;
    }
    default:
        ERR_STOP(49);
    }
    return 0;
}

/* photon_mod.f90:2875-2974.  Returns ierr; writes pk->direction only on the first branch. */
static int hg(Ctx *c, Packet *pk)
{
    float vin[3] = { pk->direction.x, pk->direction.y, pk->direction.z };
    float vout[3];
    float hgg = c->P->gSca[pk->nuP - 1];
    float random0 = rng_uniform(&c->rng);
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
    float random = rng_uniform(&c->rng);
    float phi = (2.f * 3.141592654f) * random;
    float cosp, sinp;
    dm_sincosf(phi, &sinp, &cosp);
    float denom = sqrtf(1.f - vin[2] * vin[2]);
    if (denom > 0.001f) {
        vout[0] = sint / denom * (vin[0] * vin[2] * cosp - vin[1] * sinp) + vin[0] * cost;
        vout[1] = sint / denom * (vin[1] * vin[2] * cosp + vin[0] * sinp) + vin[1] * cost;
        vout[2] = -sint * cosp * denom + vin[2] * cost;
    } else {
        vout[0] = sint * cosp;
        vout[1] = sint * sinp;
        if (vin[2] >= 0.f) vout[2] = cost; else vout[2] = -cost;
    }
    int fin0 = (vout[0] >= 0.f || vout[0] < 0.f);
    int fin1 = (vout[1] >= 0.f || vout[1] < 0.f);
    int fin2 = (vout[2] >= 0.f || vout[2] < 0.f);
    if ((fabsf(vout[0]) <= 1.f && fabsf(vout[1]) <= 1.f && fabsf(vout[2]) <= 1.f) && fin0 && fin1 && fin2) {
        pk->direction.x = vout[0];
        pk->direction.y = vout[1];
        pk->direction.z = vout[2];
        return 0;
    } else if ((fabsf(vout[0]) >= 1.f) || (fabsf(vout[1]) >= 1.f) ||
               ((fabsf(vout[2]) >= 1.f) && fin0 && fin1 && fin2)) {
        /* reference renormalises a local copy and never stores it (:2951-2952) */
        return 0;
    }
    return 1;
}

/* cell volume in 1e45 cm^3, photon_mod.f90:1469-1512 (== grid_mod.f90:2876-2965) */
static float cell_volume(const OrParams *P, const OrGrid *g, int xP, int yP, int zP)
{
    float dx = 0.f, dy = 0.f, dz = 0.f;
    const float *xa = g->xAxis - 1, *ya = g->yAxis - 1, *za = g->zAxis - 1; /* 1-based */
    if (xP > 1 && xP < g->nx) dx = fabsf(xa[xP + 1] - xa[xP - 1]) / 2.f;
    else if (xP == 1) { if (P->lgSymmetricXYZ) dx = fabsf(xa[xP + 1] - xa[xP]) / 2.f; else dx = fabsf(xa[xP + 1] - xa[xP]); }
    else if (xP == g->nx) dx = fabsf(xa[xP] - xa[xP - 1]);
    if (yP > 1 && yP < g->ny) dy = fabsf(ya[yP + 1] - ya[yP - 1]) / 2.f;
    else if (yP == 1) { if (P->lgSymmetricXYZ) dy = fabsf(ya[yP + 1] - ya[yP]) / 2.f; else dy = fabsf(ya[yP + 1] - ya[yP]); }
    else if (yP == g->ny) dy = fabsf(ya[yP] - ya[yP - 1]);
    if (zP > 1 && zP < g->nz) dz = fabsf(za[zP + 1] - za[zP - 1]) / 2.f;
    else if (zP == 1) { if (P->lgSymmetricXYZ) dz = fabsf(za[zP + 1] - za[zP]) / 2.f; else dz = fabsf(za[zP + 1] - za[zP]); }
    else if (zP == g->nz) dz = fabsf(za[zP] - za[zP - 1]);
    dx = dx / 1.e15f;
    dy = dy / 1.e15f;
    dz = dz / 1.e15f;
    return dx * dy * dz;
}

/* J estimator add, photon_mod.f90:1563-1574 and :1822-1833 */
static inline void j_add(Ctx *c, OrGrid *g, const Packet *pk, int32_t cell, float len, float dV)
{
    size_t idx = T2(g, cell, pk->nuP);
    int toDif = (!pk->lgStellar) && c->P->lgDebug;
    if (!c->atomicMode) {
        float *J = toDif ? g->Jdif : g->Jste;
        if (J) J[idx] = J[idx] + len * c->deltaE / dV;
    }
    int64_t *Q = toDif ? g->JdifQ : g->JsteQ;
    if (Q) addq(c, &Q[idx], (int64_t)llrintf(len * g->invLenUnit));
    int32_t *N = toDif ? g->JdifN : g->JsteN;
    if (N) N[idx] += 1;                        /* single-threaded runs only (test aid) */
}

#define PS_RETURN   0   /* packet finished (escaped / dropped) */
#define PS_REEMIT   1   /* absorbed: re-emit (reRun = 1) */

/* photon_mod.f90:1061-2872 */
static int path_segment(Ctx *c, Packet *enPacket, int *chTypeIn, vec3 *positionIn,
                        int32_t inX[2], int32_t inY[2], int32_t inZ[2], int32_t *gPIn)
{
    const OrParams *P = c->P;
    OrGrid *grid = c->grids - 1;              /* grid[1..nGrids] */
    vec3 vHat, rVec;
    float absTau, dlLoc, dSx, dSy, dSz, dS, dV, passProb, probSca, radius, random, tauCell;
    int32_t xP, yP, zP, gP;
    int igpp;
    int packetType = -1;
    int lgReturn;
    int safeLimit;
    int i, j;
    int rc;

    c->C->nFlights++;

    if (enPacket->iG == 1) igpp = 0; else if (enPacket->iG > 1) igpp = 1; else ERR_STOP(50);
    if (enPacket->iG <= 0 || enPacket->iG > P->nGrids) ERR_STOP(51);
    if (enPacket->xP[igpp] <= 0 || enPacket->xP[igpp] > grid[enPacket->iG].nx) ERR_STOP(52);
    if (enPacket->yP[igpp] <= 0 || enPacket->yP[igpp] > grid[enPacket->iG].ny) ERR_STOP(53);
    if (enPacket->zP[igpp] <= 0 || enPacket->zP[igpp] > grid[enPacket->iG].nz) ERR_STOP(54);

    rVec = enPacket->position;
    vHat = enPacket->direction;
    if (!(rVec.x >= 0.f || rVec.x < 0.f)) ERR_STOP(55);
    if (!(vHat.x >= 0.f || vHat.x < 0.f)) ERR_STOP(56);

    xP = enPacket->xP[igpp];
    yP = enPacket->yP[igpp];
    zP = enPacket->zP[igpp];
    gP = enPacket->iG;

    dSx = 0.f; dSy = 0.f; dSz = 0.f;
    absTau = 0.f;
    random = rng_uniform(&c->rng);
    passProb = -dm_logf(1.f - random);

    if (P->lgPlaneIonization) safeLimit = 5000; else safeLimit = 500000;   /* :1187-1192 */

    for (i = 1; i <= safeLimit; ++i) {
        c->segs++;
        for (j = 1; j <= safeLimit; ++j) {
            if (xP > grid[gP].nx || xP < 1 || yP > grid[gP].ny || yP < 1 ||
                zP > grid[gP].nz || zP < 1) ERR_STOP(58);

            if (ACTIVE(&grid[gP], xP, yP, zP) < 0) {
                /* packet is entering a subgrid, :1206-1243 */
                enPacket->xP[0] = xP;
                enPacket->yP[0] = yP;
                enPacket->zP[0] = zP;
                gP = abs(ACTIVE(&grid[gP], xP, yP, zP));
                if (gP < 1 || gP > P->nGrids) ERR_STOP(59);
                const float *xa = grid[gP].xAxis - 1, *ya = grid[gP].yAxis - 1, *za = grid[gP].zAxis - 1;
                xP = locate(grid[gP].xAxis, grid[gP].nx, rVec.x);
                if (xP == 0) xP = xP + 1;
                if (xP < grid[gP].nx) { if (rVec.x > (xa[xP + 1] + xa[xP]) / 2.f) xP = xP + 1; }
                yP = locate(grid[gP].yAxis, grid[gP].ny, rVec.y);
                if (yP == 0) yP = yP + 1;
                if (yP < grid[gP].ny) { if (rVec.y > (ya[yP + 1] + ya[yP]) / 2.f) yP = yP + 1; }
                zP = locate(grid[gP].zAxis, grid[gP].nz, rVec.z);
                if (zP == 0) zP = zP + 1;
                if (zP < grid[gP].nz) { if (rVec.z > (za[zP + 1] + za[zP]) / 2.f) zP = zP + 1; }
                enPacket->iG = gP;
                igpp = 1;
            }

            const OrGrid *g = &grid[gP];
            const float *xa = g->xAxis - 1, *ya = g->yAxis - 1, *za = g->zAxis - 1;

            if (P->lgSymmetricXYZ) {           /* :1248-1261 */
                if (rVec.x <= grid[1].xAxis[0]) { if (vHat.x < 0.f) vHat.x = -vHat.x; rVec.x = grid[1].xAxis[0]; }
                if (rVec.y <= grid[1].yAxis[0]) { if (vHat.y < 0.f) vHat.y = -vHat.y; rVec.y = grid[1].yAxis[0]; }
                if (rVec.z <= grid[1].zAxis[0]) { if (vHat.z < 0.f) vHat.z = -vHat.z; rVec.z = grid[1].zAxis[0]; }
            }

            /* x walls, :1263-1295 */
            if (vHat.x > 1.e-10f) {
                if (xP < g->nx) {
                    dSx = ((xa[xP + 1] + xa[xP]) / 2.f - rVec.x) / vHat.x;
                    if (fabsf(dSx) < 1.e-10f) { rVec.x = (xa[xP + 1] + xa[xP]) / 2.f; xP = xP + 1; }
                } else {
                    dSx = (xa[g->nx] - rVec.x) / vHat.x;
                    if (fabsf(dSx) < 1.e-10f) { rVec.x = xa[g->nx]; if (!P->lgPlaneIonization && gP == 1) { c->fateCode = 3; return PS_RETURN; } }
                }
            } else if (vHat.x < -1.e-10f) {
                if (xP > 1) {
                    dSx = ((xa[xP] + xa[xP - 1]) / 2.f - rVec.x) / vHat.x;
                    if (fabsf(dSx) < 1.e-10f) { rVec.x = (xa[xP] + xa[xP - 1]) / 2.f; xP = xP - 1; }
                } else {
                    dSx = (xa[1] - rVec.x) / vHat.x;
                    if (fabsf(dSx) < 1.e-10f) rVec.x = xa[1];
                }
            } else {
                dSx = 1.e35f;
            }
            if (!(dSx >= 0.f || dSx < 0.f)) ERR_STOP(60);

            /* y walls, :1304-1334 */
            if (vHat.y > 1.e-10f) {
                if (yP < g->ny) {
                    dSy = ((ya[yP + 1] + ya[yP]) / 2.f - rVec.y) / vHat.y;
                    if (fabsf(dSy) < 1.e-10f) { rVec.y = (ya[yP + 1] + ya[yP]) / 2.f; yP = yP + 1; }
                } else {
                    dSy = (ya[g->ny] - rVec.y) / vHat.y;
                    if (fabsf(dSy) < 1.e-10f) { rVec.y = ya[g->ny]; if (gP == 1) { c->fateCode = 3; return PS_RETURN; } }
                }
            } else if (vHat.y < -1.e-10f) {
                if (yP > 1) {
                    dSy = ((ya[yP] + ya[yP - 1]) / 2.f - rVec.y) / vHat.y;
                    if (fabsf(dSy) < 1.e-10f) { rVec.y = (ya[yP] + ya[yP - 1]) / 2.f; yP = yP - 1; }
                } else {
                    dSy = (ya[1] - rVec.y) / vHat.y;
                    if (fabsf(dSy) < 1.e-10f) rVec.y = ya[1];
                }
            } else {
                dSy = 1.e35f;
            }
            if (!(dSy >= 0.f || dSy < 0.f)) ERR_STOP(61);

            /* z walls, :1343-1373 */
            if (vHat.z > 1.e-10f) {
                if (zP < g->nz) {
                    dSz = ((za[zP + 1] + za[zP]) / 2.f - rVec.z) / vHat.z;
                    if (fabsf(dSz) < 1.e-10f) { rVec.z = (za[zP + 1] + za[zP]) / 2.f; zP = zP + 1; }
                } else {
                    dSz = (za[g->nz] - rVec.z) / vHat.z;
                    if (fabsf(dSz) < 1.e-10f) { rVec.z = za[g->nz]; if (!P->lgPlaneIonization && gP == 1) { c->fateCode = 3; return PS_RETURN; } }
                }
            } else if (vHat.z < -1.e-10f) {
                if (zP > 1) {
                    dSz = ((za[zP] + za[zP - 1]) / 2.f - rVec.z) / vHat.z;
                    if (fabsf(dSz) < 1.e-10f) { rVec.z = (za[zP] + za[zP - 1]) / 2.f; zP = zP - 1; }
                } else {
                    dSz = (za[1] - rVec.z) / vHat.z;
                    if (fabsf(dSz) < 1.e-10f) rVec.z = za[1];
                }
            } else {
                dSz = 1.e35f;
            }
            if (!(dSz >= 0.f || dSz < 0.f)) ERR_STOP(62);

            if (xP > g->nx || xP < 1 || yP > g->ny || yP < 1 || zP > g->nz || zP < 1) ERR_STOP(63);

            if (ACTIVE(g, xP, yP, zP) >= 0) break;
        }

        OrGrid *g = &grid[gP];
        const float *xa = g->xAxis - 1, *ya = g->yAxis - 1, *za = g->zAxis - 1;

        /* cater for cells on cell wall, :1395-1397 */
        if (fabsf(dSx) < 1.e-10f) dSx = xa[g->nx];
        if (fabsf(dSy) < 1.e-10f) dSy = ya[g->ny];
        if (fabsf(dSz) < 1.e-10f) dSz = za[g->nz];

        dSx = fabsf(dSx);
        dSy = fabsf(dSy);
        dSz = fabsf(dSz);

        if (dSx <= 0.f)      dS = fminf(dSy, dSz);
        else if (dSy <= 0.f) dS = fminf(dSx, dSz);
        else if (dSz <= 0.f) dS = fminf(dSx, dSy);
        else { dS = fminf(dSx, dSy); dS = fminf(dS, dSz); }

        if (dS <= 0.f) ERR_STOP(64);

        int32_t cell = ACTIVE(g, xP, yP, zP);
        float opac = g->opacity[T2(g, cell, enPacket->nuP)];
        tauCell = dS * opac;

        dV = cell_volume(P, g, xP, yP, zP);

        if ((absTau + tauCell > passProb) && (cell > 0)) {
            /* packet interacts, :1517-1814 */
            dlLoc = (passProb - absTau) / opac;
            rVec.x = rVec.x + dlLoc * vHat.x;
            rVec.y = rVec.y + dlLoc * vHat.y;
            rVec.z = rVec.z + dlLoc * vHat.z;
            if (!(rVec.x >= 0.f || rVec.x < 0.f)) ERR_STOP(65);
            if (!(rVec.y >= 0.f || rVec.y < 0.f)) ERR_STOP(66);
            if (!(rVec.z >= 0.f || rVec.z < 0.f)) ERR_STOP(67);

            if (P->lgSymmetricXYZ && gP == 1) {
                if (rVec.x <= xa[1]) { if (vHat.x < 0.f) vHat.x = -vHat.x; rVec.x = xa[1]; }
                if (rVec.y <= ya[1]) { if (vHat.y < 0.f) vHat.y = -vHat.y; rVec.y = ya[1]; }
                if (rVec.z <= za[1]) { if (vHat.z < 0.f) vHat.z = -vHat.z; rVec.z = za[1]; }
            }

            j_add(c, g, enPacket, cell, dlLoc, dV);

            {
                float tx = rVec.x / 1.e10f, ty = rVec.y / 1.e10f, tz = rVec.z / 1.e10f;
                float rr = sqrtf(tx * tx + ty * ty + tz * tz) * 1.e10f;
                if (rr >= P->R_out && P->R_out > 0.f) {
                    rc = escape_tally(c, enPacket);
                    if (rc) return rc;
                    c->fateCode = 1;
                    return PS_RETURN;
                }
            }

            if (P->lgDust) {
                probSca = g->scaOpac[T2(g, cell, enPacket->nuP)] / opac;
                random = rng_uniform(&c->rng);
                random = 1.f - random;
                int lgScattered;
                if (random > probSca) lgScattered = 0;
                else if (random <= probSca) lgScattered = 1;
                else ERR_STOP(68);

                if (!lgScattered) {
                    c->C->absInt = c->C->absInt + 1.f;
                    c->C->nAbs++;
                    if (!P->lgGas) packetType = CH_DUSTEMI;
                    else packetType = CH_DIFFUSE;
                    break;
                } else {
                    c->C->scaInt = c->C->scaInt + 1.f;
                    c->C->nSca++;
                    /* sublimation check, :1722-1748 */
                    int comp = P->lgMultiDustChemistry ? g->dustAbunIndex[cell] : 1;
                    int nSp = P->nSpeciesPart[comp - 1];
                    int nS;
                    for (nS = 1; nS <= nSp; ++nS) {
                        float ab = P->grainAbun[(comp - 1) + P->nDustComp * (nS - 1)];
                        float Td = g->Tdust[nS + (P->nSpeciesMax + 1) * (0 + (size_t)(P->nSizes + 1) * cell)];
                        if (ab > 0.f && Td < P->TdustSublime[P->dustComPoint[comp - 1] - 1 + nS - 1]) break;
                    }
                    if (nS > nSp) ERR_STOP(69);

                    enPacket->xP[igpp] = xP;
                    enPacket->yP[igpp] = yP;
                    enPacket->zP[igpp] = zP;
                    if (ACTIVE(g, enPacket->xP[igpp], enPacket->yP[igpp], enPacket->zP[igpp]) < 0) ERR_STOP(70);

                    {
                        Packet np;
                        rc = init_photon_packet(c, &np, enPacket->nuP, rVec, enPacket->direction, 0, 0,
                                                enPacket->xP, enPacket->yP, enPacket->zP, gP, 1);
                        if (rc) return rc;
                        *enPacket = np;
                    }
                    if (!P->lgIsotropic && !enPacket->lgStellar) {
                        int ihg;
                        for (ihg = 1; ihg <= 10; ++ihg) {
                            if (hg(c, enPacket) == 0) break;
                        }
                    }
                    vHat.x = enPacket->direction.x;
                    vHat.y = enPacket->direction.y;
                    vHat.z = enPacket->direction.z;
                    if (!(enPacket->direction.x >= 0.f || enPacket->direction.x < 0.f)) ERR_STOP(71);
                    absTau = 0.f;
                    random = rng_uniform(&c->rng);
                    passProb = -dm_logf(1.f - random);
                }
            } else {
                c->C->absInt = c->C->absInt + 1.f;
                c->C->nAbs++;
                if (!P->lgGas) ERR_STOP(72);
                packetType = CH_DIFFUSE;
                break;
            }
        } else {
            /* the packet is not absorbed within this cell, :1817-2731 */
            j_add(c, g, enPacket, cell, dS, dV);
            absTau = absTau + tauCell;
            rVec.x = rVec.x + dS * vHat.x;
            rVec.y = rVec.y + dS * vHat.y;
            rVec.z = rVec.z + dS * vHat.z;

            /* keep track of where you are on mother grid, :1842-1958 */
            if (gP > 1) {
                const OrGrid *m = &grid[g->motherP];
                const float *mx = m->xAxis - 1, *my = m->yAxis - 1, *mz = m->zAxis - 1;
                if (enPacket->xP[0] <= 0 || enPacket->yP[0] <= 0 || enPacket->zP[0] <= 0 ||
                    enPacket->xP[0] > m->nx || enPacket->yP[0] > m->ny || enPacket->zP[0] > m->nz) {
                    enPacket->xP[0] = locate(m->xAxis, m->nx, rVec.x);
                    if (enPacket->xP[0] < m->nx) {
                        /* sic: a + b/2, not (a+b)/2 (:1854-1855); xP(1)=0 would index xAxis(0): guarded */
                        if (enPacket->xP[0] >= 1 && rVec.x > (mx[enPacket->xP[0]] + mx[enPacket->xP[0] + 1] / 2.f))
                            enPacket->xP[0] = enPacket->xP[0] + 1;
                    }
                    enPacket->yP[0] = locate(m->yAxis, m->ny, rVec.y);
                    if (enPacket->yP[0] < m->ny) {
                        if (enPacket->yP[0] >= 1 && rVec.y > (my[enPacket->yP[0]] + my[enPacket->yP[0] + 1] / 2.f))
                            enPacket->yP[0] = enPacket->yP[0] + 1;
                    }
                    enPacket->zP[0] = locate(m->zAxis, m->nz, rVec.z);
                    if (enPacket->zP[0] < m->nz) {
                        if (enPacket->zP[0] >= 1 && rVec.z > (mz[enPacket->zP[0]] + mz[enPacket->zP[0] + 1] / 2.f))
                            enPacket->zP[0] = enPacket->zP[0] + 1;
                    }
                } else {
                    if (vHat.x > 0.f) {
                        if (enPacket->xP[0] < m->nx) {
                            if (rVec.x > (mx[enPacket->xP[0]] + mx[enPacket->xP[0] + 1]) / 2.f) enPacket->xP[0] = enPacket->xP[0] + 1;
                        }
                    } else {
                        if (enPacket->xP[0] > 1) {
                            if (rVec.x <= (mx[enPacket->xP[0] - 1] + mx[enPacket->xP[0]]) / 2.f) enPacket->xP[0] = enPacket->xP[0] - 1;
                        }
                    }
                    if (vHat.y > 0.f) {
                        if (enPacket->yP[0] < m->ny) {
                            if (rVec.y > (my[enPacket->yP[0]] + my[enPacket->yP[0] + 1]) / 2.f) enPacket->yP[0] = enPacket->yP[0] + 1;
                        }
                    } else {
                        if (enPacket->yP[0] > 1) {
                            if (rVec.y <= (my[enPacket->yP[0] - 1] + my[enPacket->yP[0]]) / 2.f) enPacket->yP[0] = enPacket->yP[0] - 1;
                        }
                    }
                    if (vHat.z > 0.f) {
                        if (enPacket->zP[0] < m->nz) {
                            if (rVec.z > (mz[enPacket->zP[0]] + mz[enPacket->zP[0] + 1]) / 2.f) enPacket->zP[0] = enPacket->zP[0] + 1;
                        }
                    } else {
                        if (enPacket->zP[0] > 1) {
                            if (rVec.z <= (mz[enPacket->zP[0] - 1] + mz[enPacket->zP[0]]) / 2.f) enPacket->zP[0] = enPacket->zP[0] - 1;
                        }
                    }
                }
            }

            /* :1961-1976 */
            if (dS == dSx && vHat.x > 0.f) xP = xP + 1;
            else if (dS == dSx && vHat.x < 0.f) xP = xP - 1;
            else if (dS == dSy && vHat.y > 0.f) yP = yP + 1;
            else if (dS == dSy && vHat.y < 0.f) yP = yP - 1;
            else if (dS == dSz && vHat.z > 0.f) zP = zP + 1;
            else if (dS == dSz && vHat.z < 0.f) zP = zP - 1;
            /* else: reference only prints a warning */

            /* be 6/6/06, :1986-2194 */
            if (!P->lgPlaneIonization && !P->lgSymmetricXYZ) {
                lgReturn = 0;
                if (rVec.y <= grid[gP].yAxis[0] - grid[gP].geoCorrY || yP < 1) {
                    if (gP == 1) { yP = 1; lgReturn = 1; }
                    else if (gP > 1) {
                        int mp = grid[gP].motherP - 1;
                        xP = enPacket->xP[mp]; yP = enPacket->yP[mp]; zP = enPacket->zP[mp];
                        gP = grid[gP].motherP;
                    } else ERR_STOP(73);
                }
                if (rVec.y > grid[gP].yAxis[grid[gP].ny - 1] + grid[gP].geoCorrY || yP > grid[gP].ny) {
                    if (gP == 1) { yP = grid[gP].ny; lgReturn = 1; }
                    else if (gP > 1) {
                        int mp = grid[gP].motherP - 1;
                        xP = enPacket->xP[mp]; yP = enPacket->yP[mp]; zP = enPacket->zP[mp];
                        gP = grid[gP].motherP;
                    } else ERR_STOP(74);
                }
                if ((rVec.x <= grid[gP].xAxis[0] - grid[gP].geoCorrX || xP < 1) && gP == 1) { xP = 1; lgReturn = 1; }
                if ((rVec.x <= grid[gP].xAxis[0] - grid[gP].geoCorrX || xP < 1) && gP > 1) {
                    int mp = grid[gP].motherP - 1;
                    xP = enPacket->xP[mp]; yP = enPacket->yP[mp]; zP = enPacket->zP[mp];
                    gP = grid[gP].motherP;
                }
                if ((rVec.x >= grid[gP].xAxis[grid[gP].nx - 1] + grid[gP].geoCorrX || xP > grid[gP].nx) && gP == 1) { xP = grid[gP].nx; lgReturn = 1; }
                if ((rVec.x >= grid[gP].xAxis[grid[gP].nx - 1] + grid[gP].geoCorrX || xP > grid[gP].nx) && gP > 1) {
                    int mp = grid[gP].motherP - 1;
                    xP = enPacket->xP[mp]; yP = enPacket->yP[mp]; zP = enPacket->zP[mp];
                    gP = grid[gP].motherP;
                }
                if ((rVec.z <= grid[gP].zAxis[0] - grid[gP].geoCorrZ || zP < 1) && gP == 1) { zP = 1; lgReturn = 1; }
                if ((rVec.z <= grid[gP].zAxis[0] - grid[gP].geoCorrZ || zP < 1) && gP > 1) {
                    int mp = grid[gP].motherP - 1;
                    xP = enPacket->xP[mp]; yP = enPacket->yP[mp]; zP = enPacket->zP[mp];
                    gP = grid[gP].motherP;
                }
                if ((rVec.z >= grid[gP].zAxis[grid[gP].nz - 1] + grid[gP].geoCorrZ || zP > grid[gP].nz) && gP == 1) { zP = grid[gP].nz; lgReturn = 1; }
                if ((rVec.z >= grid[gP].zAxis[grid[gP].nz - 1] + grid[gP].geoCorrZ || zP > grid[gP].nz) && gP > 1) {
                    int mp = grid[gP].motherP - 1;
                    xP = enPacket->xP[mp]; yP = enPacket->yP[mp]; zP = enPacket->zP[mp];
                    gP = grid[gP].motherP;
                }
                if (lgReturn) {
                    rc = escape_tally(c, enPacket);
                    if (rc) return rc;
                    c->fateCode = 1;
                    return PS_RETURN;
                }
            }

            if (P->lgPlaneIonization) {        /* :2199-2414 */
                lgReturn = 0;
#define TO_MOTHER_PLANE() do { xP = enPacket->xP[0]; yP = enPacket->yP[0]; zP = enPacket->zP[0]; gP = 1; igpp = 0; } while (0)
                if (rVec.y <= grid[gP].yAxis[0] - grid[gP].geoCorrY || yP < 1) {
                    if (gP == 1) { yP = 1; lgReturn = 1; }
                    else if (gP > 1) TO_MOTHER_PLANE();
                    else ERR_STOP(77);
                }
                if (rVec.y > grid[gP].yAxis[grid[gP].ny - 1] + grid[gP].geoCorrY || yP > grid[gP].ny) {
                    if (gP == 1) { yP = grid[gP].ny; lgReturn = 1; }
                    else if (gP > 1) TO_MOTHER_PLANE();
                    else ERR_STOP(78);
                }
                if (rVec.x <= grid[1].xAxis[0] || xP < 1) {
                    xP = 1;
                    rVec.x = grid[gP].xAxis[0];
                    vHat.x = -vHat.x;
                }
                if ((rVec.x <= grid[gP].xAxis[0] - grid[gP].geoCorrX || xP < 1) && gP > 1) TO_MOTHER_PLANE();
                {
                    int nxg = grid[gP].nx, im = nxg <= grid[1].nx ? nxg : grid[1].nx;   /* grid(1)%xAxis(grid(gP)%nx) */
                    if (rVec.x >= grid[1].xAxis[im - 1] || xP > nxg) {
                        xP = nxg;
                        rVec.x = grid[gP].xAxis[nxg - 1];
                        vHat.x = -vHat.x;
                    }
                }
                if ((rVec.x >= grid[gP].xAxis[grid[gP].nx - 1] + grid[gP].geoCorrX || xP > grid[gP].nx) && gP > 1) TO_MOTHER_PLANE();
                if (rVec.z <= grid[1].zAxis[0] || zP < 1) {
                    zP = 1;
                    rVec.z = grid[gP].yAxis[0];            /* sic: yAxis(1), :2278 */
                    vHat.z = -vHat.z;
                }
                if ((rVec.z <= grid[gP].zAxis[0] - grid[gP].geoCorrZ || zP < 1) && gP > 1) TO_MOTHER_PLANE();
                {
                    int nzg = grid[gP].nz, im = nzg <= grid[1].nz ? nzg : grid[1].nz;
                    if (rVec.z >= grid[1].zAxis[im - 1] || zP > nzg) {
                        zP = nzg;
                        rVec.z = grid[gP].zAxis[nzg - 1];
                        vHat.z = -vHat.z;
                    }
                }
                if ((rVec.z >= grid[gP].zAxis[grid[gP].nz - 1] + grid[gP].geoCorrZ || zP > grid[gP].nz) && gP > 1) TO_MOTHER_PLANE();
#undef TO_MOTHER_PLANE
                if (lgReturn) {
                    rc = escape_tally(c, enPacket);
                    if (rc) return rc;
                    c->fateCode = 1;
                    return PS_RETURN;
                }
            }

            /* :2417-2419 */
            {
                float tx = rVec.x / 1.e10f, ty = rVec.y / 1.e10f, tz = rVec.z / 1.e10f;
                radius = 1.e10f * sqrtf(tx * tx + ty * ty + tz * tz);
            }

            if (!P->lgPlaneIonization) {
                if ((!P->lgSymmetricXYZ && (rVec.x <= grid[1].xAxis[0] - grid[1].geoCorrX ||
                                            rVec.y <= grid[1].yAxis[0] - grid[1].geoCorrY ||
                                            rVec.z <= grid[1].zAxis[0] - grid[1].geoCorrZ)) ||
                    (rVec.x >= grid[gP].xAxis[grid[gP].nx - 1] + grid[gP].geoCorrX) ||
                    (rVec.y >= grid[gP].yAxis[grid[gP].ny - 1] + grid[gP].geoCorrY) ||
                    (rVec.z >= grid[gP].zAxis[grid[gP].nz - 1] + grid[gP].geoCorrZ) ||
                    xP > grid[gP].nx || yP > grid[gP].ny || zP > grid[gP].nz) {
                    if (gP == 1) {
                        /* :2433-2438 clamp local indices (dead: return follows) */
                        rc = escape_tally(c, enPacket);
                        if (rc) return rc;
                        c->fateCode = 1;
                        return PS_RETURN;
                    } else if (gP > 1) {
                        xP = enPacket->xP[0];
                        yP = enPacket->yP[0];
                        zP = enPacket->zP[0];
                        gP = 1;
                        igpp = 0;
                        if ((radius >= P->R_out && P->R_out >= 0.f) ||
                            (rVec.x >= grid[1].xAxis[grid[1].nx - 1] + grid[1].geoCorrX) ||
                            (rVec.y >= grid[1].yAxis[grid[1].ny - 1] + grid[1].geoCorrY) ||
                            (rVec.z >= grid[1].zAxis[grid[1].nz - 1] + grid[1].geoCorrZ) ||
                            (!P->lgSymmetricXYZ && (rVec.x <= grid[1].xAxis[0] - grid[1].geoCorrX ||
                                                    rVec.y <= grid[1].yAxis[0] - grid[1].geoCorrY ||
                                                    rVec.z <= grid[1].zAxis[0] - grid[1].geoCorrZ))) {
                            if (xP > grid[gP].nx) xP = grid[gP].nx;
                            if (yP > grid[gP].ny) yP = grid[gP].ny;
                            if (zP > grid[gP].nz) zP = grid[gP].nz;
                            if (xP < 1) xP = 1;
                            if (yP < 1) yP = 1;
                            if (zP < 1) zP = 1;
                            rc = escape_tally(c, enPacket);
                            if (rc) return rc;
                            c->fateCode = 1;
                            return PS_RETURN;
                        }
                    } else ERR_STOP(75);
                }

                if (P->lgSymmetricXYZ) {       /* :2674-2699 */
                    if (rVec.x <= grid[1].xAxis[0] || (gP == 1 && xP < 1)) {
                        if (vHat.x < 0.f) vHat.x = -vHat.x;
                        enPacket->xP[0] = 1;
                        xP = 1;
                        rVec.x = grid[gP].xAxis[0];
                    }
                    if (rVec.y <= grid[1].yAxis[0] || (gP == 1 && yP < 1)) {
                        if (vHat.y < 0.f) vHat.y = -vHat.y;
                        enPacket->yP[0] = 1;
                        yP = 1;
                        rVec.y = grid[gP].yAxis[0];
                    }
                    if (rVec.z <= grid[1].zAxis[0] || (gP == 1 && zP < 1)) {
                        if (vHat.z < 0.f) vHat.z = -vHat.z;
                        enPacket->zP[0] = 1;
                        zP = 1;
                        rVec.z = grid[1].zAxis[0];
                    }
                }
            }

            if (gP > 1) {                      /* :2703-2726 */
                const OrGrid *s = &grid[gP];
                if (((rVec.x <= s->xAxis[0] || xP < 1) && vHat.x <= 0.f) ||
                    ((rVec.y <= s->yAxis[0] || yP < 1) && vHat.y <= 0.f) ||
                    ((rVec.z <= s->zAxis[0] || zP < 1) && vHat.z <= 0.f) ||
                    ((rVec.x >= s->xAxis[s->nx - 1] || xP > s->nx) && vHat.x >= 0.f) ||
                    ((rVec.y >= s->yAxis[s->ny - 1] || yP > s->ny) && vHat.y >= 0.f) ||
                    ((rVec.z >= s->zAxis[s->nz - 1] || zP > s->nz) && vHat.z >= 0.f)) {
                    xP = enPacket->xP[0];
                    yP = enPacket->yP[0];
                    zP = enPacket->zP[0];
                    gP = 1;
                    igpp = 0;
                }
            }
        }

        /* :2733-2834 */
        if (!P->lgPlaneIonization && gP == 1 &&
            (xP > grid[gP].nx || yP > grid[gP].ny || zP > grid[gP].nz)) {
            rc = escape_tally(c, enPacket);
            if (rc) return rc;
            c->fateCode = 1;
            return PS_RETURN;
        }
    }

    if (i >= safeLimit) {                      /* :2838-2846 */
        c->fateCode = 3;
        return PS_RETURN;
    }

    if (gP == 1) igpp = 0; else if (gP > 1) igpp = 1; else ERR_STOP(76);
    enPacket->xP[igpp] = xP;
    enPacket->yP[igpp] = yP;
    enPacket->zP[igpp] = zP;

    *chTypeIn = packetType;
    *positionIn = rVec;
    for (int k = 0; k < 2; ++k) { inX[k] = enPacket->xP[k]; inY[k] = enPacket->yP[k]; inZ[k] = enPacket->zP[k]; }
    *gPIn = gP;
    return PS_REEMIT;
}

/* photon_mod.f90:289-487 */
static int energy_packet_run(Ctx *c, int *chType, vec3 *position, int32_t xP[2], int32_t yP[2],
                             int32_t zP[2], int32_t *gP, int *rR, int32_t *lastNuP)
{
    const OrParams *P = c->P;
    Packet enPacket;
    int igpr;
    int32_t difSourceL[3] = { -1, -1, -1 };
    const int32_t noCellLoc[3] = { -1, -1, -1 };
    int rc;
    *rR = 0;
    if (*gP == 1) igpr = 0; else if (*gP > 1) igpr = 1; else ERR_STOP(80);
    switch (*chType) {
    case CH_STELLAR:
        rc = new_photon_packet(c, &enPacket, *chType, *position, xP, yP, zP, gP, noCellLoc);
        break;
    case CH_DIFFEXT:
        difSourceL[0] = xP[igpr]; difSourceL[1] = yP[igpr]; difSourceL[2] = zP[igpr];
        rc = new_photon_packet(c, &enPacket, *chType, *position, xP, yP, zP, gP, difSourceL);
        break;
    case CH_DIFFUSE:
    case CH_DUSTEMI:
        rc = new_photon_packet(c, &enPacket, *chType, *position, xP, yP, zP, gP, noCellLoc);
        break;
    default:
        ERR_STOP(81);
    }
    if (rc) return rc;
    *lastNuP = enPacket.nuP;

    if (!P->lgDust && enPacket.nu < P->ionEdge1 && !enPacket.lgLine) {
        rc = escape_tally(c, &enPacket);       /* :370-465 */
        if (rc) return rc;
        c->C->nEarlyEscaped++;
        c->fateCode = 5;
        return 0;
    }
    if (!enPacket.lgLine) {
        rc = path_segment(c, &enPacket, chType, position, xP, yP, zP, gP);
        if (rc < 0) return rc;
        *rR = rc;                              /* PS_REEMIT -> reRun = 1 */
        return 0;
    } else {
        if (P->lgDebug) {                      /* :478-482 */
            OrGrid *g = &c->grids[*gP - 1];
            int32_t cell = ACTIVE(g, enPacket.xP[igpr], enPacket.yP[igpr], enPacket.zP[igpr]);
            size_t idx = T2(g, cell, enPacket.nuP);
            if (g->linePackets && !c->atomicMode) g->linePackets[idx] = g->linePackets[idx] + c->deltaE;
            if (g->linePacketsQ) addq(c, &g->linePacketsQ[idx], 1);
        }
        c->C->nLinePackets++;
        c->fateCode = 2;
    }
    return 0;
}

/* one trip of the packet loop, photon_mod.f90:93-170 */
static int run_packet(Ctx *c, int64_t pid, uint64_t seed, int32_t gpLoc, const int32_t *cellLoc, int32_t *fate)
{
    const OrParams *P = c->P;
    int32_t inX[2] = { -1, -1 }, inY[2] = { -1, -1 }, inZ[2] = { -1, -1 };
    int chTypeIn, reRun = 0, igp, i, rc;
    int32_t gPIn, lastNuP = 0;
    vec3 positionIn;

    /* Philox keying, same rule as the CUDA library (capi.cu run_transport): a star's packets are
     * (packet id, iStar); the extra diffuse source of cell (gpLoc, cellLoc) has streams of its own
     * -- counter word 3 = 0x80000000 + linear index of the emitting cell, grid in bits 48+ of the id */
    if (c->iStar >= 1 || !cellLoc || gpLoc < 1) {
        rng_init(&c->rng, seed, (uint64_t)pid, (uint32_t)c->iStar);
    } else {
        const OrGrid *dg = &c->grids[gpLoc - 1];
        uint32_t lin = (uint32_t)((cellLoc[0] - 1) + dg->nx * ((cellLoc[1] - 1) + dg->ny * (cellLoc[2] - 1)));
        rng_init(&c->rng, seed, ((uint64_t)gpLoc << 48) + (uint64_t)pid, 0x80000000u + lin);
    }
    c->segs = 0;
    c->fateCode = 0;

    if (c->iStar >= 1) {
        const int32_t *si = &P->starIndeces[4 * (c->iStar - 1)];
        chTypeIn = CH_STELLAR;
        if (si[3] == 1) igp = 0; else if (si[3] > 1) igp = 1; else ERR_STOP(90);
        inX[igp] = si[0]; inY[igp] = si[1]; inZ[igp] = si[2];
        positionIn.x = P->starPosition[3 * (c->iStar - 1) + 0];
        positionIn.y = P->starPosition[3 * (c->iStar - 1) + 1];
        positionIn.z = P->starPosition[3 * (c->iStar - 1) + 2];
        gPIn = si[3];
    } else {
        const OrGrid *g = &c->grids[gpLoc - 1];
        chTypeIn = CH_DIFFEXT;
        if (gpLoc == 1) igp = 0; else if (gpLoc > 1) igp = 1; else ERR_STOP(91);
        inX[igp] = cellLoc[0]; inY[igp] = cellLoc[1]; inZ[igp] = cellLoc[2];
        positionIn.x = g->xAxis[cellLoc[0] - 1];
        positionIn.y = g->yAxis[cellLoc[1] - 1];
        positionIn.z = g->zAxis[cellLoc[1] - 1];   /* sic: zAxis(cellLoc(2)), :150 */
        gPIn = gpLoc;
    }
    for (i = 1; i <= 5000; ++i) {              /* recursionLimit, constants_mod.f90:56 */
        rc = energy_packet_run(c, &chTypeIn, &positionIn, inX, inY, inZ, &gPIn, &reRun, &lastNuP);
        if (rc) return rc;
        if (reRun == 0) break;
    }
    if (i >= 5000) { c->C->trapped++; if (reRun) c->fateCode = 4; }
    if (c->fateCode == 3) c->C->nDropped++;
    c->C->nSegments += c->segs;
    if (fate) {
        fate[0] = (int32_t)c->segs;
        fate[1] = i > 5000 ? 5000 : i;
        fate[2] = lastNuP;
        fate[3] = c->fateCode;
    }
    return 0;
}

int oracle_transport(const OrParams *P, OrGrid *grids, int32_t iStar, int64_t firstId, int64_t n,
                     uint64_t seed, int32_t gpLoc, const int32_t *cellLoc, OrCounters *C,
                     int64_t *qphotCounts, int32_t *fate)
{
    Ctx c;
    memset(&c, 0, sizeof(c));
    c.P = P; c.grids = grids; c.iStar = iStar; c.C = C; c.qphotCounts = qphotCounts;
    c.deltaE = P->deltaE[iStar];
    c.atomicMode = 0;
    if (iStar < 0 || iStar > P->nStars) return -1;
    for (int64_t k = 0; k < n; ++k) {
        int rc = run_packet(&c, firstId + k, seed, gpLoc, cellLoc, fate ? fate + 4 * k : NULL);
        if (rc) return rc;
    }
    return 0;
}

/* photon_mod.f90:180-266 */
int oracle_transport_reslines(const OrParams *P, OrGrid *grids, int32_t iStar, uint64_t seed,
                              int32_t rank, int32_t nranks, OrCounters *C, int64_t *nRun)
{
    Ctx c;
    memset(&c, 0, sizeof(c));
    c.P = P; c.grids = grids; c.iStar = iStar; c.C = C; c.qphotCounts = NULL;
    c.deltaE = P->deltaE[iStar];
    int64_t iCell = 0, gid = 0, run = 0;
    for (int igrid = 1; igrid <= P->nGrids; ++igrid) {
        OrGrid *g = &grids[igrid - 1];
        int igp = igrid == 1 ? 0 : 1;
        for (int ix = 1; ix <= g->nx; ++ix)
            for (int iy = 1; iy <= g->ny; ++iy)
                for (int iz = 1; iz <= g->nz; ++iz) {
                    iCell = iCell + 1;
                    int32_t cell = ACTIVE(g, ix, iy, iz);
                    int32_t cnt = (cell > 0 && g->resLinePackets) ? g->resLinePackets[cell] : 0;
                    int mine = ((iCell - (rank + 1)) % nranks) == 0;
                    for (int iPhot = 1; iPhot <= cnt; ++iPhot, ++gid) {
                        if (!mine) continue;
                        int32_t inX[2] = { -1, -1 }, inY[2] = { -1, -1 }, inZ[2] = { -1, -1 };
                        vec3 pos = { g->xAxis[ix - 1], g->yAxis[iy - 1], g->zAxis[iz - 1] };
                        inX[igp] = ix; inY[igp] = iy; inZ[igp] = iz;
                        if (igrid > 1) {          /* location on the mother grid, :222-239 */
                            const OrGrid *m = &grids[g->motherP - 1];
                            inX[0] = locate(m->xAxis, m->nx, pos.x);
                            if (inX[0] >= 1 && inX[0] < m->nx && pos.x > (m->xAxis[inX[0] - 1] + m->xAxis[inX[0]]) / 2.f) inX[0] = inX[0] + 1;
                            inY[0] = locate(m->yAxis, m->ny, pos.y);
                            if (inY[0] >= 1 && inY[0] < m->ny && pos.y > (m->yAxis[inY[0] - 1] + m->yAxis[inY[0]]) / 2.f) inY[0] = inY[0] + 1;
                            inZ[0] = locate(m->zAxis, m->nz, pos.z);
                            if (inZ[0] >= 1 && inZ[0] < m->nz && pos.z > (m->zAxis[inZ[0] - 1] + m->zAxis[inZ[0]]) / 2.f) inZ[0] = inZ[0] + 1;
                        }
                        rng_init(&c.rng, seed, ((uint64_t)1 << 40) + (uint64_t)gid, (uint32_t)iStar);
                        c.segs = 0; c.fateCode = 0;
                        int chTypeIn = CH_DIFFUSE, reRun = 0, i, rc;
                        int32_t gPIn = igrid, lastNuP = 0;
                        vec3 positionIn = pos;
                        for (i = 1; i <= 5000; ++i) {
                            rc = energy_packet_run(&c, &chTypeIn, &positionIn, inX, inY, inZ, &gPIn, &reRun, &lastNuP);
                            if (rc) return rc;
                            if (reRun == 0) break;
                        }
                        if (i >= 5000) C->trapped++;
                        if (c.fateCode == 3) C->nDropped++;
                        C->nSegments += c.segs;
                        run++;
                    }
                }
    }
    if (nRun) *nRun = run;
    return 0;
}

/* ------------------------------------------------------------------------- */
typedef struct MtArg {
    const OrParams *P; OrGrid *grids; int32_t iStar; int64_t first, n; uint64_t seed;
    OrCounters C; int64_t *qphotCounts; int rc;
} MtArg;

static void *mt_worker(void *vp)
{
    MtArg *a = (MtArg *)vp;
    Ctx c;
    memset(&c, 0, sizeof(c));
    c.P = a->P; c.grids = a->grids; c.iStar = a->iStar; c.C = &a->C; c.qphotCounts = a->qphotCounts;
    c.deltaE = a->P->deltaE[a->iStar];
    c.atomicMode = 1;
    a->rc = 0;
    for (int64_t k = 0; k < a->n; ++k) {
        int rc = run_packet(&c, a->first + k, a->seed, 0, NULL, NULL);
        if (rc) { a->rc = rc; break; }
    }
    return NULL;
}

int oracle_transport_mt(const OrParams *P, OrGrid *grids, int32_t iStar, int64_t firstId, int64_t n,
                        uint64_t seed, int32_t nThreads, OrCounters *C, int64_t *qphotCounts)
{
    if (nThreads < 1) nThreads = 1;
    if (iStar < 1 || iStar > P->nStars) return -1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nThreads);
    MtArg *args = (MtArg *)calloc(nThreads, sizeof(MtArg));
    /* the reference's own split, iteration_mod.f90:477-493 */
    int64_t load = n / nThreads, rest = n % nThreads, first = firstId;
    for (int t = 0; t < nThreads; ++t) {
        int64_t mine = load + (t < rest ? 1 : 0);
        args[t].P = P; args[t].grids = grids; args[t].iStar = iStar; args[t].first = first;
        args[t].n = mine; args[t].seed = seed; args[t].qphotCounts = qphotCounts;
        first += mine;
        pthread_create(&th[t], NULL, mt_worker, &args[t]);
    }
    int rc = 0;
    for (int t = 0; t < nThreads; ++t) {
        pthread_join(th[t], NULL);
        if (args[t].rc) rc = args[t].rc;
        C->Qphot += args[t].C.Qphot; C->absInt += args[t].C.absInt; C->scaInt += args[t].C.scaInt;
        C->nAbs += args[t].C.nAbs; C->nSca += args[t].C.nSca; C->trapped += args[t].C.trapped;
        C->nLinePackets += args[t].C.nLinePackets; C->nDropped += args[t].C.nDropped;
        C->nSegments += args[t].C.nSegments; C->nFlights += args[t].C.nFlights;
        C->nEscaped += args[t].C.nEscaped; C->nEarlyEscaped += args[t].C.nEarlyEscaped;
    }
    free(th); free(args);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* opacity assembly: ionization_mod.f90:26-129,349-484 + iteration_mod.f90:166-227 */
/* ------------------------------------------------------------------------- */
static void in_opacity(float *opacity /*1-based*/, const float *xSecArray /*1-based*/, int nbins,
                       int xSecP, int nuLowP, int nuHighP, float den)
{
    /* inOpacity with b = 0 (ionization_mod.f90:448-482) */
    int k = xSecP - nuLowP;
    int iup = nuHighP < nbins ? nuHighP : nbins;
    if (iup < nuLowP) iup = nuLowP;
    for (int i = nuLowP; i <= iup; ++i) opacity[i] = opacity[i] + xSecArray[i + k] * den;
}

void oracle_opacity(const OrOpacityIn *in, float *opacityOut, float *scaOut, float *absOut)
{
    const int nR = in->nCells + 1, nb = in->nbins;
    const float *xs = in->xSecArray - 1;
    float *row = (float *)malloc(sizeof(float) * (nb + 2));
    float density[31][31];
    for (size_t i = 0; i < (size_t)nR * nb; ++i) opacityOut[i] = 0.f;
    for (int cell = 1; cell <= in->nCells; ++cell) {
        /* density(n,i), ionization_mod.f90:65-80 */
        memset(density, 0, sizeof(density));
        for (int n = 1; n <= 30; ++n) {
            int imax = n < in->nstages ? n : in->nstages;
            for (int i = 1; i <= imax; ++i) {
                if (!in->lgElementOn[n - 1]) break;
                int xr = in->elementXref[n - 1];
                float ion = in->ionDen[(size_t)cell + (size_t)nR * ((size_t)(xr - 1) + (size_t)in->nElementsUsed * (size_t)(i - 1))];
                float ab = in->elemAbun[(size_t)(in->abIndex[cell] - 1) + (size_t)in->nAbComp * (size_t)(n - 1)];
                density[n][i] = ion * ab * in->Hden[cell];
            }
        }
        for (int i = 0; i <= nb + 1; ++i) row[i] = 0.f;
        /* addOpacity, :349-443 */
        if (in->ff1) row[1] = row[1] + in->ff1[cell];
        in_opacity(row, xs, nb, in->HlevXSecP1, in->HlevNuP1, nb, density[1][1]);
        in_opacity(row, xs, nb, in->HeISingXSecP1, in->HeIlevNuP1, nb, density[2][1]);
        in_opacity(row, xs, nb, in->HeIIXSecP1, in->HeIIlevNuP1, nb, density[2][2]);
        for (int el = 3; el <= 30; ++el) {
            if (!in->lgElementOn[el - 1]) continue;
            int imax = el < in->nstages ? el : in->nstages;
            for (int ion = 1; ion <= imax; ++ion) {       /* putOpacity, :418-443 */
                if (density[el][ion] > 0.f) {
                    int ns = in->nShells[(el - 1) + 30 * (ion - 1)];
                    for (int sh = 1; sh <= ns; ++sh) {
                        size_t b = (size_t)(el - 1) + 30 * ((size_t)(ion - 1) + 30 * (size_t)(sh - 1));
                        int nuLowP = in->elementP[b + 0 * 30 * 30 * 7];
                        int nuHighP = in->elementP[b + 1 * 30 * 30 * 7];
                        int xSecP = in->elementP[b + 2 * 30 * 30 * 7];
                        in_opacity(row, xs, nb, xSecP, nuLowP, nuHighP, density[el][ion]);
                    }
                }
            }
        }
        for (int i = 1; i <= nb; ++i) opacityOut[(size_t)(i - 1) * nR + cell] = row[i];
    }
    free(row);
    if (!in->lgDust) return;
    /* dust contribution, iteration_mod.f90:166-227 */
    for (size_t i = 0; i < (size_t)nR * nb; ++i) { scaOut[i] = 0.f; absOut[i] = 0.f; }
    for (int cell = 1; cell <= in->nCells; ++cell) {
        int nsp = in->lgMultiDustChemistry ? in->dustAbunIndex[cell] : 1;
        if (nsp < 1 || nsp > in->nDustComp) continue;
        int dcp = in->dustComPoint[nsp - 1];
        for (int nS = 1; nS <= in->nSpeciesPart[nsp - 1]; ++nS) {
            for (int ai = 1; ai <= in->nSizes; ++ai) {
                float Td = in->Tdust[(size_t)nS + (size_t)(in->nSpeciesMax + 1) * ((size_t)ai + (size_t)(in->nSizes + 1) * (size_t)cell)];
                if (Td < in->TdustSublime[dcp - 1 + nS - 1]) {
                    float ga = in->grainAbun[(size_t)(nsp - 1) + (size_t)in->nDustComp * (size_t)(nS - 1)];
                    int sp = in->dustScaXsecP[(size_t)(nS + dcp - 1 - 1) + (size_t)in->nSpeciesTot * (size_t)(ai - 1)];
                    int ap = in->dustAbsXsecP[(size_t)(nS + dcp - 1 - 1) + (size_t)in->nSpeciesTot * (size_t)(ai - 1)];
                    for (int f = 1; f <= nb; ++f) {
                        size_t o = (size_t)(f - 1) * nR + cell;
                        scaOut[o] = scaOut[o] + ga * in->grainWeight[ai - 1] * in->Ndust[cell] * xs[sp + f - 1];
                        absOut[o] = absOut[o] + ga * in->grainWeight[ai - 1] * in->Ndust[cell] * xs[ap + f - 1];
                    }
                }
            }
        }
        for (int f = 1; f <= nb; ++f) {
            size_t o = (size_t)(f - 1) * nR + cell;
            opacityOut[o] = opacityOut[o] + (scaOut[o] + absOut[o]);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* dust-only closure: getFlux, setDustPDF, getDustT                            */
/* ------------------------------------------------------------------------- */
/* continuum_mod.f90:359-416, cShape 'blackbody' */
static float get_flux(float energy, float temperature)
{
    const float hPlanck = 6.6262e-27f, hcRyd_k = 157893.94f;
    float constant = 0.5250229f / hPlanck;
    if (hcRyd_k * energy / temperature > 86.f) {
        /* Wien: real*real*real*real * exp(dble(..)) -> double product, rounded on assignment */
        float pre = constant * energy * energy * energy;
        return (float)((double)pre * dm_exp_d((double)(-hcRyd_k * energy / temperature)));
    }
    float denominator = dm_expf(hcRyd_k * energy / temperature) - 1.f;
    if (denominator <= 0.f) return 3.32154e-6f * energy * energy * temperature / hPlanck;
    return constant * energy * energy * energy / denominator;
}
float oracle_get_flux(float energy, float temperature) { return get_flux(energy, temperature); }

#define TDUST(in, T, nS, ai, cell) (T)[(size_t)(nS) + (size_t)((in)->nSpeciesMax + 1) * ((size_t)(ai) + (size_t)((in)->nSizes + 1) * (size_t)(cell))]

void oracle_dust_pdf(const OrDustIn *in, const float *Tdust, float *dustPDF)
{
    const int nR = in->nCells + 1, nb = in->nbins;
    const float *xs = in->xSecArray - 1;
    for (size_t i = 0; i < (size_t)nR * nb; ++i) dustPDF[i] = 0.f;
    for (int cell = 1; cell <= in->nCells; ++cell) {
        int nspE = in->lgMultiDustChemistry ? in->dustAbunIndex[cell] : 1;
        if (nspE < 1 || nspE > in->nDustComp) continue;
        int dcp = in->dustComPoint[nspE - 1];
        for (int n = 1; n <= in->nSpeciesPart[nspE - 1]; ++n)
            for (int ai = 1; ai <= in->nSizes; ++ai) {
                float treal = TDUST(in, Tdust, n, ai, cell);
                if (treal > 0.f && treal < in->TdustSublime[dcp - 1 + n - 1]) {
                    int ap = in->dustAbsXsecP[(size_t)(n + dcp - 1 - 1) + (size_t)in->nSpeciesTot * (size_t)(ai - 1)];
                    float ga = in->grainAbun[(size_t)(nspE - 1) + (size_t)in->nDustComp * (size_t)(n - 1)];
                    for (int i = 1; i <= nb; ++i) {
                        float bb = get_flux(in->nuArray[i - 1], treal);
                        size_t o = (size_t)(i - 1) * nR + cell;
                        dustPDF[o] = dustPDF[o] + xs[ap + i - 1] * bb * in->widFlx[i - 1] * in->grainWeight[ai - 1] * ga;
                    }
                }
            }
        for (int i = 2; i <= nb; ++i) {
            size_t o = (size_t)(i - 1) * nR + cell;
            dustPDF[o] = dustPDF[o - nR] + dustPDF[o];
        }
        float last = dustPDF[(size_t)(nb - 1) * nR + cell];
        for (int i = 1; i <= nb; ++i) {
            size_t o = (size_t)(i - 1) * nR + cell;
            dustPDF[o] = dustPDF[o] / last;
        }
        dustPDF[(size_t)(nb - 1) * nR + cell] = 1.f;
    }
}

void oracle_dust_update(const OrDustIn *in, const float *Jste, const float *Jdif, float XHILimit,
                        float *Tdust, int32_t *lgConverged)
{
    const int nR = in->nCells + 1, nb = in->nbins, nT = in->nTemps;
    const float *xs = in->xSecArray - 1;
    const float Pi = 3.141592654f;
    float *radField = (float *)malloc(sizeof(float) * nb);
    float *row = (float *)malloc(sizeof(float) * nT);
    for (int cell = 1; cell <= in->nCells; ++cell) {
        int nspU = in->lgMultiDustChemistry ? in->dustAbunIndex[cell] : 1;
        /* updateCell returns before anything else for a cell no packet crossed
         * (update_mod.f90:104-149): Tdust and lgConverged keep their previous values */
        int lgHit = 0;
        for (int i = 0; i < nb && !lgHit; ++i) {
            size_t o = (size_t)i * nR + cell;
            lgHit = Jste[o] > 0.f || (in->lgDebug && Jdif && Jdif[o] > 0.f);
        }
        if (!lgHit) continue;
        float XOldHI = TDUST(in, Tdust, 0, 0, cell);
        for (int i = 0; i < nb; ++i) {
            size_t o = (size_t)i * nR + cell;
            radField[i] = (in->lgDebug && Jdif) ? (Jste[o] + Jdif[o]) / Pi : Jste[o] / Pi;
        }
        for (int nS = 0; nS <= in->nSpeciesMax; ++nS)
            for (int ai = 0; ai <= in->nSizes; ++ai) TDUST(in, Tdust, nS, ai, cell) = 0.f;
        if (nspU >= 1 && nspU <= in->nDustComp) {
            for (int nS = 1; nS <= in->nSpeciesPart[nspU - 1]; ++nS) {
                for (int ai = 1; ai <= in->nSizes; ++ai) {
                    float dustAbsIntegral = 0.f;
                    int ap = in->dustAbsXsecP[(size_t)(nS - 1) + (size_t)in->nSpeciesTot * (size_t)(ai - 1)];   /* sic: local nS */
                    for (int i = 1; i <= nb; ++i) dustAbsIntegral = dustAbsIntegral + xs[ap + i - 1] * radField[i - 1];
                    for (int t = 0; t < nT; ++t)
                        row[t] = in->dustEmIntegral[(size_t)(nS - 1) + (size_t)in->nSpeciesTot * ((size_t)(ai - 1) + (size_t)in->nSizes * (size_t)t)];
                    int iT = locate(row, nT, dustAbsIntegral);
                    float T;
                    if (iT <= 0) T = 1.f;
                    else if (iT >= nT) T = (float)nT;
                    else T = (float)iT + (dustAbsIntegral - row[iT - 1]) * ((float)(iT + 1) - (float)iT) / (row[iT] - row[iT - 1]);
                    TDUST(in, Tdust, nS, ai, cell) = T;
                    TDUST(in, Tdust, nS, 0, cell) = TDUST(in, Tdust, nS, 0, cell) + T * in->grainWeight[ai - 1];
                }
                float ga = in->grainAbun[(size_t)(nspU - 1) + (size_t)in->nDustComp * (size_t)(nS - 1)];
                TDUST(in, Tdust, 0, 0, cell) = TDUST(in, Tdust, 0, 0, cell) + TDUST(in, Tdust, nS, 0, cell) * ga;
            }
        }
        float deltaXHI = (TDUST(in, Tdust, 0, 0, cell) - XOldHI) / XOldHI;
        lgConverged[cell] = (fabsf(deltaXHI) <= XHILimit) ? 1 : 0;
    }
    free(radField); free(row);
}

/* ------------------------------------------------------------------------- */
/* photo-rate pre-integration, update_mod.f90:170-262 and :1160-1214           */
/* ------------------------------------------------------------------------- */
void oracle_photo_integrals(int32_t nCells, int32_t nbins, int32_t nBands, const int32_t *off,
                            const int32_t *low, const int32_t *high, const float *xSecArray,
                            const float *nuArray, const float *J, float *nPhoto, float *heat)
{
    const float hcRyd = 2.1799153e-11f;
    const size_t nR = (size_t)nCells + 1;
    const float *xs = xSecArray - 1, *nu = nuArray - 1;
    for (int b = 0; b < nBands; ++b) {
        int hi = high[b] < nbins ? high[b] : nbins;
        for (int cell = 0; cell <= nCells; ++cell) {
            float np = 1.e-20f, ht = 0.f;
            int heatOn = 1;
            for (int j = low[b]; j <= hi; ++j) {
                float phXSec = xs[off[b] + (j - low[b])];
                float Jc = J[(size_t)(j - 1) * nR + cell];
                if (phXSec < 1.e-35f) { heatOn = 0; phXSec = 0.f; }
                if (Jc > 0.f) {
                    np = np + Jc * phXSec / (hcRyd * nu[j]);
                    if (heatOn) ht = ht + phXSec * Jc * (nu[j] - nu[low[b]]) / nu[j];
                }
            }
            if (nPhoto) nPhoto[(size_t)b * nR + cell] = np;
            if (heat) heat[(size_t)b * nR + cell] = ht;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* unit-test hooks                                                            */
/* ------------------------------------------------------------------------- */
void oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out4)
{
    uint32_t c[4] = { c0, c1, c2, c3 };
    philox4x32_10(c, k0, k1);
    memcpy(out4, c, sizeof(c));
}

void oracle_uniforms(uint64_t seed, uint64_t pid, uint32_t stream, int32_t n, float *out)
{
    Rng r; rng_init(&r, seed, pid, stream);
    for (int i = 0; i < n; ++i) out[i] = rng_uniform(&r);
}

int32_t oracle_locate(const float *xa, int32_t n, float x) { return locate(xa, n, x); }

int32_t oracle_getnu2(const float *probDen, int64_t stride, int32_t nbins, uint64_t seed, uint64_t pid, uint32_t stream)
{
    Rng r; rng_init(&r, seed, pid, stream);
    return get_nu2(&r, probDen, (size_t)stride, nbins);
}

void oracle_random_unit_vector(uint64_t seed, uint64_t pid, uint32_t stream, float *out3)
{
    Rng r; rng_init(&r, seed, pid, stream);
    vec3 v = random_unit_vector(&r);
    out3[0] = v.x; out3[1] = v.y; out3[2] = v.z;
}

int32_t oracle_hg(float g, const float *vin, uint64_t seed, uint64_t pid, uint32_t stream, float *vout)
{
    OrParams P; memset(&P, 0, sizeof(P));
    float gs[1] = { g };
    P.gSca = gs;
    Ctx c; memset(&c, 0, sizeof(c));
    c.P = &P;
    rng_init(&c.rng, seed, pid, stream);
    Packet pk; memset(&pk, 0, sizeof(pk));
    pk.nuP = 1;
    pk.direction.x = vin[0]; pk.direction.y = vin[1]; pk.direction.z = vin[2];
    int ierr = hg(&c, &pk);
    vout[0] = pk.direction.x; vout[1] = pk.direction.y; vout[2] = pk.direction.z;
    return ierr;
}

void oracle_detmath(int32_t which, const float *in, float *out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        float s, c;
        switch (which) {
        case 0: out[i] = dm_logf(in[i]); break;
        case 1: dm_sincosf(in[i], &s, &c); out[i] = s; break;
        case 2: dm_sincosf(in[i], &s, &c); out[i] = c; break;
        case 3: out[i] = dm_acosf(in[i]); break;
        case 4: out[i] = dm_atanf(in[i]); break;
        case 5: out[i] = dm_expf(in[i]); break;
        default: out[i] = 0.f;
        }
    }
}

/* double-precision detmath (which: 0 = dm_exp_d, the Wien branch of getFlux) */
void oracle_detmath_d(int32_t which, const double *in, double *out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) out[i] = which == 0 ? dm_exp_d(in[i]) : 0.0;
}

int32_t oracle_escape_bins(const OrParams *P, const float *dir, int32_t *idirT, int32_t *idirP)
{
    vec3 d = { dir[0], dir[1], dir[2] };
    return escape_bins(P, d, idirT, idirP);
}

float oracle_cell_volume(const OrParams *P, const OrGrid *g, int32_t xP, int32_t yP, int32_t zP)
{
    return cell_volume(P, g, xP, yP, zP);
}
