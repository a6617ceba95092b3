// types.h -- device-side views of the reference's grid_type / common_mod globals.
#pragma once
#include <cstdint>

namespace mcb {

// Hot members of grid_type (common_mod.f90:241-302) as device pointers.
// Layout in HBM (all as in the reference unless noted):
//   active   int32 [x-1 + nx*((y-1) + ny*(z-1))]            (4 B / cell)
//   opacity, scaOpac float [(nu-1)*(nCells+1) + cell]       (nu-planes, cell fastest)
//   pdfT     float [cell*nbins + (nu-1)]                    TRANSPOSED on upload so one
//            re-emission CDF row is contiguous (binary search touches <=2 lines/probe)
//   xWall[i] i=0..nx: W[0]=x(1), W[i]=(x(i+1)+x(i))/2, W[nx]=x(nx)  -- the mid-point cell
//            walls of photon_mod.f90:1263-1295 precomputed with the same float32 expression
//   JsteQ/JdifQ  uint64 [(nu-1)*(nCells+1) + cell]  fixed-point path length (2^-e cm units)
//   escQ     uint32 [cell + (nCells+1)*(nu + (nbins+1)*ang)] escaped packet counts
//   lineQ    uint32 [(line-1)*(nCells+1) + cell]    line packet counts (debug)
struct DevGrid {
    int nx, ny, nz, nCells, motherP;
    int dense;                       // 1: every cell active, id = 1 + (z-1) + nz*((y-1) + ny*(x-1))
    float geoX, geoY, geoZ;          // geoCorr (grid_mod.f90:809-812)
    // per-axis constants of the escape tests, same float32 expressions as the reference:
    // a1 = axis(1), aN = axis(n), lo = axis(1)-geoCorr, hi = axis(n)+geoCorr
    float x1, y1, z1, xN, yN, zN;
    float xLo, yLo, zLo, xHi, yHi, zHi;
    float invLenUnit;                // 2^-e, path-length quantum of the J tally
    const float *xAxis, *yAxis, *zAxis;
    const float *xWall, *yWall, *zWall;
    const int *active;
    const float *opacity, *scaOpac;
    const float *pdfT;               // recPDF (gas) or dustPDF (dust only), transposed
    const float *totalLines;
    const float *linePDF;            // reference layout, debug only
    const unsigned char *canScatter; // per cell: an unsublimated species exists (photon_mod.f90:1722-1748)
    unsigned long long *JsteQ, *JdifQ;
    unsigned int *escQ, *lineQ;      // packet counts: < 2^32 per rank and call
    int *nuTouched;                  // [nbins+1] 1 = some packet was emitted in this bin since the last fold:
                                     // only those nu-planes of JsteQ/escQ can be non-zero (exchange + fold skip the rest)
};

struct DevParams {
    int nGrids, nbins, nStars, nAngleBins, totT, totP, nLines;
    int lgDust, lgGas, lgSym, lgIso, lgDebug, lgMultistars, lgPlane;
    int safeLimit;                   // photon_mod.f90:1187-1192: 500000, or 5000 in plane-parallel mode
    int *planeDist;                  // planeIonDistribution(grid(1)%nx, grid(1)%nz), plane mode only
    float dTheta, dPhi, R_out, ionEdge1;
    const float *nuArray, *gSca;
    const float *starCdf;            // [(s)*nbins + nu-1], s = 0..nStars
    const float *starPos;            // [3*(i-1)+k]
    const int *starIdx;              // [4*(i-1)+k]
    const int *starCell;             // active id of the star cell, per star
    const int *vpPtheta, *vpPphi;    // 0:totT, 0:totP
    const float *vpTheta, *vpPhi;    // 0:nAngleBins
};

enum Counter {
    C_ABS = 0, C_SCA, C_TRAPPED, C_LINE, C_DROPPED, C_SEGMENTS, C_FLIGHTS, C_ESCAPED, C_EARLY,
    C_COUNT
};

// one source cell of the resonance-line packet transfer (photon_mod.f90:180-266)
struct ResCell {
    int grid;                        // 1-based grid
    short x, y, z;                   // cell indices in that grid
    short mx, my, mz;                // mother-grid slot (sub-grids: nearest mother point, :222-239)
    float px, py, pz;                // cell centre
    unsigned long long gid;          // global index of the cell's first packet (all ranks, loop order)
};

struct TransportArgs {
    DevParams P;
    DevGrid g1;                      // copy of grids[0] (constant-bank access for the mother grid)
    const DevGrid *grids;            // device array [nGrids]
    int iStar;                       // >=1 stellar, 0 extra diffuse source
    int difGrid, difX, difY, difZ;   // diffuse source cell (iStar==0)
    long long firstId, n;            // global id of this rank's first packet, packets of this rank
    unsigned long long seed;         // Philox key of this call: context seed advanced by the epoch (capi.cu: call_seed)
    unsigned int rngStream;          // Philox counter word 3: iStar for a star; 0x80000000 + linear index of the
                                     // emitting cell for the extra diffuse source (every cell its own streams)
    unsigned long long pidBase;      // added to the packet index: difGrid << 48 for the extra diffuse source
    unsigned long long *nextPacket;  // work counter
    const unsigned int *order;       // optional: packet indices sorted by first frequency bin
    int aggSteps;                    // first steps of a flight with warp-aggregated tallies (0 = off)
    int batch;                       // lanes that must wait for a rare phase before it runs
    unsigned long long *counters;    // [C_COUNT]
    unsigned long long *qphotCounts; // [nbins]
    int *errFlag;
    int *fates;                      // optional [4*n]
    const ResCell *resCells;         // resonance-line transfer: this rank's source cells, or NULL
    const unsigned int *resPrefix;   // [nResCells+1] packets before each source cell (this rank)
    int nResCells;
    unsigned int *segsArr;           // wave-front + trace: segments of earlier flights per packet
};

// ---- wave-front pipeline (wavefront.cu) --------------------------------------------------
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
    int directA;                         // wave 0 with pre-ordered packets: EMIT writes recA directly, no sort
    int flyBatch;                        // idle lanes of a warp that trigger the store/claim pass of the FLY kernel
    int escCompact;                      // ESCAPE events carry the tally element instead of the record position
    int stepBudget;                      // cell crossings per flight per wave (longer flights continue
                                         // in the next wave, so one straggler cannot hold a wave open)
    unsigned int *hist, *cursor;         // [nbins+1]
    unsigned long long *nextFlight;      // FLY work counter
};

}  // namespace mcb
