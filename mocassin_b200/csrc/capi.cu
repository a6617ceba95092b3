// capi.cu -- the C ABI of libmocassin_b200.so (include/mcb200.h): context, device
// memory, uploads in the reference's layouts, kernel launches, fold, downloads.
#include "../../include/mcb200.h"
#include "types.h"
#include "detmath.cuh"
#include "philox.cuh"
#include "dust.h"
#include "transport_core.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <chrono>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace mcb {
cudaError_t launch_transport(const TransportArgs &a, bool multi, int gridBlocks, cudaStream_t stream, const WfArgs *resume);
int transport_blocks_per_sm(bool multi);
cudaError_t launch_order(const TransportArgs &a, unsigned short *key, unsigned int *hist, unsigned int *cursor,
                         unsigned int *order, int numSMs, cudaStream_t stream);
cudaError_t launch_transpose_pdf(const float *src, float *dst, int nRows, int nb, cudaStream_t s);
cudaError_t launch_check_monotone(const float *pdfT, int nRows, int nb, int *bad, cudaStream_t s);
cudaError_t launch_fold_j(unsigned long long *Q, float *J, const float *dV, int nRows, size_t first, size_t total,
                          double lenUnit, float deltaE, int blocks, cudaStream_t s);
struct P2PPeers {
    unsigned long long *Q[16];
    float *J[16];
    int nranks, rank;
};
cudaError_t launch_p2p_reduce_fold(const P2PPeers &P, const float *dV, int nRows, size_t first, size_t total,
                                   double lenUnit, float deltaE, int blocks, cudaStream_t s);
struct P2PPush {
    unsigned long long *recv[16];
    int nranks, rank;
};
cudaError_t launch_p2p_push(const P2PPush &P, const unsigned long long *Q, size_t off, size_t count, size_t slotStride,
                            size_t rOff, int blocks, cudaStream_t s);
struct P2PPackedPeers {
    unsigned int *lo[16];
    unsigned int *hi[16];
    unsigned char *flag[16];             // see tables.cu
    int nranks, rank;
};
int p2p_pack_block();
cudaError_t launch_p2p_push_packed(const P2PPackedPeers &P, const unsigned long long *Q, size_t off, size_t count, size_t rOff,
                                   size_t slotStride, size_t flagStride, int blocks, cudaStream_t s);
cudaError_t launch_p2p_sum_fold_packed(const P2PPeers &P, const unsigned int *lo, const unsigned int *hi, const unsigned char *flag,
                                       size_t slotStride, size_t flagStride, size_t rOff, const float *dV, int nRows, size_t first,
                                       size_t total, double lenUnit, float deltaE, int blocks, cudaStream_t s);
cudaError_t launch_p2p_sum_fold(const P2PPeers &P, const unsigned long long *recv, size_t slotStride, size_t rOff, const float *dV,
                                int nRows, size_t first, size_t total, double lenUnit, float deltaE, int blocks, cudaStream_t s);
cudaError_t launch_fold_count(unsigned int *Q, float *E, size_t total, float deltaE, int blocks,
                              cudaStream_t s);
cudaError_t launch_merge_sets(unsigned long long *J0, unsigned long long *J1, size_t nJ, unsigned int *E0,
                              unsigned int *E1, size_t nE, int *f0, int *f1, int nf, int blocks, cudaStream_t s);
cudaError_t wf_launch_event(const WfArgs &w, bool multi, int ev, int blocks, cudaStream_t s);
cudaError_t wf_launch_fly(const WfArgs &w, bool multi, int blocks, cudaStream_t s);
cudaError_t wf_launch_escape_compact(const WfArgs &w, int blocks, cudaStream_t s);
bool wf_esc_compact_built();
cudaError_t wf_launch_sort(const WfArgs &w, bool multi, int numSMs, cudaStream_t s);
int wf_fly_blocks_per_sm(bool multi);
size_t wf_rec_bytes();
size_t wf_recx_bytes();
struct OpacityArgs {
    int nRows, nb;
    int nSpeciesDen;
    const float *den;
    const float *ff1;
    const float *xSec;
    const int *nuStart;
    const int *nuBandSpecies;
    const int *nuBandXs;
    int nDustTerms;
    const float *Ndust;
    const unsigned char *dustOn;
    const float *dustCoef;
    const int *dustCompOfCell;
    const int *dustTermOn;
    const int *dustScaP, *dustAbsP;
    float *opacity, *scaOpac, *absOpac;
};
cudaError_t launch_opacity(const OpacityArgs &A, cudaStream_t s);
cudaError_t launch_esc_compact(unsigned int *Q, size_t off, size_t len, unsigned long long *list,
                               unsigned long long *count, unsigned long long capacity, int clear, int blocks,
                               cudaStream_t s);
cudaError_t launch_esc_clear_list(unsigned int *Q, const unsigned long long *list, unsigned long long n, int blocks, cudaStream_t s);
cudaError_t launch_esc_scatter(unsigned int *Q, size_t total, const unsigned long long *list, unsigned long long n,
                               int blocks, cudaStream_t s);
cudaError_t launch_sed_sum(const unsigned int *escQ, size_t nR, int firstPlane, int nPlanes,
                           unsigned long long *sedQ, cudaStream_t s);
cudaError_t launch_contcube(const float *esc, size_t nR, int nb, int nAngles, float *out, cudaStream_t s);
cudaError_t launch_gather_cells(const float *table, size_t nR, int nb, int first, int stride, int nMine, float *out, cudaStream_t s);
cudaError_t launch_gather_rows(const float *table, size_t nR, int nb, const int *cells, int nWanted, float *out, cudaStream_t s);
cudaError_t launch_dust_mask(const float *Tdust, const int *compOfCell, const int *dustComPoint, const int *nSpeciesPart,
                             const float *Tsub, int nRows, int nSpeciesTot, int nSizes, int s0, int s1, unsigned char *on,
                             cudaStream_t s);
}  // namespace mcb

using namespace mcb;

namespace {

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count)
    {
        if (p && n == count) return cudaSuccess;
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void **)&p, count * sizeof(T));
    }
    cudaError_t upload(const T *h, size_t count, cudaStream_t s)
    {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    cudaError_t zero(cudaStream_t s) { return n ? cudaMemsetAsync(p, 0, n * sizeof(T), s) : cudaSuccess; }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

struct GridState {
    bool set = false;
    int nx = 0, ny = 0, nz = 0, nCells = 0, motherP = 0, dense = 0;
    float geo[3] = {0, 0, 0};
    int lenExp = 0;                       // path-length unit = 2^lenExp cm
    std::vector<float> hx, hy, hz;        // host copies of the axes
    std::vector<int> hactive;
    std::vector<int> resLinePackets;      // (0:nCells) extra packets of the resonance-line transfer
    DevBuf<float> xAxis, yAxis, zAxis, xWall, yWall, zWall;
    DevBuf<int> active;
    DevBuf<float> opacity, scaOpac, absOpac, pdfT, totalLines, linePDF, dV, stage;
    DevBuf<float> contI;                  // mcb200_fetch_contcube: (0:nCells, 0:nAngleBins)
    DevBuf<float> cellStage;              // mcb200_fetch_estimators_cells: compact rows of one rank's cells
    DevBuf<unsigned char> canScatter;
    DevBuf<unsigned long long> JsteQ, JdifQ;
    DevBuf<unsigned int> escQ, lineQ;
    DevBuf<int> nuTouched;
    // second tally set (option "tally_set"=1): lets the caller all-reduce one half of a call's
    // tallies while the other half is still being transported (overlap of the exchange)
    DevBuf<unsigned long long> JsteQ2;
    DevBuf<unsigned int> escQ2;
    DevBuf<int> nuTouched2;
    DevBuf<float> Jste, Jdif, esc, linePk;
    // mcb200_fetch_escaped_sparse: the entries written by the previous call and where
    std::vector<unsigned int> sparsePrev;
    const float *sparsePrevPtr = nullptr;
    bool sparsePrevDense = false;         // the previous call fell back to the dense copy
    std::vector<char> folded;             // nu-planes of the pending call already folded by mcb200_reduce_range
    // mcb200_exchange reduce-scatters the touched J planes: after it this rank holds the global sum
    // only of its own share [off + rank*count, +count) of every range (and of the all-reduced tail
    // [off + nranks*count, off + len)); the fold then runs on that share and the float32 results are
    // all-gathered in place (capi.cu: fold_pending)
    struct JRange { size_t off, len, count; };
    std::vector<JRange> jShards;
    bool jShardsP2P = false;              // the shares are still to be summed: by the fused peer-memory kernel in the fold
    // peer mappings of every rank's JsteQ / Jste (cudaIpc), for the fused reduce + fold over NVLink
    std::vector<void *> peerQ, peerJ;     // [rank]; own entry = local pointer
    std::vector<void *> peerRecv;         // [rank]: every rank's receive buffer of the push variant (own = recvQ.p)
    DevBuf<unsigned long long> recvQ;     // (nranks-1) slots of slotStride partial sums pushed by the peers
    DevBuf<unsigned char> flagStage;      // packed push: flag bytes per destination rank, staged locally
    size_t slotStride = 0, flagStride = 0;
    bool jPacked = false;                 // the pending push went out packed (low words + flagged high words)
    void *peerBaseQ = nullptr, *peerBaseJ = nullptr;   // local pointers the mappings were made for
    // dust closure (mcb200_dust_update / mcb200_dust_pdf): device copy of the dust state
    DevBuf<float> Tdust;
    DevBuf<int> dustAbun, lgConverged;
    bool haveOpacity = false, havePdf = false;
};

}  // namespace

struct mcb200_ctx {
    int device = 0, rank = 0, nranks = 1;
    uint64_t seed = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;    // async PDF upload (option async_pdfs): overlaps wave 0
    cudaEvent_t pdfReady = nullptr;
    bool asyncPdfs = false, pdfPending = false;
    int tallySet = 0;                     // tally set new transport calls write to
    bool pending2 = false;                // set 1 holds tallies not yet merged into set 0
    int partIndex = 0, partCount = 1;     // sub-range of this rank's share (option part / parts)
    DevBuf<int> pdfBad;                   // deferred monotonicity verdicts, one per grid
    int numSMs = 0;
    bool haveCfg = false;
    mcb200_config cfg{};
    std::vector<GridState> grids;
    DevBuf<DevGrid> dGrids;
    std::vector<DevGrid> hGrids;
    bool gridsDirty = true;
    DevBuf<float> nuArray, gSca, starCdf, starPos, vpTheta, vpPhi, xSec;
    DevBuf<int> starIdx, starCell, vpPtheta, vpPphi;
    std::vector<float> hNu, hStarPos, hXsec;
    std::vector<int> hStarIdx;
    bool haveSpectra = false, haveStars = false, haveView = false;
    // dust species tables (host)
    std::vector<int> nSpeciesPart, dustComPoint;
    std::vector<float> grainAbun, TdustSublime;
    bool haveDustSpecies = false;
    // dust closure tables (mcb200_set_dust_tables)
    DevBuf<float> widFlx, grainWeight, emT, dGrainAbun, dSublime;
    DevBuf<int> absP, dSpeciesPart, dComPoint;
    DevBuf<unsigned long long> nConv;
    DevBuf<unsigned long long> sparseList, sparseCount;   // mcb200_fetch_escaped_sparse
    unsigned long long *sparseHost = nullptr;             // pinned staging of the (index, value) list
    size_t sparseHostCap = 0;
    // temporaries of mcb200_assemble_opacity
    DevBuf<float> opDen, opFf, opNd, opCoef;
    DevBuf<int> opStart, opSpec, opXs, opComp, opTermOn, opScaP, opAbsP;
    DevBuf<unsigned char> opOn;
    int nSpeciesTot = 0, nTemps = 0;
    bool haveDustTables = false;
    // work buffers
    DevBuf<unsigned long long> nextPacket, counters, qphot;
    DevBuf<int> errFlag, fates, flag;
    DevBuf<unsigned short> sortKey;
    DevBuf<unsigned int> sortHist, sortCursor, sortOrder;
    int orderMode = -1;                   // -1 auto, 0 off, 1 on: process packets in frequency order
    int waveMode = -1;                    // -1 auto, 0 persistent kernel, 1 wave-front pipeline
    DevBuf<unsigned char> wfRecA, wfRecB, wfRecXA, wfRecXB;
    DevBuf<unsigned int> wfEv0, wfEv1, wfEv2, wfEv3, wfCounts, wfHist, wfCursor, wfSegs;
    int stepBudget = 96;
    int wave0Order = 1;                   // wave 0 of the wave-front pipeline emits in first-frequency order: 0 off, 1 for >= 2^17 packets, 2 always
    unsigned int wave0Count = 0;
    bool wave0Exact = false;              // exact slots instead of appended records (measured slower)
    int wave0Blocks = 4;                  // CTAs per SM of the pre-ordered wave-0 emission
    bool escCompact = true;               // option esc_compact (only in builds with MCB_ESC_COMPACT)
    int flyBatch = 8;                     // FLY kernel: lanes of a warp that must be idle before records are stored / claimed
    int64_t tailThreshold = 32768;        // alive packets below which the persistent kernel finishes the batch
    DevBuf<unsigned char> wfArgsDev;
    DevBuf<int> planeDist;                // planeIonDistribution(grid(1)%nx, grid(1)%nz)
    DevBuf<ResCell> resCells;
    DevBuf<unsigned int> resPrefix;
    DevBuf<unsigned short> wfFlyKey;
    DevBuf<unsigned long long> wfNext;
    std::vector<cudaEvent_t> flyEv;       // start/stop of the FLY kernel of every wave (mcb200_counters.fly_ms)
    int lastWaves = 0, lastLaunches = 0, lastFoldLaunches = 0, rangeFoldLaunches = 0;
    int aggSteps = 0, batch = 12;
    bool trace = false;
    int blocksPerSM = 0;                  // 0 = occupancy default
    // SED(nu, angle) = sum over cells and grids of escapedPackets (writeSED): integer counts of
    // the pending call on the device, raw float sums and cumulative counts on the host
    DevBuf<unsigned long long> escList, escCount;   // sparse escape-count exchange (mcb200_escaped_compact)
    DevBuf<unsigned long long> sedQ;
    std::vector<float> sed;
    std::vector<long long> sedCount;
    bool sedLocal = false;                // option sed_local: sedQ is taken from this rank's escQ at
                                          // the end of the transport call and exchanged instead of escQ
    bool sedReady = false;                // sedQ already holds the pending call's counts
    // library-owned NCCL communicator (mcb200_comm_init) for hosts without a device-aware collective
    void *comm = nullptr;
    DevBuf<unsigned long long> commSizes, commPad, commGather;
    int64_t lastExchangeBytes = 0;        // bytes this rank handed to NCCL in the last mcb200_exchange
    int lastExchangeSparse = 0;           // grids whose escape counts went as sparse lists
    bool exchangeDense = false;           // option exchange_dense: mcb200_exchange never takes the sparse path
    bool exchangeAllReduce = false;       // option exchange_allreduce: all-reduce the J planes (round-1 path) instead of
                                          // reduce-scatter -> fold the share -> all-gather float32
    int p2pMode = -1;                     // option exchange_p2p: -1 auto (peer memory when it can be mapped), 0 NCCL only, 1 required
    int p2pPushBlocks = 4;                // option exchange_push_blocks: CTAs per SM of the push kernels (4 x 256 threads
                                          // leave half of every SM to the NCCL kernels of the escape-count exchange)
    bool p2pPack = true;                  // option exchange_pack: packed push (4 bytes per element + flagged high words)
    int p2pPush = 1;                      // option exchange_push: peers push their partial sums (copies), the owner sums locally;
                                          // 0 = the owner pulls them with peer loads inside the merge kernel
    int p2pState = 0;                     // 0 not tried, 1 usable, -1 unavailable (lastP2PWhy)
    std::string lastP2PWhy;
    int lastExchangePath = 0;             // 0 none, 1 all-reduce, 2 reduce-scatter/all-gather (NCCL), 3 fused peer-memory kernel
    double lastPhaseMs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // MCB200_TRACE_EXCHANGE=1: host-clock marks of the exchange + merge of the last call, printed by rank 0
    int traceOn = -1;
    std::chrono::steady_clock::time_point traceT0;
    std::vector<std::pair<const char *, double>> xtrace;
    DevBuf<unsigned char> ipcBuf;
    DevBuf<int> barrierWord;
    bool pdfSlabs = false;                // option pdf_slabs: 1/nranks of the PDF table per rank over PCIe, all-gather over NVLink
    int64_t lastPdfH2D = 0;               // bytes the last mcb200_set_pdfs moved host -> device for the CDF table
    int64_t epoch = 0;                    // option epoch: advances the Philox key (set to the Lucy iteration number)
    cudaStream_t sideStream = nullptr;    // fold of the escape counts / SED beside the link-bound J merge (multi-rank)
    cudaEvent_t sideEv0 = nullptr, sideEv1 = nullptr, evMid = nullptr, evPush0 = nullptr, pushDone = nullptr;
    bool evMidSet = false;
    bool solo = false;                    // option solo
    int soloRank = 0, soloNranks = 1;
    bool keepSharded = false;             // option keep_sharded: skip the all-gather (Jste stays valid only on the owner's share)
    bool exchanged = false;               // the pending tallies are already global (mcb200_exchange ran)
    bool deferFold = false;               // option defer_fold: a single rank keeps its tallies pending like a multi-rank run
    // pending fold
    bool pending = false;
    float pendingDeltaE = 0.f;
    std::string err;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
};

namespace {

int fail(mcb200_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? MCB200_ENOMEM : MCB200_ENODEV, \
                        "%s: %s", #call, cudaGetErrorString(e__));                            \
    } while (0)

#define NEED_CTX()                                   \
    do {                                             \
        if (!ctx) return MCB200_EINVAL;              \
        cudaSetDevice(ctx->device);                  \
    } while (0)

GridState *grid_of(mcb200_ctx *ctx, int iG)
{
    if (!ctx->haveCfg || iG < 1 || iG > ctx->cfg.nGrids) return nullptr;
    return &ctx->grids[iG - 1];
}

size_t tsize(const mcb200_ctx *ctx, const GridState &g) { return (size_t)(g.nCells + 1) * (size_t)ctx->cfg.nbins; }
size_t esize(const mcb200_ctx *ctx, const GridState &g)
{
    return (size_t)(g.nCells + 1) * (size_t)(ctx->cfg.nbins + 1) * (size_t)(ctx->cfg.nAngleBins + 1);
}
size_t lsize(const mcb200_ctx *ctx, const GridState &g) { return (size_t)(g.nCells + 1) * (size_t)ctx->cfg.nLines; }

// mid-point walls, same float32 expression as photon_mod.f90:1266 / :1281
std::vector<float> make_walls(const std::vector<float> &a)
{
    size_t n = a.size();
    std::vector<float> w(n + 1);
    w[0] = a[0];
    for (size_t i = 1; i < n; ++i) {
        volatile float s = a[i] + a[i - 1];
        w[i] = s / 2.f;
    }
    w[n] = a[n - 1];
    return w;
}

// cell width / 1e15 per axis, photon_mod.f90:1469-1508
std::vector<float> make_widths(const std::vector<float> &a, bool sym)
{
    size_t n = a.size();
    std::vector<float> w(n, 0.f);
    for (size_t i = 0; i < n; ++i) {
        volatile float d;
        if (i > 0 && i + 1 < n) { d = std::fabs(a[i + 1] - a[i - 1]); d = d / 2.f; }
        else if (i == 0) { d = std::fabs(a[1] - a[0]); if (sym) d = d / 2.f; }
        else d = std::fabs(a[n - 1] - a[n - 2]);
        w[i] = d / 1.e15f;
    }
    return w;
}

int sync_grids(mcb200_ctx *ctx)
{
    if (!ctx->gridsDirty) return MCB200_OK;
    int nG = ctx->cfg.nGrids;
    ctx->hGrids.assign(nG, DevGrid{});
    for (int i = 0; i < nG; ++i) {
        GridState &g = ctx->grids[i];
        DevGrid &d = ctx->hGrids[i];
        d.nx = g.nx; d.ny = g.ny; d.nz = g.nz; d.nCells = g.nCells; d.motherP = g.motherP;
        d.dense = g.dense;
        d.geoX = g.geo[0]; d.geoY = g.geo[1]; d.geoZ = g.geo[2];
        d.x1 = g.hx.front(); d.y1 = g.hy.front(); d.z1 = g.hz.front();
        d.xN = g.hx.back();  d.yN = g.hy.back();  d.zN = g.hz.back();
        {
            volatile float t;
            t = d.x1 - d.geoX; d.xLo = t; t = d.y1 - d.geoY; d.yLo = t; t = d.z1 - d.geoZ; d.zLo = t;
            t = d.xN + d.geoX; d.xHi = t; t = d.yN + d.geoY; d.yHi = t; t = d.zN + d.geoZ; d.zHi = t;
        }
        d.invLenUnit = std::ldexp(1.0f, -g.lenExp);
        d.xAxis = g.xAxis.p; d.yAxis = g.yAxis.p; d.zAxis = g.zAxis.p;
        d.xWall = g.xWall.p; d.yWall = g.yWall.p; d.zWall = g.zWall.p;
        d.active = g.active.p;
        d.opacity = g.opacity.p; d.scaOpac = g.scaOpac.p;
        d.pdfT = g.pdfT.p; d.totalLines = g.totalLines.p; d.linePDF = g.linePDF.p;
        d.canScatter = g.canScatter.p;
        d.JsteQ = g.JsteQ.p; d.JdifQ = g.JdifQ.p; d.escQ = g.escQ.p; d.lineQ = g.lineQ.p;
        d.nuTouched = g.nuTouched.p;
        if (ctx->tallySet == 1) { d.JsteQ = g.JsteQ2.p; d.escQ = g.escQ2.p; d.nuTouched = g.nuTouched2.p; }
    }
    CU(ctx->dGrids.upload(ctx->hGrids.data(), nG, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->gridsDirty = false;
    return MCB200_OK;
}

int ensure_estimators(mcb200_ctx *ctx, GridState &g)
{
    size_t ts = tsize(ctx, g), es = esize(ctx, g);
    bool fresh = (g.JsteQ.n != ts);
    if (!fresh) return MCB200_OK;
    CU(g.JsteQ.alloc(ts)); CU(g.JsteQ.zero(ctx->stream));
    CU(g.Jste.alloc(ts));  CU(g.Jste.zero(ctx->stream));
    CU(g.escQ.alloc(es));  CU(g.escQ.zero(ctx->stream));
    CU(g.nuTouched.alloc(ctx->cfg.nbins + 1)); CU(g.nuTouched.zero(ctx->stream));
    CU(g.esc.alloc(es));   CU(g.esc.zero(ctx->stream));
    if (ctx->cfg.lgDebug) {
        CU(g.JdifQ.alloc(ts)); CU(g.JdifQ.zero(ctx->stream));
        CU(g.Jdif.alloc(ts));  CU(g.Jdif.zero(ctx->stream));
        size_t ls = lsize(ctx, g);
        CU(g.lineQ.alloc(ls)); CU(g.lineQ.zero(ctx->stream));
        CU(g.linePk.alloc(ls)); CU(g.linePk.zero(ctx->stream));
    }
    ctx->gridsDirty = true;
    return MCB200_OK;
}

int ensure_second_set(mcb200_ctx *ctx, GridState &g)
{
    size_t ts = tsize(ctx, g), es = esize(ctx, g);
    if (g.JsteQ2.n == ts) return MCB200_OK;
    CU(g.JsteQ2.alloc(ts)); CU(g.JsteQ2.zero(ctx->stream));
    CU(g.escQ2.alloc(es)); CU(g.escQ2.zero(ctx->stream));
    CU(g.nuTouched2.alloc(ctx->cfg.nbins + 1)); CU(g.nuTouched2.zero(ctx->stream));
    ctx->gridsDirty = true;
    return MCB200_OK;
}

// ---- NCCL, bound at run time -----------------------------------------------------------------
// The library has no link-time dependency on NCCL: single-rank hosts never need it, and a
// Python host has torch's copy in the process already (dlopen by SONAME returns that one).
// Search order: $MCB200_NCCL_LIB, libnccl.so.2, libnccl.so.  Types restated from nccl.h (2.x ABI).
struct NcclId { char internal[128]; };
enum { kNcclInt32 = 2, kNcclUint32 = 3, kNcclUint64 = 5, kNcclFloat32 = 7, kNcclSum = 0, kNcclMax = 2 };
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    std::string why;
};

NcclApi &nccl_api()
{
    static NcclApi api;
    if (api.handle || !api.why.empty()) return api;
    const char *env = std::getenv("MCB200_NCCL_LIB");
    const char *names[3] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (api.handle) break;
    }
    if (!api.handle) { api.why = "libnccl.so.2 not found (set MCB200_NCCL_LIB)"; return api; }
    bool ok = true;
    auto sym = [&](const char *n) { void *p = dlsym(api.handle, n); if (!p) { ok = false; api.why = std::string("NCCL symbol missing: ") + n; } return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(sym("ncclReduceScatter"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    return api;
}

#define NC(call)                                                                                   \
    do {                                                                                           \
        int r__ = (call);                                                                          \
        if (r__ != 0) return fail(ctx, MCB200_ECOMM, "%s: %s", #call, nccl_api().GetErrorString(r__)); \
    } while (0)

// contiguous runs [first,last] of touched frequency bins (gaps of <= 2 bins are bridged)
std::vector<std::pair<int, int>> touched_ranges(const std::vector<int> &flag, int bridge = 3)
{
    std::vector<std::pair<int, int>> r;
    int n = (int)flag.size();
    for (int i = 0; i < n; ++i) {
        if (!flag[i]) continue;
        if (!r.empty() && i - r.back().second <= bridge) r.back().second = i;
        else r.emplace_back(i, i);
    }
    return r;
}

size_t sed_size(const mcb200_ctx *ctx) { return (size_t)(ctx->cfg.nbins + 1) * (size_t)(ctx->cfg.nAngleBins + 1); }

int ensure_sed(mcb200_ctx *ctx)
{
    size_t n = sed_size(ctx);
    if (ctx->sedQ.n != n) {
        CU(ctx->sedQ.alloc(n)); CU(ctx->sedQ.zero(ctx->stream));
        ctx->sed.assign(n, 0.f); ctx->sedCount.assign(n, 0);
    }
    return MCB200_OK;
}

// sedQ += per-plane sums of the escape counts of every grid (tally set `set`)
int sed_tally(mcb200_ctx *ctx, int set, int *launches, cudaStream_t st = nullptr)
{
    int rc = ensure_sed(ctx);
    if (rc) return rc;
    if (!st) st = ctx->stream;
    const int nb = ctx->cfg.nbins;
    for (auto &g : ctx->grids) {
        const unsigned int *q = set == 1 ? g.escQ2.p : g.escQ.p;
        const int *touched = set == 1 ? g.nuTouched2.p : g.nuTouched.p;
        if (!q) continue;
        size_t nR = (size_t)g.nCells + 1;
        std::vector<int> flag(nb + 1, 1);
        if (touched) {
            CU(cudaMemcpyAsync(flag.data(), touched, sizeof(int) * (nb + 1), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        if (set == 0 && !g.folded.empty())           // planes a ranged fold has already tallied and cleared
            for (int nu = 0; nu <= nb; ++nu) if (g.folded[nu]) flag[nu] = 0;
        for (auto &rg : touched_ranges(flag, g.folded.empty() ? 3 : 1))
            for (int ang = 0; ang <= ctx->cfg.nAngleBins; ++ang)
            {
                CU(launch_sed_sum(q, nR, rg.first + (nb + 1) * ang, rg.second - rg.first + 1, ctx->sedQ.p, st));
                if (launches) ++*launches;
            }
    }
    return MCB200_OK;
}

// fold the tallies of the nu-planes [nu0, nu1] of one grid: J planes nu >= 1, escape-count planes
// nu0..nu1 of every viewing angle.  Asynchronous on the library stream.
int fold_planes(mcb200_ctx *ctx, GridState &g, int nu0, int nu1, int *launches, bool withJ = true, cudaStream_t st = nullptr)
{
    if (!st) st = ctx->stream;
    const int nb = ctx->cfg.nbins, blocks = ctx->numSMs * 8;
    const size_t nR = (size_t)g.nCells + 1;
    const double lenUnit = std::ldexp(1.0, g.lenExp);
    int p0 = nu0 < 1 ? 1 : nu0, p1 = nu1;
    if (withJ && p1 >= p0) {
        size_t off = (size_t)(p0 - 1) * nR, len = (size_t)(p1 - p0 + 1) * nR;
        CU(launch_fold_j(g.JsteQ.p + off, g.Jste.p + off, g.dV.p, (int)nR, off, len, lenUnit, ctx->pendingDeltaE, blocks, st));
        if (launches) ++*launches;
        if (ctx->cfg.lgDebug) {
            CU(launch_fold_j(g.JdifQ.p + off, g.Jdif.p + off, g.dV.p, (int)nR, off, len, lenUnit, ctx->pendingDeltaE, blocks, st));
            if (launches) ++*launches;
        }
    }
    for (int ang = 0; ang <= ctx->cfg.nAngleBins; ++ang) {
        size_t off = nR * ((size_t)nu0 + (size_t)(nb + 1) * (size_t)ang);
        size_t len = (size_t)(nu1 - nu0 + 1) * nR;
        CU(launch_fold_count(g.escQ.p + off, g.esc.p + off, len, ctx->pendingDeltaE, blocks, st));
        if (launches) ++*launches;
    }
    return MCB200_OK;
}

// stream-ordered barrier over the ranks: a one-word all-reduce completes on a rank only after every
// rank's stream has reached it
int comm_barrier(mcb200_ctx *ctx)
{
    CU(ctx->barrierWord.alloc(1));
    NC(nccl_api().AllReduce(ctx->barrierWord.p, ctx->barrierWord.p, 1, kNcclInt32, kNcclMax, ctx->comm, ctx->stream));
    return MCB200_OK;
}

void p2p_close(GridState &g, int rank)
{
    for (size_t r = 0; r < g.peerQ.size(); ++r) {
        if ((int)r == rank) continue;
        if (g.peerQ[r]) cudaIpcCloseMemHandle(g.peerQ[r]);
        if (g.peerJ[r]) cudaIpcCloseMemHandle(g.peerJ[r]);
        if (r < g.peerRecv.size() && g.peerRecv[r]) cudaIpcCloseMemHandle(g.peerRecv[r]);
    }
    g.peerQ.clear(); g.peerJ.clear(); g.peerRecv.clear();
    g.peerBaseQ = g.peerBaseJ = nullptr;
}

// Map every rank's JsteQ and Jste of this grid into this process (cudaIpc handles, all-gathered
// through the communicator).  Collective: every rank calls it at the same point.  Leaves
// ctx->p2pState = 1 when every rank mapped every peer, -1 otherwise (the exchange then stays on NCCL).
int p2p_setup(mcb200_ctx *ctx, GridState &g)
{
    NcclApi &N = nccl_api();
    const int world = ctx->nranks, rank = ctx->rank;
    cudaStream_t s = ctx->stream;
    // do the mappings of every rank still match the buffers?  (max over ranks of "mine changed")
    int stale = (g.peerQ.empty() || g.peerBaseQ != (void *)g.JsteQ.p || g.peerBaseJ != (void *)g.Jste.p) ? 1 : 0;
    CU(ctx->barrierWord.alloc(1));
    CU(cudaMemcpyAsync(ctx->barrierWord.p, &stale, sizeof(int), cudaMemcpyHostToDevice, s));
    NC(N.AllReduce(ctx->barrierWord.p, ctx->barrierWord.p, 1, kNcclInt32, kNcclMax, ctx->comm, s));
    CU(cudaMemcpyAsync(&stale, ctx->barrierWord.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (!stale) return MCB200_OK;
    p2p_close(g, rank);
    // receive buffer of the push variant: one slot per peer, each large enough for this rank's share
    // of every range of the J table
    // a share of every range, each starting on a flag-block boundary inside the slot (up to 64 ranges)
    g.slotStride = (g.JsteQ.n / (size_t)world + 1 + (size_t)65 * (size_t)p2p_pack_block()) / (size_t)p2p_pack_block() * (size_t)p2p_pack_block();
    g.flagStride = g.slotStride / (size_t)p2p_pack_block() + 2;
    {   // 8 bytes per element (64-bit partial sums, or low + high words of the packed push) + one flag byte per block
        const size_t words = g.slotStride * (size_t)(world - 1) + (g.flagStride * (size_t)(world - 1) + 7) / 8 + 1;
        if (g.recvQ.n != words) CU(g.recvQ.alloc(words));
    }
    struct Pair { cudaIpcMemHandle_t q, j, rcv; int ok; int pad[3]; };
    static_assert(sizeof(Pair) % 8 == 0, "handle record must be a multiple of 8 bytes");
    std::vector<Pair> all((size_t)world);
    Pair mine{};
    mine.ok = 1;
    if (world > 16) { mine.ok = 0; ctx->lastP2PWhy = "more than 16 ranks"; }
    if (mine.ok && cudaIpcGetMemHandle(&mine.q, g.JsteQ.p) != cudaSuccess) { mine.ok = 0; ctx->lastP2PWhy = "cudaIpcGetMemHandle(JsteQ) failed"; cudaGetLastError(); }
    if (mine.ok && cudaIpcGetMemHandle(&mine.j, g.Jste.p) != cudaSuccess) { mine.ok = 0; ctx->lastP2PWhy = "cudaIpcGetMemHandle(Jste) failed"; cudaGetLastError(); }
    if (mine.ok && cudaIpcGetMemHandle(&mine.rcv, g.recvQ.p) != cudaSuccess) { mine.ok = 0; ctx->lastP2PWhy = "cudaIpcGetMemHandle(recvQ) failed"; cudaGetLastError(); }
    CU(ctx->ipcBuf.alloc(sizeof(Pair) * (size_t)(world + 1)));
    CU(cudaMemcpyAsync(ctx->ipcBuf.p + sizeof(Pair) * (size_t)world, &mine, sizeof(Pair), cudaMemcpyHostToDevice, s));
    NC(N.AllGather(ctx->ipcBuf.p + sizeof(Pair) * (size_t)world, ctx->ipcBuf.p, sizeof(Pair) / 8, kNcclUint64, ctx->comm, s));
    CU(cudaMemcpyAsync(all.data(), ctx->ipcBuf.p, sizeof(Pair) * (size_t)world, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    int ok = 1;
    for (auto &h : all) ok = ok && h.ok;
    g.peerQ.assign((size_t)world, nullptr); g.peerJ.assign((size_t)world, nullptr); g.peerRecv.assign((size_t)world, nullptr);
    for (int r = 0; r < world && ok; ++r) {
        if (r == rank) { g.peerQ[r] = g.JsteQ.p; g.peerJ[r] = g.Jste.p; g.peerRecv[r] = g.recvQ.p; continue; }
        if (cudaIpcOpenMemHandle(&g.peerQ[r], all[r].q, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&g.peerJ[r], all[r].j, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&g.peerRecv[r], all[r].rcv, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            ok = 0;
            ctx->lastP2PWhy = std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError());
        }
    }
    // every rank must take the same path
    CU(cudaMemcpyAsync(ctx->barrierWord.p, &ok, sizeof(int), cudaMemcpyHostToDevice, s));
    NC(N.AllReduce(ctx->barrierWord.p, ctx->barrierWord.p, 1, kNcclInt32, 3 /* ncclMin */, ctx->comm, s));
    CU(cudaMemcpyAsync(&ok, ctx->barrierWord.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (!ok) {
        p2p_close(g, rank);
        if (ctx->lastP2PWhy.empty()) ctx->lastP2PWhy = "a peer could not map the buffers";
        ctx->p2pState = -1;
        return MCB200_OK;
    }
    g.peerBaseQ = g.JsteQ.p; g.peerBaseJ = g.Jste.p;
    ctx->p2pState = 1;
    return MCB200_OK;
}

// position of the next range's share inside a receive slot: on a flag-block boundary
size_t slot_advance(size_t rOff, size_t count)
{
    const size_t blk = (size_t)p2p_pack_block();
    return (rOff + count + 3 + blk - 1) / blk * blk;   // (+3: a share starts at its element index mod 4 inside the slot)
}

void trace_mark(mcb200_ctx *ctx, const char *what, bool reset = false)
{
    if (ctx->traceOn < 0) { const char *e = getenv("MCB200_TRACE_EXCHANGE"); ctx->traceOn = e && *e == '1' ? 1 : 0; }
    if (!ctx->traceOn) return;
    if (reset) { ctx->xtrace.clear(); ctx->traceT0 = std::chrono::steady_clock::now(); }
    ctx->xtrace.push_back({what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ctx->traceT0).count()});
}

int ensure_side_stream(mcb200_ctx *ctx)
{
    if (!ctx->sideStream) {
        CU(cudaStreamCreateWithFlags(&ctx->sideStream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->sideEv0, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->sideEv1, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->pushDone, cudaEventDisableTiming));
        CU(cudaEventCreate(&ctx->evPush0));
        CU(cudaEventCreate(&ctx->evMid));
    }
    return MCB200_OK;
}

// Push variant of the peer-memory merge, first phase, issued by mcb200_exchange as soon as the shares
// are known, on the side stream: it runs on the copy engines (or a streaming kernel) and the links,
// beside the escape-count exchange that mcb200_exchange does on the library stream.  Every rank
// copies, for every peer, the peer's share of its partial sums into the peer's receive buffer
// (device-to-device through the mapped buffers: posted writes), then clears the shares it has
// handed over.  The receive buffers are free: every rank passed the closing barrier of the previous merge.
int p2p_push_phase(mcb200_ctx *ctx, GridState &g, bool first)
{
    const int world = ctx->nranks, rank = ctx->rank;
    int rc = ensure_side_stream(ctx);
    if (rc) return rc;
    cudaStream_t ss = ctx->sideStream;
    CU(cudaEventRecord(ctx->sideEv0, ctx->stream));
    CU(cudaStreamWaitEvent(ss, ctx->sideEv0, 0));
    if (first) CU(cudaEventRecord(ctx->evPush0, ss));
    g.jPacked = false;
    if (ctx->p2pPack && g.jShards.size() <= 64) {
        if (g.flagStage.n != (size_t)world * g.flagStride) CU(g.flagStage.alloc((size_t)world * g.flagStride));
        P2PPackedPeers PP{};
        PP.nranks = world; PP.rank = rank;
        for (int r = 0; r < world; ++r) {
            PP.lo[r] = (unsigned int *)g.peerRecv[r];
            PP.hi[r] = PP.lo[r] + (size_t)(world - 1) * g.slotStride;
            PP.flag[r] = g.flagStage.p + (size_t)r * g.flagStride;
        }
        size_t rOff = 0;
        for (auto &r : g.jShards) {
            CU(launch_p2p_push_packed(PP, g.JsteQ.p, r.off, r.count, rOff, g.slotStride, g.flagStride, ctx->numSMs * ctx->p2pPushBlocks, ss));
            rOff = slot_advance(rOff, r.count);
            ctx->lastExchangeBytes += (int64_t)r.count * (int64_t)(world - 1) * 4;   // + the flagged high words (not counted)
        }
        for (int d = 1; d < world; ++d) {            // the flag bytes: one copy per peer
            const int r = (rank + d) % world, slot = rank < r ? rank : rank - 1;
            unsigned char *dst = (unsigned char *)((unsigned long long *)g.peerRecv[r] + (size_t)(world - 1) * g.slotStride) + (size_t)slot * g.flagStride;
            CU(cudaMemcpyAsync(dst, g.flagStage.p + (size_t)r * g.flagStride, rOff / (size_t)p2p_pack_block(), cudaMemcpyDeviceToDevice, ss));
        }
        CU(cudaEventRecord(ctx->evMid, ss));
        ctx->evMidSet = true;
        CU(cudaEventRecord(ctx->pushDone, ss));      // the merge may start here ...
        // ... while what was handed over is cleared behind it on the side stream (disjoint from the share the
        // merge kernel reads and clears; the library stream joins the side stream at the end of the fold).
        // A separate pass, because a store to a line whose load is still in flight inside the push kernel
        // takes a slow path: measured 2.4x on the whole kernel
        for (auto &r : g.jShards) {
            if (rank > 0) CU(cudaMemsetAsync(g.JsteQ.p + r.off, 0, (size_t)rank * r.count * 8, ss));
            if (rank + 1 < world) CU(cudaMemsetAsync(g.JsteQ.p + r.off + (size_t)(rank + 1) * r.count, 0, (size_t)(world - 1 - rank) * r.count * 8, ss));
        }
        g.jPacked = true;
        return MCB200_OK;
    }
    {
        size_t rOff = 0;
        if (ctx->p2pPush == 2) {
            P2PPush PP{};
            PP.nranks = world; PP.rank = rank;
            for (int r = 0; r < world; ++r) PP.recv[r] = (unsigned long long *)g.peerRecv[r];
            for (auto &r : g.jShards) {
                CU(launch_p2p_push(PP, g.JsteQ.p, r.off, r.count, g.slotStride, rOff, ctx->numSMs * ctx->p2pPushBlocks, ss));
                rOff = slot_advance(rOff, r.count);
                ctx->lastExchangeBytes += (int64_t)r.count * (int64_t)(world - 1) * 12;
            }
        } else {
            for (auto &r : g.jShards) {
                for (int d = 1; d < world; ++d) {
                    const int peer = (rank + d) % world;                       // staggered: no two ranks start on the same target
                    const int slot = rank < peer ? rank : rank - 1;           // this rank's slot in the peer's buffer
                    unsigned long long *dst = (unsigned long long *)g.peerRecv[peer] + (size_t)slot * g.slotStride + rOff;
                    CU(cudaMemcpyAsync(dst, g.JsteQ.p + r.off + (size_t)peer * r.count, r.count * 8, cudaMemcpyDeviceToDevice, ss));
                }
                rOff = slot_advance(rOff, r.count);
                ctx->lastExchangeBytes += (int64_t)r.count * (int64_t)(world - 1) * 12;
            }
        }
    }
    CU(cudaEventRecord(ctx->evMid, ss));
    ctx->evMidSet = true;
    {
        for (auto &r : g.jShards) {          // handed over: clear (the owner never reads these)
            const size_t mine = r.off + (size_t)rank * r.count, tail = r.off + (size_t)world * r.count;
            if (mine > r.off) CU(cudaMemsetAsync(g.JsteQ.p + r.off, 0, (mine - r.off) * 8, ss));
            if (tail > mine + r.count) CU(cudaMemsetAsync(g.JsteQ.p + mine + r.count, 0, (tail - mine - r.count) * 8, ss));
        }
    }
    CU(cudaEventRecord(ctx->pushDone, ss));
    return MCB200_OK;
}

// J planes after the reduce-scatter of mcb200_exchange (comm_exchange): this rank holds the global
// integer sums of its share of every exchanged range, the tail of each range on every rank.  Fold
// the share (1/nranks of the work), clear the other ranks' shares (partial sums that have been
// handed over), then all-gather the float32 results in place: every rank ends with the estimator
// the all-reduce + full fold gave, bit for bit (same integer sums, same fold arithmetic), with
// 12 instead of 16 bytes per element on the wire and the fold divided by nranks.
int fold_shards(mcb200_ctx *ctx, GridState &g, int *launches)
{
    NcclApi &N = nccl_api();
    const int blocks = ctx->numSMs * 8, world = ctx->nranks, rank = ctx->rank;
    const size_t nR = (size_t)g.nCells + 1;
    const double lenUnit = std::ldexp(1.0, g.lenExp);
    cudaStream_t s = ctx->stream;
    if (g.jShardsP2P) {
        // fused peer-memory path: every rank's transport kernels are behind the collectives of
        // mcb200_exchange on its stream, so the partial sums are complete everywhere
        P2PPeers P{};
        P.nranks = world; P.rank = rank;
        for (int r = 0; r < world; ++r) { P.Q[r] = (unsigned long long *)g.peerQ[r]; P.J[r] = (float *)g.peerJ[r]; }
        if (ctx->p2pPush) {
            // push variant: every rank copies, for every peer, the peer's share of its partial sums into
            // the peer's receive buffer (device-to-device copies through the mapped buffers: posted
            // writes at copy-engine speed), a barrier, then one local kernel per range sums, folds and
            // stores the float32 result into all ranks' Jste
            // the pushes were issued by mcb200_exchange on the side stream (p2p_push_phase)
            CU(cudaStreamWaitEvent(s, ctx->pushDone, 0));
            size_t rOff = 0;
            int rcb = comm_barrier(ctx);
            if (rcb) return rcb;
            rOff = 0;
            for (auto &r : g.jShards) {
                const size_t mine = r.off + (size_t)rank * r.count, tail = r.off + (size_t)world * r.count;
                if (g.jPacked) {
                    const unsigned int *lo = (const unsigned int *)g.recvQ.p, *hi = lo + (size_t)(world - 1) * g.slotStride;
                    const unsigned char *fl = (const unsigned char *)(g.recvQ.p + (size_t)(world - 1) * g.slotStride);
                    CU(launch_p2p_sum_fold_packed(P, lo, hi, fl, g.slotStride, g.flagStride, rOff, g.dV.p, (int)nR, mine, r.count, lenUnit,
                                                  ctx->pendingDeltaE, ctx->numSMs * 16, s));
                } else
                CU(launch_p2p_sum_fold(P, g.recvQ.p, g.slotStride, rOff, g.dV.p, (int)nR, mine, r.count, lenUnit, ctx->pendingDeltaE,
                                       ctx->numSMs * 16, s));
                CU(launch_fold_j(g.JsteQ.p + tail, g.Jste.p + tail, g.dV.p, (int)nR, tail, r.off + r.len - tail, lenUnit, ctx->pendingDeltaE, blocks, s));
                if (launches) *launches += 2;
                rOff = slot_advance(rOff, r.count);
            }
        } else
        for (auto &r : g.jShards) {
            const size_t mine = r.off + (size_t)rank * r.count, tail = r.off + (size_t)world * r.count;
            CU(launch_p2p_reduce_fold(P, g.dV.p, (int)nR, mine, r.count, lenUnit, ctx->pendingDeltaE, ctx->numSMs * 16, s));
            CU(launch_fold_j(g.JsteQ.p + tail, g.Jste.p + tail, g.dV.p, (int)nR, tail, r.off + r.len - tail, lenUnit, ctx->pendingDeltaE, blocks, s));
            if (launches) *launches += 2;
            ctx->lastExchangeBytes += (int64_t)r.count * (int64_t)(world - 1) * 12;
        }
        // nobody may clear (or overwrite, in the next transport call) partial sums a peer is still
        // reading, and nobody may read Jste before every owner has stored its share
        int rc = comm_barrier(ctx);
        if (rc) return rc;
        if (!ctx->p2pPush)               // (push variant: cleared right behind the pushes, on the side stream)
            for (auto &r : g.jShards) {
                const size_t mine = r.off + (size_t)rank * r.count, tail = r.off + (size_t)world * r.count;
                if (mine > r.off) CU(cudaMemsetAsync(g.JsteQ.p + r.off, 0, (mine - r.off) * 8, s));
                if (tail > mine + r.count) CU(cudaMemsetAsync(g.JsteQ.p + mine + r.count, 0, (tail - mine - r.count) * 8, s));
            }
        g.jShards.clear();
        g.jShardsP2P = false;
        return MCB200_OK;
    }
    const int nSets = ctx->cfg.lgDebug && g.JdifQ.p ? 2 : 1;
    for (int set = 0; set < nSets; ++set) {
        unsigned long long *Q = set ? g.JdifQ.p : g.JsteQ.p;
        float *J = set ? g.Jdif.p : g.Jste.p;
        for (auto &r : g.jShards) {
            const size_t mine = r.off + (size_t)rank * r.count, tail = r.off + (size_t)world * r.count;
            CU(launch_fold_j(Q + mine, J + mine, g.dV.p, (int)nR, mine, r.count, lenUnit, ctx->pendingDeltaE, blocks, s));
            CU(launch_fold_j(Q + tail, J + tail, g.dV.p, (int)nR, tail, r.off + r.len - tail, lenUnit, ctx->pendingDeltaE, blocks, s));
            if (launches) *launches += 2;
            if (mine > r.off) CU(cudaMemsetAsync(Q + r.off, 0, (mine - r.off) * 8, s));
            if (tail > mine + r.count) CU(cudaMemsetAsync(Q + mine + r.count, 0, (tail - mine - r.count) * 8, s));
        }
        if (!ctx->keepSharded) {
            NC(N.GroupStart());
            for (auto &r : g.jShards)
                if (r.count) {
                    NC(N.AllGather(J + r.off + (size_t)rank * r.count, J + r.off, r.count, kNcclFloat32, ctx->comm, s));
                    ctx->lastExchangeBytes += (int64_t)r.count * 4;
                }
            NC(N.GroupEnd());
        }
    }
    g.jShards.clear();
    return MCB200_OK;
}

int fold_pending(mcb200_ctx *ctx)
{
    if (!ctx->pending) return MCB200_OK;
    int blocks = ctx->numSMs * 8;
    const int nb = ctx->cfg.nbins;
    int launches = 0;
    cudaStream_t s = ctx->stream;
    if (ctx->pending2) {
        // integer merge first (exact), then ONE fold: fold(Q0)+fold(Q1) would round differently
        for (auto &g : ctx->grids) {
            if (!g.JsteQ2.p) continue;
            CU(launch_merge_sets(g.JsteQ.p, g.JsteQ2.p, g.JsteQ.n, g.escQ.p, g.escQ2.p, g.escQ.n,
                                 g.nuTouched.p, g.nuTouched2.p, nb + 1, blocks, s));
            ++launches;
        }
        ctx->pending2 = false;
    }
    // Multi-rank: the merge of the J planes (fold_shards) is bound by the links, the escape-count and
    // SED work by HBM and small.  The first goes on the library stream, the rest on a side stream, so
    // that they overlap instead of queueing up behind each other.
    bool anySharded = false;
    for (auto &g : ctx->grids) anySharded = anySharded || !g.jShards.empty();
    cudaStream_t es = s;
    if (anySharded) {
        { int rcs = ensure_side_stream(ctx); if (rcs) return rcs; }
        es = ctx->sideStream;
        CU(cudaEventRecord(ctx->sideEv0, s));
        CU(cudaStreamWaitEvent(es, ctx->sideEv0, 0));
        CU(cudaEventRecord(ctx->ev0, s));
        for (auto &g : ctx->grids)
            if (!g.jShards.empty()) { int rc = fold_shards(ctx, g, &launches); if (rc) return rc; }
        CU(cudaEventRecord(ctx->ev1, s));
    }
    {
        // SED: counts of this call per (nu, angle) over all cells and grids, folded like
        // escapedPackets: SED += float(count) * deltaE
        if (!ctx->sedReady) { int rc = sed_tally(ctx, 0, &launches, es); if (rc) return rc; }
        size_t n = sed_size(ctx);
        std::vector<unsigned long long> q(n);
        CU(cudaMemcpyAsync(q.data(), ctx->sedQ.p, n * 8, cudaMemcpyDeviceToHost, es));
        CU(cudaStreamSynchronize(es));
        CU(cudaMemsetAsync(ctx->sedQ.p, 0, ctx->sedQ.n * 8, es));
        for (size_t k = 0; k < n; ++k)
            if (q[k]) {
                volatile float add = (float)q[k] * ctx->pendingDeltaE;
                volatile float sum = ctx->sed[k] + add;
                ctx->sed[k] = sum;
                ctx->sedCount[k] += (long long)q[k];
            }
        ctx->sedReady = false;
    }
    for (auto &g : ctx->grids) {
        // only nu-planes in which a packet was emitted can hold tallies
        std::vector<int> flag(nb + 1, 1);
        if (g.nuTouched.p) {
            CU(cudaMemcpyAsync(flag.data(), g.nuTouched.p, sizeof(int) * (nb + 1), cudaMemcpyDeviceToHost, es));
            CU(cudaStreamSynchronize(es));
        }
        // planes a ranged fold (mcb200_reduce_range) has already taken are skipped
        if (!g.folded.empty())
            for (int nu = 0; nu <= nb; ++nu) if (g.folded[nu]) flag[nu] = 0;
        // J planes of a grid whose tallies were not scattered over the ranks are folded here, all of
        // them on every rank; the escape counts always are
        const bool withJ = !anySharded;
        for (auto &rg : touched_ranges(flag, g.folded.empty() ? 3 : 1)) {
            int rc = fold_planes(ctx, g, rg.first, rg.second, &launches, withJ, es);
            if (rc) return rc;
        }
        g.folded.clear();
        if (ctx->cfg.lgDebug && g.lineQ.n) {
            CU(launch_fold_count(g.lineQ.p, g.linePk.p, g.lineQ.n, ctx->pendingDeltaE, blocks, es));
            ++launches;
        }
        if (g.nuTouched.p) CU(cudaMemsetAsync(g.nuTouched.p, 0, g.nuTouched.n * sizeof(int), es));
    }
    if (anySharded) {
        CU(cudaEventRecord(ctx->sideEv1, es));
        CU(cudaStreamWaitEvent(s, ctx->sideEv1, 0));
    }
    CU(cudaStreamSynchronize(s));
    if (anySharded) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->lastPhaseMs[2] = ms;
        ctx->lastPhaseMs[3] = 0.0;
        if (ctx->evMidSet) {             // push variant: the share of the pushes in the merge
            CU(cudaEventElapsedTime(&ms, ctx->evPush0, ctx->evMid));
            ctx->lastPhaseMs[3] = ms;
            ctx->evMidSet = false;
        }
    }
    ctx->pending = false;
    ctx->exchanged = false;
    ctx->lastFoldLaunches = launches;
    return MCB200_OK;
}

// non-zero (index, count) pairs of the escape counts of tally set `set` -> ctx->escList, entries
// cleared (scattering every rank's list back rebuilds the sum); only the touched nu-planes are scanned
int esc_compact(mcb200_ctx *ctx, GridState &g, int set, int64_t *nEntries)
{
    unsigned int *Q = set == 1 ? g.escQ2.p : g.escQ.p;
    const int *touched = set == 1 ? g.nuTouched2.p : g.nuTouched.p;
    if (!Q) return fail(ctx, MCB200_ESTATE, "tally set %d not allocated", set);
    const int nb = ctx->cfg.nbins;
    const size_t nR = (size_t)g.nCells + 1;
    cudaStream_t s = ctx->stream;
    std::vector<int> flag(nb + 1, 1);
    if (touched) {
        CU(cudaMemcpyAsync(flag.data(), touched, sizeof(int) * (nb + 1), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    auto ranges = touched_ranges(flag);
    const int blocks = ctx->numSMs * 8;
    CU(ctx->escCount.alloc(1));
    // One scan of the touched planes: the non-zero entries go into the list the previous call sized
    // (with head room) and are NOT cleared yet -- an overflow only counts, and the scan is repeated
    // once with a longer list; then the listed entries are cleared through the list.
    unsigned long long n = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        CU(ctx->escCount.zero(s));
        const unsigned long long cap = ctx->escList.n / 2;
        for (auto &rg : ranges)
            for (int ang = 0; ang <= ctx->cfg.nAngleBins; ++ang) {
                size_t off = nR * ((size_t)rg.first + (size_t)(nb + 1) * (size_t)ang);
                size_t len = (size_t)(rg.second - rg.first + 1) * nR;
                CU(launch_esc_compact(Q, off, len, cap ? ctx->escList.p : nullptr, ctx->escCount.p, cap, 0, blocks, s));
            }
        CU(cudaMemcpyAsync(&n, ctx->escCount.p, 8, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (n <= cap) break;
        CU(ctx->escList.alloc((size_t)(2 * (n + n / 4 + 1024))));
    }
    CU(launch_esc_clear_list(Q, ctx->escList.p, n, blocks, s));
    *nEntries = (int64_t)n;
    return MCB200_OK;
}

// wave-front schedule (wavefront.cu): event kernels + frequency sort + FLY kernel per wave
int run_wavefront(mcb200_ctx *ctx, const TransportArgs &a, bool multi, int64_t mine)
{
    const int nb = ctx->cfg.nbins;
    cudaStream_t s = ctx->stream;
    size_t n = (size_t)mine;
    CU(ctx->wfRecA.alloc(n * wf_rec_bytes())); CU(ctx->wfRecB.alloc(n * wf_rec_bytes()));
    if (multi) { CU(ctx->wfRecXA.alloc(n * wf_recx_bytes())); CU(ctx->wfRecXB.alloc(n * wf_recx_bytes())); }
    CU(ctx->wfFlyKey.alloc(n));
    CU(ctx->wfEv0.alloc(n)); CU(ctx->wfEv1.alloc(n)); CU(ctx->wfEv2.alloc(n)); CU(ctx->wfEv3.alloc(n));
    CU(ctx->wfCounts.alloc(8)); CU(ctx->wfNext.alloc(1));
    CU(ctx->wfHist.alloc(nb + 1)); CU(ctx->wfCursor.alloc(nb + 1));
    CU(ctx->wfHist.zero(s));
    WfArgs w{};
    w.t = a;
    w.recA = reinterpret_cast<PacketRec *>(ctx->wfRecA.p); w.recB = reinterpret_cast<PacketRec *>(ctx->wfRecB.p);
    w.recxA = reinterpret_cast<PacketRecX *>(ctx->wfRecXA.p); w.recxB = reinterpret_cast<PacketRecX *>(ctx->wfRecXB.p);
    w.flyKey = ctx->wfFlyKey.p;
    w.flyCount = ctx->wfCounts.p; w.evCount = ctx->wfCounts.p + 1;
    w.evList[0] = ctx->wfEv0.p; w.evList[1] = ctx->wfEv1.p; w.evList[2] = ctx->wfEv2.p; w.evList[3] = ctx->wfEv3.p;
    w.stepBudget = ctx->stepBudget < 1 ? 1 : ctx->stepBudget;
    w.flyBatch = ctx->flyBatch;
    // compact escape entries: single grid, no viewing angles (the tally element is then known without
    // the direction), no per-packet trace, element index + flag bit in 32 bits
    w.escCompact = (wf_esc_compact_built() && ctx->escCompact && !multi && ctx->cfg.nAngleBins == 0 && !a.fates &&
                    esize(ctx, ctx->grids[0]) < ((size_t)1 << 31)) ? 1 : 0;
    if (a.fates) { CU(ctx->wfSegs.alloc(n)); CU(ctx->wfSegs.zero(s)); w.t.segsArr = ctx->wfSegs.p; }
    w.hist = ctx->wfHist.p; w.cursor = ctx->wfCursor.p; w.nextFlight = ctx->wfNext.p;
    int flyBps = ctx->blocksPerSM > 0 ? ctx->blocksPerSM : wf_fly_blocks_per_sm(multi);
    if (flyBps < 1) flyBps = 1;
    const int flyBlocks = ctx->numSMs * flyBps;
    auto evBlocks = [&](uint64_t cnt) {
        uint64_t b = (cnt + 255) / 256, cap = (uint64_t)ctx->numSMs * 16;
        return (int)(b < 1 ? 1 : (b > cap ? cap : b));
    };
    // wave 0: every packet is emitted.  With option wave0_order the packets are emitted in the
    // order of their first frequency bin (first_nu_kernel replays that draw) straight into recA:
    // 4 B of index per packet are sorted instead of 64 B of record.
    CU(ctx->wfCounts.zero(s));
    w.inList = nullptr; w.inCount = nullptr;
    w.directA = 0;
    if (!a.resCells && mine > 0 && (ctx->wave0Order == 2 || (ctx->wave0Order == 1 && mine >= (1 << 17)))) {
        CU(ctx->sortKey.alloc(n)); CU(ctx->sortOrder.alloc(n));
        CU(ctx->sortHist.alloc(nb + 1)); CU(ctx->sortCursor.alloc(nb + 1));
        CU(launch_order(a, ctx->sortKey.p, ctx->sortHist.p, ctx->sortCursor.p, ctx->sortOrder.p, ctx->numSMs, s));
        w.t.order = ctx->sortOrder.p;
        w.directA = ctx->wave0Exact ? 2 : 1;
        ctx->lastLaunches += 3;
    }
    {
        // pre-ordered: a small grid keeps the grid-stride window (= how far arrival order can
        // deviate from frequency order) below one frequency bin's worth of packets
        int eb = evBlocks((uint64_t)mine), capOrdered = ctx->numSMs * ctx->wave0Blocks;
        if (w.directA && ctx->wave0Blocks > 0 && eb > capOrdered) eb = capOrdered;
        CU(wf_launch_event(w, multi, 0, eb, s));
        if (w.directA == 2) {            // slot i holds packet order[i] (or a hole)
            ctx->wave0Count = (unsigned int)mine;
            CU(cudaMemcpyAsync(w.flyCount, &ctx->wave0Count, sizeof(unsigned int), cudaMemcpyHostToDevice, s));
        }
    }
    ctx->lastLaunches++;
    w.t.order = nullptr;
    bool sorted = w.directA != 0;
    w.directA = 0;
    for (;;) {
        if (!sorted) CU(wf_launch_sort(w, multi, ctx->numSMs, s));
        sorted = false;
        CU(cudaMemsetAsync(w.evCount, 0, 4 * sizeof(unsigned int), s));
        CU(ctx->wfNext.zero(s));
        if (ctx->flyEv.size() < 2 * (size_t)(ctx->lastWaves + 1)) {
            cudaEvent_t e0, e1;
            CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
            ctx->flyEv.push_back(e0); ctx->flyEv.push_back(e1);
        }
        CU(cudaEventRecord(ctx->flyEv[2 * ctx->lastWaves], s));
        CU(wf_launch_fly(w, multi, flyBlocks, s));
        CU(cudaEventRecord(ctx->flyEv[2 * ctx->lastWaves + 1], s));
        ctx->lastLaunches += 4;
        ctx->lastWaves++;
        unsigned int hc[5];
        CU(cudaMemcpyAsync(hc, ctx->wfCounts.p, sizeof(hc), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (w.escCompact && hc[1 + EV_ESCAPE]) {
            // escapes end here: tallied from the compact entries, not carried into the next wave or the tail
            w.inList = w.evList[EV_ESCAPE]; w.inCount = &w.evCount[EV_ESCAPE];
            CU(wf_launch_escape_compact(w, evBlocks(hc[1 + EV_ESCAPE]), s));
            CU(cudaMemsetAsync(&w.evCount[EV_ESCAPE], 0, sizeof(unsigned int), s));
            ctx->lastLaunches++;
            hc[1 + EV_ESCAPE] = 0;
        }
        uint64_t alive = (uint64_t)hc[1] + hc[2] + hc[3] + hc[4];
        if (alive == 0) break;
        if ((int64_t)alive <= ctx->tailThreshold) {
            if (ctx->pdfPending) CU(cudaStreamWaitEvent(s, ctx->pdfReady, 0));
            // thin tail: one persistent kernel carries the remaining packets to completion
            CU(ctx->wfArgsDev.alloc(sizeof(WfArgs)));
            CU(cudaMemcpyAsync(ctx->wfArgsDev.p, &w, sizeof(WfArgs), cudaMemcpyHostToDevice, s));
            TransportArgs ta = w.t;
            ta.n = (long long)alive; ta.order = nullptr; ta.batch = ctx->batch < 1 ? 1 : ctx->batch; ta.aggSteps = 0;
            CU(ctx->nextPacket.zero(s));
            int tb = (int)((alive + 255) / 256);
            int cap = ctx->numSMs * 3;
            CU(launch_transport(ta, multi, tb < cap ? tb : cap, s, reinterpret_cast<const WfArgs *>(ctx->wfArgsDev.p)));
            ctx->lastLaunches++;
            break;
        }
        CU(cudaMemsetAsync(w.flyCount, 0, sizeof(unsigned int), s));
        if (ctx->pdfPending) CU(cudaStreamWaitEvent(s, ctx->pdfReady, 0));   // re-emission needs the PDFs now
        for (int ev = 3; ev >= 0; --ev) {
            if (!hc[1 + ev]) continue;
            w.inList = w.evList[ev]; w.inCount = &w.evCount[ev];
            CU(wf_launch_event(w, multi, ev, evBlocks(hc[1 + ev]), s));
            ctx->lastLaunches++;
        }
        if (hc[1] + hc[2] + hc[4] == 0) break;   // only escapes were pending: nothing can fly any more
    }
    return MCB200_OK;
}

// interpolation_mod.f90:48-81 on the host (ascending axis)
int host_locate(const std::vector<float> &xa, float x)
{
    int n = (int)xa.size();
    if (x > xa[n - 1]) return n;
    if (x < xa[0]) return 0;
    int first = n;                       // first 0-based index with xa > x
    for (int i = 0; i < n; ++i) if (xa[i] > x) { first = i; break; }
    if (first >= n) return 1;
    return first > 1 ? first : 1;
}

// source cells of the resonance-line transfer owned by this rank (photon_mod.f90:187-262)
int build_res_cells(mcb200_ctx *ctx, std::vector<ResCell> &cells, std::vector<unsigned int> &prefix)
{
    long long iCell = 0;
    unsigned long long gid = 0, local = 0;
    cells.clear(); prefix.clear();
    for (int ig = 0; ig < ctx->cfg.nGrids; ++ig) {
        GridState &g = ctx->grids[ig];
        for (int ix = 1; ix <= g.nx; ++ix)
            for (int iy = 1; iy <= g.ny; ++iy)
                for (int iz = 1; iz <= g.nz; ++iz) {
                    ++iCell;
                    int cell = g.hactive[(size_t)(ix - 1) + (size_t)g.nx * ((size_t)(iy - 1) + (size_t)g.ny * (size_t)(iz - 1))];
                    int cnt = (cell > 0 && !g.resLinePackets.empty()) ? g.resLinePackets[cell] : 0;
                    if (cnt <= 0) continue;
                    if (((iCell - (ctx->rank + 1)) % ctx->nranks) == 0) {
                        ResCell rc{};
                        rc.grid = ig + 1; rc.x = (short)ix; rc.y = (short)iy; rc.z = (short)iz;
                        rc.px = g.hx[ix - 1]; rc.py = g.hy[iy - 1]; rc.pz = g.hz[iz - 1];
                        rc.mx = rc.my = rc.mz = -1;
                        if (ig > 0) {
                            GridState &m = ctx->grids[g.motherP - 1];
                            int a = host_locate(m.hx, rc.px);
                            if (a >= 1 && a < m.nx && rc.px > (m.hx[a - 1] + m.hx[a]) / 2.f) a = a + 1;
                            int b = host_locate(m.hy, rc.py);
                            if (b >= 1 && b < m.ny && rc.py > (m.hy[b - 1] + m.hy[b]) / 2.f) b = b + 1;
                            int c = host_locate(m.hz, rc.pz);
                            if (c >= 1 && c < m.nz && rc.pz > (m.hz[c - 1] + m.hz[c]) / 2.f) c = c + 1;
                            rc.mx = (short)a; rc.my = (short)b; rc.mz = (short)c;
                        }
                        rc.gid = gid;
                        cells.push_back(rc);
                        prefix.push_back((unsigned int)local);
                        local += (unsigned long long)cnt;
                    }
                    gid += (unsigned long long)cnt;
                }
    }
    if (local >= (1ull << 31)) return fail(ctx, MCB200_EINVAL, "too many resonance-line packets on one rank");
    prefix.push_back((unsigned int)local);
    return MCB200_OK;
}

int run_transport(mcb200_ctx *ctx, int iStar, int difGrid, const int32_t *cellLoc, int64_t nGlobal,
                  float deltaE, mcb200_counters *out, bool resLines = false)
{
    const mcb200_config &cfg = ctx->cfg;
    if (!ctx->haveCfg) return fail(ctx, MCB200_ESTATE, "mcb200_set_config not called");
    if (!ctx->haveSpectra || !ctx->haveStars) return fail(ctx, MCB200_ESTATE, "spectra/stars not set");
    if (cfg.nAngleBins > 0 && !ctx->haveView) return fail(ctx, MCB200_ESTATE, "viewpoints not set");
    if (nGlobal < 0) return fail(ctx, MCB200_EINVAL, "negative packet count");
    if (nGlobal >= ((int64_t)1 << 32)) return fail(ctx, MCB200_EINVAL, "more than 2^32 packets in one call (packet counts are 32-bit): split the call");
    if (iStar < 0 || iStar > cfg.nStars) return fail(ctx, MCB200_EINVAL, "iStar out of range");
    for (int i = 0; i < cfg.nGrids; ++i) {
        GridState &g = ctx->grids[i];
        if (!g.set) return fail(ctx, MCB200_ESTATE, "grid %d not set", i + 1);
        if (!g.haveOpacity) return fail(ctx, MCB200_ESTATE, "opacity of grid %d not set", i + 1);
        if (cfg.lgDust && !g.scaOpac.p) return fail(ctx, MCB200_ESTATE, "scaOpac of grid %d not set", i + 1);
        if (cfg.lgDust && !g.canScatter.p) return fail(ctx, MCB200_ESTATE, "dust state of grid %d not set", i + 1);
        if (!g.havePdf) return fail(ctx, MCB200_ESTATE, "re-emission PDFs of grid %d not set", i + 1);
        int rc = ensure_estimators(ctx, g);
        if (rc) return rc;
        if (ctx->tallySet == 1) { rc = ensure_second_set(ctx, g); if (rc) return rc; }
    }
    if (ctx->tallySet == 1 && (cfg.lgDebug || ctx->nranks == 1))
        return fail(ctx, MCB200_ESTATE, "tally_set=1 is for multi-rank, non-debug runs");
    // a previous call with a different deltaE must be folded first (single rank), or
    // reduced by the caller (multi rank)
    if (ctx->pending) {
        if (ctx->nranks == 1 && !ctx->deferFold) { int rc = fold_pending(ctx); if (rc) return rc; }
        else if (ctx->exchanged)
            return fail(ctx, MCB200_ESTATE, "pending tallies already exchanged over the ranks: call mcb200_reduce before the next transport");
        else if (ctx->pendingDeltaE != deltaE)
            return fail(ctx, MCB200_ESTATE, "pending tallies with a different deltaE: call mcb200_reduce first");
    }
    int rc = sync_grids(ctx);
    if (rc) return rc;

    // the reference's split over ranks, iteration_mod.f90:477-493
    int64_t load = nGlobal / ctx->nranks, rest = nGlobal % ctx->nranks;
    int64_t mine = load + (ctx->rank < rest ? 1 : 0);
    int64_t first = (int64_t)ctx->rank * load + (ctx->rank < rest ? ctx->rank : rest);
    if (ctx->partCount > 1) {            // contiguous sub-range of this rank's share
        int64_t pl = mine / ctx->partCount, pr = mine % ctx->partCount, pi = ctx->partIndex;
        first += pi * pl + (pi < pr ? pi : pr);
        mine = pl + (pi < pr ? 1 : 0);
    }

    TransportArgs a{};
    DevParams &P = a.P;
    P.nGrids = cfg.nGrids; P.nbins = cfg.nbins; P.nStars = cfg.nStars; P.nAngleBins = cfg.nAngleBins;
    P.totT = cfg.totAngleBinsTheta; P.totP = cfg.totAngleBinsPhi; P.nLines = cfg.nLines;
    P.lgDust = cfg.lgDust; P.lgGas = cfg.lgGas; P.lgSym = cfg.lgSymmetricXYZ; P.lgIso = cfg.lgIsotropic;
    P.lgDebug = cfg.lgDebug; P.lgMultistars = cfg.lgMultistars; P.lgPlane = cfg.lgPlaneIonization;
    P.safeLimit = cfg.lgPlaneIonization ? 5000 : 500000;
    P.planeDist = nullptr;
    if (cfg.lgPlaneIonization) {
        size_t np = (size_t)ctx->grids[0].nx * ctx->grids[0].nz;
        if (ctx->planeDist.n != np) { CU(ctx->planeDist.alloc(np)); CU(ctx->planeDist.zero(ctx->stream)); }
        P.planeDist = ctx->planeDist.p;
    }
    P.dTheta = cfg.dTheta; P.dPhi = cfg.dPhi; P.R_out = cfg.R_out; P.ionEdge1 = cfg.ionEdge1;
    P.nuArray = ctx->nuArray.p; P.gSca = ctx->gSca.p; P.starCdf = ctx->starCdf.p;
    P.starPos = ctx->starPos.p; P.starIdx = ctx->starIdx.p; P.starCell = ctx->starCell.p;
    P.vpPtheta = ctx->vpPtheta.p; P.vpPphi = ctx->vpPphi.p; P.vpTheta = ctx->vpTheta.p; P.vpPhi = ctx->vpPhi.p;
    a.g1 = ctx->hGrids[0];
    a.grids = ctx->dGrids.p;
    a.iStar = iStar;
    if (iStar == 0) {
        if (!cellLoc || difGrid < 1 || difGrid > cfg.nGrids) return fail(ctx, MCB200_EINVAL, "bad diffuse source");
        const GridState &g = ctx->grids[difGrid - 1];
        if (cellLoc[0] < 1 || cellLoc[0] > g.nx || cellLoc[1] < 1 || cellLoc[1] > g.ny || cellLoc[2] < 1 || cellLoc[2] > g.nz)
            return fail(ctx, MCB200_EINVAL, "diffuse source cell out of range");
        a.difGrid = difGrid; a.difX = cellLoc[0]; a.difY = cellLoc[1]; a.difZ = cellLoc[2];
    }
    // Philox keying.  The key is the context seed advanced by the epoch (option "epoch": the host sets it
    // to the Lucy iteration number, so that no two iterations replay the same histories); a star's
    // packets are (global packet index, iStar); the extra diffuse source of cell (gpLoc, cellLoc) gets
    // streams of its own: counter word 3 = 0x80000000 + the cell's linear index, grid in the packet id.
    a.firstId = first; a.n = mine;
    a.seed = ctx->seed + (unsigned long long)ctx->epoch * 0x9E3779B97F4A7C15ull;
    a.rngStream = (unsigned int)iStar; a.pidBase = 0ull;
    if (iStar == 0) {
        const GridState &dg = ctx->grids[difGrid - 1];
        a.rngStream = 0x80000000u + (unsigned int)((cellLoc[0] - 1) + dg.nx * ((cellLoc[1] - 1) + dg.ny * (cellLoc[2] - 1)));
        a.pidBase = (unsigned long long)difGrid << 48;
    }
    a.resCells = nullptr; a.resPrefix = nullptr; a.nResCells = 0;
    if (resLines) {
        if (!cfg.lgGas || !cfg.lgDust) return fail(ctx, MCB200_ESTATE, "resonance-line transfer needs gas and dust (photon_mod.f90:180,910)");
        std::vector<ResCell> cells;
        std::vector<unsigned int> prefix;
        int rcb = build_res_cells(ctx, cells, prefix);
        if (rcb) return rcb;
        mine = (int64_t)prefix.back();
        a.n = mine; a.firstId = 0;
        if (mine > 0) {
            CU(ctx->resCells.upload(cells.data(), cells.size(), ctx->stream));
            CU(ctx->resPrefix.upload(prefix.data(), prefix.size(), ctx->stream));
            a.resCells = ctx->resCells.p; a.resPrefix = ctx->resPrefix.p; a.nResCells = (int)cells.size();
        }
    }
    CU(ctx->nextPacket.alloc(1)); CU(ctx->nextPacket.zero(ctx->stream));
    CU(ctx->counters.alloc(C_COUNT)); CU(ctx->counters.zero(ctx->stream));
    CU(ctx->qphot.alloc(cfg.nbins)); CU(ctx->qphot.zero(ctx->stream));
    CU(ctx->errFlag.alloc(1)); CU(ctx->errFlag.zero(ctx->stream));
    a.nextPacket = ctx->nextPacket.p; a.counters = ctx->counters.p; a.qphotCounts = ctx->qphot.p;
    a.errFlag = ctx->errFlag.p;
    a.fates = nullptr;
    a.segsArr = nullptr;
    if (ctx->trace) {
        CU(ctx->fates.alloc((size_t)4 * (size_t)(mine > 0 ? mine : 1)));
        CU(ctx->fates.zero(ctx->stream));
        a.fates = ctx->fates.p;
    }
    bool multi = cfg.nGrids > 1;
    int bps = ctx->blocksPerSM > 0 ? ctx->blocksPerSM : transport_blocks_per_sm(multi);
    if (bps < 1) bps = 1;
    int blocks = ctx->numSMs * bps;
    int64_t maxUseful = (mine + 255) / 256;
    if (maxUseful < 1) maxUseful = 1;
    if (blocks > maxUseful) blocks = (int)maxUseful;

    // frequency-ordered processing pays when the nu-planes do not all fit in L2 anyway
    size_t tableBytes = 0;
    for (auto &g : ctx->grids) tableBytes += tsize(ctx, g) * 12;       // opacity 4 B + JsteQ 8 B
    bool ordered = !resLines && (ctx->orderMode == 1 || (ctx->orderMode < 0 && tableBytes > (size_t)48 << 20 && mine >= (1 << 16)));
    if (mine >= ((int64_t)1 << 32)) return fail(ctx, MCB200_EINVAL, "more than 2^32 packets per rank in one call: split the call");
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    a.order = nullptr;
    a.batch = ctx->batch < 1 ? 1 : ctx->batch;
    a.aggSteps = 0;
    bool waveSel = ctx->waveMode == 1 || (ctx->waveMode < 0 && mine >= (1 << 17));
    if (ordered && mine > 0 && !waveSel) {
        CU(ctx->sortKey.alloc((size_t)mine)); CU(ctx->sortOrder.alloc((size_t)mine));
        CU(ctx->sortHist.alloc(cfg.nbins + 1)); CU(ctx->sortCursor.alloc(cfg.nbins + 1));
        CU(launch_order(a, ctx->sortKey.p, ctx->sortHist.p, ctx->sortCursor.p, ctx->sortOrder.p, ctx->numSMs, ctx->stream));
        a.order = ctx->sortOrder.p;
        a.aggSteps = ctx->aggSteps;
    }
    bool wave = ctx->waveMode == 1 || (ctx->waveMode < 0 && mine >= (1 << 17));
    ctx->lastWaves = 0; ctx->lastLaunches = (ordered && mine > 0 && !wave ? 3 : 0) + (mine > 0 && !wave ? 1 : 0);
    if (mine > 0 && wave) {
        if (mine >= ((int64_t)1 << 31)) return fail(ctx, MCB200_EINVAL, "more than 2^31 packets per rank in one wave-front call: split the call");
        a.order = nullptr;
        int rcw = run_wavefront(ctx, a, multi, mine);
        if (rcw) return rcw;
    } else if (mine > 0) {
        if (ctx->pdfPending) CU(cudaStreamWaitEvent(ctx->stream, ctx->pdfReady, 0));
        CU(launch_transport(a, multi, blocks, ctx->stream, nullptr));
    }
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));

    if (ctx->pdfPending) {                                // deferred verdict of the async PDF upload
        CU(cudaStreamSynchronize(ctx->copyStream));
        std::vector<int> bad(cfg.nGrids, 0);
        CU(cudaMemcpy(bad.data(), ctx->pdfBad.p, sizeof(int) * cfg.nGrids, cudaMemcpyDeviceToHost));
        ctx->pdfPending = false;
        for (int i = 0; i < cfg.nGrids; ++i)
            if (bad[i]) { ctx->grids[i].havePdf = false; return fail(ctx, MCB200_ETABLE, "grid %d: re-emission PDF is not non-decreasing along nu", i + 1); }
    }
    unsigned long long hc[C_COUNT];
    int herr = 0;
    CU(cudaMemcpy(hc, ctx->counters.p, sizeof(hc), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&herr, ctx->errFlag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (out) {
        std::vector<unsigned long long> q(cfg.nbins);
        CU(cudaMemcpy(q.data(), ctx->qphot.p, sizeof(unsigned long long) * cfg.nbins, cudaMemcpyDeviceToHost));
        double Q = 0.0;
        for (int i = 0; i < cfg.nbins; ++i)
            if (q[i]) Q += (double)q[i] * ((double)deltaE / (2.1799153e-11 * (double)ctx->hNu[i]));
        out->nPackets = mine;
        out->nAbs = (int64_t)hc[C_ABS]; out->nSca = (int64_t)hc[C_SCA]; out->trapped = (int64_t)hc[C_TRAPPED];
        out->nLinePackets = (int64_t)hc[C_LINE]; out->nDropped = (int64_t)hc[C_DROPPED];
        out->nSegments = (int64_t)hc[C_SEGMENTS]; out->nFlights = (int64_t)hc[C_FLIGHTS];
        out->nEscaped = (int64_t)hc[C_ESCAPED]; out->nEarlyEscaped = (int64_t)hc[C_EARLY];
        out->Qphot = Q;
        out->kernel_ms = ms;
        out->total_ms = ms;
        out->nLaunches = ctx->lastLaunches;
        out->nWaves = ctx->lastWaves;
        out->fly_ms = 0.0;
        for (int wv = 0; wv < ctx->lastWaves && 2 * (size_t)wv + 1 < ctx->flyEv.size(); ++wv) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, ctx->flyEv[2 * wv], ctx->flyEv[2 * wv + 1]) == cudaSuccess) out->fly_ms += t;
        }
    }
    ctx->pending = true;
    if (ctx->tallySet == 1) ctx->pending2 = true;
    ctx->pendingDeltaE = deltaE;
    if (ctx->sedLocal && ctx->nranks > 1 && !herr) {
        // the caller exchanges sedQ (a few KB) instead of the escape counts (5 GB at 128^3 x 600)
        int nl = 0;
        int rcs = sed_tally(ctx, ctx->tallySet, &nl);
        if (rcs) return rcs;
        ctx->sedReady = true;
        if (out) out->nLaunches += nl;
    }
    if (herr) return fail(ctx, MCB200_EPACKET, "a packet hit reference stop condition %d (see oracle/mc_oracle.c ERR_STOP codes)", herr);
    if (ctx->nranks == 1 && !ctx->deferFold) {
        int rcf = fold_pending(ctx);
        if (rcf) return rcf;
        CU(cudaEventRecord(ctx->ev2, ctx->stream));
        CU(cudaEventSynchronize(ctx->ev2));
        float ms2 = 0.f;
        CU(cudaEventElapsedTime(&ms2, ctx->ev0, ctx->ev2));
        if (out) { out->total_ms = ms2; out->nLaunches += ctx->lastFoldLaunches; }
    }
    return MCB200_OK;
}

__global__ void detmath_kernel(int which, const float *in, float *out, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s, c;
    switch (which) {
    case 0: out[i] = dm_logf(in[i]); break;
    case 1: dm_sincosf(in[i], s, c); out[i] = s; break;
    case 2: dm_sincosf(in[i], s, c); out[i] = c; break;
    case 3: out[i] = dm_acosf(in[i]); break;
    case 4: out[i] = dm_atanf(in[i]); break;
    case 5: out[i] = dm_expf(in[i]); break;
    // 6: the transport's division (div_rn) of in[i] by in[(i+1) mod n]; 7: the IEEE divider
    case 6: { float v = in[(i + 1) % n]; out[i] = div_rn(in[i], v, 1.f / fabsf(v)); break; }
    case 7: out[i] = in[i] / in[(i + 1) % n]; break;
    default: out[i] = 0.f;
    }
}

__global__ void uniforms_kernel(unsigned long long seed, unsigned long long pid, unsigned int stream, int n, float *out)
{
    if (threadIdx.x || blockIdx.x) return;
    Rng r;
    r.init(seed, pid, stream);
    for (int i = 0; i < n; ++i) out[i] = r.uniform();
}


// in-place sum / max over ranks of `count` elements on the library stream
int comm_allreduce(mcb200_ctx *ctx, void *buf, size_t count, int dtype, int op)
{
    if (!count) return MCB200_OK;
    NC(nccl_api().AllReduce(buf, buf, count, dtype, op, ctx->comm, ctx->stream));
    ctx->lastExchangeBytes += (int64_t)count * (dtype == kNcclUint64 ? 8 : 4);
    return MCB200_OK;
}

// escape counts of one grid: all-gather of every rank's non-zero (index, count) pairs when that
// moves fewer bytes than the dense all-reduce of the touched planes (PacketEngine._exchange_escaped_sparse)
int comm_exchange_escaped(mcb200_ctx *ctx, GridState &g, const std::vector<std::pair<int, int>> &ranges, int64_t preCompacted = -1)
{
    NcclApi &N = nccl_api();
    const int nb = ctx->cfg.nbins, world = ctx->nranks;
    const size_t nR = (size_t)g.nCells + 1;
    cudaStream_t s = ctx->stream;
    size_t planes = 0;
    for (auto &rg : ranges) planes += (size_t)(rg.second - rg.first + 1);
    const unsigned long long dense = 2ull * 4ull * nR * (unsigned long long)(ctx->cfg.nAngleBins + 1) * planes;
    int64_t n = preCompacted;
    int rc = 0;
    if (n < 0) {
        rc = esc_compact(ctx, g, 0, &n);
        if (rc) return rc;
    }
    CU(ctx->commSizes.alloc((size_t)world + 1));
    unsigned long long mine = (unsigned long long)n;
    CU(cudaMemcpyAsync(ctx->commSizes.p + world, &mine, 8, cudaMemcpyHostToDevice, s));
    NC(N.AllGather(ctx->commSizes.p + world, ctx->commSizes.p, 1, kNcclUint64, ctx->comm, s));
    std::vector<unsigned long long> all(world);
    CU(cudaMemcpyAsync(all.data(), ctx->commSizes.p, 8 * (size_t)world, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    trace_mark(ctx, "exchange: escape lists compacted, sizes gathered");
    unsigned long long maxn = 0;
    for (auto v : all) maxn = v > maxn ? v : maxn;
    if (ctx->exchangeDense || 16ull * maxn * (unsigned long long)world >= dense) {
        // lists too long: put this rank's entries back and all-reduce the touched planes
        CU(launch_esc_scatter(g.escQ.p, g.escQ.n, ctx->escList.p, (unsigned long long)n, ctx->numSMs * 8, s));
        for (auto &rg : ranges)
            for (int ang = 0; ang <= ctx->cfg.nAngleBins; ++ang) {
                size_t off = nR * ((size_t)rg.first + (size_t)(nb + 1) * (size_t)ang);
                rc = comm_allreduce(ctx, g.escQ.p + off, (size_t)(rg.second - rg.first + 1) * nR, kNcclUint32, kNcclSum);
                if (rc) return rc;
            }
        return MCB200_OK;
    }
    ctx->lastExchangeSparse++;
    if (maxn == 0) return MCB200_OK;
    // equal-sized slots, zero padded: a (0, 0) pair adds nothing
    if (ctx->commPad.n < 2 * maxn) CU(ctx->commPad.alloc((size_t)(2 * maxn)));
    if (ctx->commGather.n < 2 * maxn * (unsigned long long)world) CU(ctx->commGather.alloc((size_t)(2 * maxn * world)));
    CU(cudaMemsetAsync(ctx->commPad.p, 0, 16 * (size_t)maxn, s));
    if (n) CU(cudaMemcpyAsync(ctx->commPad.p, ctx->escList.p, 16 * (size_t)n, cudaMemcpyDeviceToDevice, s));
    NC(N.AllGather(ctx->commPad.p, ctx->commGather.p, (size_t)(2 * maxn), kNcclUint64, ctx->comm, s));
    ctx->lastExchangeBytes += (int64_t)(16 * maxn);
    CU(launch_esc_scatter(g.escQ.p, g.escQ.n, ctx->commGather.p, maxn * (unsigned long long)world, ctx->numSMs * 8, s));
    return MCB200_OK;
}

// Upload of a (0:nCells, nbins) host table every rank holds identically (the reference's ranks do,
// after their MPI_ALLREDUCE of recPDFTemp, iteration_mod.f90:357-424): with option "pdf_slabs" each
// rank sends only its 1/nranks slab of nu-planes over PCIe and the slabs are all-gathered over
// NVLink into the staging buffer (planes padded to a multiple of nranks).  `cs`: stream to work on.
int upload_table_slabs(mcb200_ctx *ctx, GridState &g, const float *src, cudaStream_t cs)
{
    const size_t nR = (size_t)g.nCells + 1;
    const int nb = ctx->cfg.nbins, world = ctx->nranks, rank = ctx->rank;
    const bool slabs = ctx->pdfSlabs && world > 1 && ctx->comm && !ctx->solo;
    if (!slabs) {
        CU(g.stage.alloc(nR * (size_t)nb));
        CU(cudaMemcpyAsync(g.stage.p, src, nR * (size_t)nb * sizeof(float), cudaMemcpyHostToDevice, cs));
        ctx->lastPdfH2D = (int64_t)(nR * (size_t)nb * sizeof(float));
        return MCB200_OK;
    }
    const int per = (nb + world - 1) / world;                 // planes per rank (the last slabs may be short or empty)
    CU(g.stage.alloc(nR * (size_t)per * (size_t)world));
    const int p0 = rank * per, p1 = (p0 + per < nb ? p0 + per : nb);
    size_t mine = p1 > p0 ? (size_t)(p1 - p0) * nR : 0;
    if (mine) CU(cudaMemcpyAsync(g.stage.p + (size_t)p0 * nR, src + (size_t)p0 * nR, mine * sizeof(float), cudaMemcpyHostToDevice, cs));
    ctx->lastPdfH2D = (int64_t)(mine * sizeof(float));
    NC(nccl_api().AllGather(g.stage.p + (size_t)p0 * nR, g.stage.p, (size_t)per * nR, kNcclFloat32, ctx->comm, cs));
    return MCB200_OK;
}

int comm_exchange(mcb200_ctx *ctx)
{
    const int nb = ctx->cfg.nbins, blocks = ctx->numSMs * 8;
    cudaStream_t s = ctx->stream;
    ctx->lastExchangeBytes = 0;
    ctx->lastExchangeSparse = 0;
    if (ctx->pending2) {                  // two tally sets: merge as integers first, exchange the sum
        for (auto &g : ctx->grids) {
            if (!g.JsteQ2.p) continue;
            CU(launch_merge_sets(g.JsteQ.p, g.JsteQ2.p, g.JsteQ.n, g.escQ.p, g.escQ2.p, g.escQ.n,
                                 g.nuTouched.p, g.nuTouched2.p, nb + 1, blocks, s));
        }
        ctx->pending2 = false;
    }
    if (ctx->sedLocal) {                  // per-(nu, angle) escape counts instead of the per-cell array
        if (!ctx->sedReady) { int rc = sed_tally(ctx, 0, nullptr); if (rc) return rc; ctx->sedReady = true; }
        int rc = comm_allreduce(ctx, ctx->sedQ.p, ctx->sedQ.n, kNcclUint64, kNcclSum);
        if (rc) return rc;
    }
    bool pushed = false;
    trace_mark(ctx, "exchange: start", true);
    for (size_t ig = 0; ig < ctx->grids.size(); ++ig) {
        GridState &g = ctx->grids[ig];
        if (!g.set || !g.JsteQ.p) continue;
        const size_t nR = (size_t)g.nCells + 1;
        int rc = 0;
        // Peer-memory push path: this rank's escape counts are compacted FIRST -- a scan of its own touched
        // planes that needs nothing from the peers, so it runs while the slower ranks are still transporting,
        // and it is out of the way of the push kernels (which would otherwise hold every SM until they end:
        // measured 7.4 ms from "pushes issued" to "lists compacted" at N=2, against ~1 ms for the scan alone)
        int64_t escN = -1;
        if (!ctx->sedLocal && !ctx->exchangeAllReduce && ctx->nranks > 1 && ctx->p2pMode != 0 && !ctx->cfg.lgDebug &&
            ctx->p2pState == 1 && ctx->p2pPush) {
            rc = esc_compact(ctx, g, 0, &escN);
            if (rc) return rc;
            trace_mark(ctx, "exchange: own escape counts compacted");
        }
        // which nu-planes did any rank touch
        rc = comm_allreduce(ctx, g.nuTouched.p, g.nuTouched.n, kNcclInt32, kNcclMax);
        if (rc) return rc;
        std::vector<int> flag(nb + 1, 0);
        CU(cudaMemcpyAsync(flag.data(), g.nuTouched.p, sizeof(int) * (nb + 1), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        trace_mark(ctx, "exchange: flags max-reduced (ranks in step)");
        auto ranges = touched_ranges(flag);
        g.jShards.clear();
        g.jShardsP2P = false;
        const bool allreduce = ctx->exchangeAllReduce || ctx->nranks == 1;
        bool p2p = false;
        if (!allreduce && ctx->p2pMode != 0 && !ctx->cfg.lgDebug && ctx->p2pState >= 0) {
            rc = p2p_setup(ctx, g);
            if (rc) return rc;
            p2p = ctx->p2pState == 1;
        }
        if (!allreduce && ctx->p2pMode == 1 && !p2p)
            return fail(ctx, MCB200_ECOMM, "option exchange_p2p=1: peer memory unavailable (%s)",
                        ctx->cfg.lgDebug ? "debug tallies go through NCCL" : ctx->lastP2PWhy.c_str());
        ctx->lastExchangePath = allreduce ? 1 : (p2p ? 3 : 2);
        const bool escLater = p2p && ctx->p2pPush;   // push variant: beside the pushes, see below
        if (!ctx->sedLocal && !escLater) {
            // the escape counts first: their exchange has data-dependent sizes (host round trips), which
            // must not queue up behind the bulk transfer of the J planes on the same stream
            rc = comm_exchange_escaped(ctx, g, ranges, escN);
            if (rc) return rc;
        }
        NC(nccl_api().GroupStart());
        for (auto &rg : ranges) {
            int p0 = rg.first < 1 ? 1 : rg.first, p1 = rg.second;
            if (p1 < p0) continue;
            size_t off = (size_t)(p0 - 1) * nR, len = (size_t)(p1 - p0 + 1) * nR;
            const int nSets = ctx->cfg.lgDebug && g.JdifQ.p ? 2 : 1;
            if (allreduce) {
                for (int set = 0; set < nSets; ++set) {
                    rc = comm_allreduce(ctx, (set ? g.JdifQ.p : g.JsteQ.p) + off, len, kNcclUint64, kNcclSum);
                    if (rc) return rc;
                }
                continue;
            }
            // every rank owns the sum of elements [off + r*count, +count) of the range: delivered by a
            // reduce-scatter in place, or left to the fused peer-memory kernel of the fold; the
            // < nranks elements left over at the end of the range are all-reduced
            const size_t count = len / (size_t)ctx->nranks, tail = len - count * (size_t)ctx->nranks;
            for (int set = 0; set < nSets; ++set) {
                unsigned long long *Q = (set ? g.JdifQ.p : g.JsteQ.p) + off;
                if (count && !p2p) {
                    NC(nccl_api().ReduceScatter(Q, Q + (size_t)ctx->rank * count, count, kNcclUint64, kNcclSum, ctx->comm, s));
                    ctx->lastExchangeBytes += (int64_t)(count * (size_t)ctx->nranks) * 8;
                }
                rc = comm_allreduce(ctx, Q + count * (size_t)ctx->nranks, tail, kNcclUint64, kNcclSum);
                if (rc) return rc;
            }
            g.jShards.push_back({off, len, count});
        }
        NC(nccl_api().GroupEnd());
        trace_mark(ctx, "exchange: range tails issued");
        g.jShardsP2P = p2p && !g.jShards.empty();
        if (g.jShardsP2P && ctx->p2pPush) {
            // the bulk of the exchange starts now, on the side stream (copy engines + links) ...
            rc = p2p_push_phase(ctx, g, !pushed);
            if (rc) return rc;
            pushed = true;
            trace_mark(ctx, "exchange: pushes issued (side stream)");
        }
        // ... beside the escape counts on the library stream
        if (!ctx->sedLocal && escLater) {
            rc = comm_exchange_escaped(ctx, g, ranges, escN);
            if (rc) return rc;
            trace_mark(ctx, "exchange: escape counts issued");
        }
        if (ctx->cfg.lgDebug && g.lineQ.n) {
            rc = comm_allreduce(ctx, g.lineQ.p, g.lineQ.n, kNcclUint32, kNcclSum);
            if (rc) return rc;
        }
        if (ig == 0 && ctx->cfg.lgPlaneIonization && ctx->planeDist.n) {
            rc = comm_allreduce(ctx, ctx->planeDist.p, ctx->planeDist.n, kNcclInt32, kNcclSum);
            if (rc) return rc;
        }
    }
    CU(cudaStreamSynchronize(s));
    trace_mark(ctx, "exchange: library stream drained");
    return MCB200_OK;
}

}  // namespace

extern "C" {

int mcb200_create(mcb200_ctx **pctx, int32_t device, int32_t rank, int32_t nranks, uint64_t seed)
{
    if (!pctx || nranks < 1 || rank < 0 || rank >= nranks) return MCB200_EINVAL;
    *pctx = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return MCB200_ENODEV;
    if (device < 0 || device >= ndev) return MCB200_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return MCB200_ENODEV;
    mcb200_ctx *ctx = new mcb200_ctx();
    ctx->device = device; ctx->rank = rank; ctx->nranks = nranks; ctx->seed = seed;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return MCB200_ENODEV; }
    ctx->numSMs = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MCB200_ENODEV; }
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaEventCreate(&ctx->ev2);
    cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->pdfReady, cudaEventDisableTiming);
    *pctx = ctx;
    return MCB200_OK;
}

int mcb200_destroy(mcb200_ctx *ctx)
{
    if (!ctx) return MCB200_EINVAL;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->pdfReady) cudaEventDestroy(ctx->pdfReady);
    if (ctx->copyStream) { cudaStreamSynchronize(ctx->copyStream); cudaStreamDestroy(ctx->copyStream); }
    if (ctx->sideStream) { cudaStreamSynchronize(ctx->sideStream); cudaStreamDestroy(ctx->sideStream); }
    if (ctx->sideEv0) cudaEventDestroy(ctx->sideEv0);
    if (ctx->sideEv1) cudaEventDestroy(ctx->sideEv1);
    if (ctx->evMid) cudaEventDestroy(ctx->evMid);
    if (ctx->evPush0) cudaEventDestroy(ctx->evPush0);
    if (ctx->pushDone) cudaEventDestroy(ctx->pushDone);
    for (auto e : ctx->flyEv) cudaEventDestroy(e);
    if (ctx->sparseHost) cudaFreeHost(ctx->sparseHost);
    for (auto &g : ctx->grids) p2p_close(g, ctx->rank);
    if (ctx->comm) { nccl_api().CommDestroy(ctx->comm); ctx->comm = nullptr; }
    ctx->grids.clear();
    cudaStream_t s = ctx->stream;
    delete ctx;
    if (s) cudaStreamDestroy(s);
    return MCB200_OK;
}

const char *mcb200_last_error(const mcb200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int mcb200_set_config(mcb200_ctx *ctx, const mcb200_config *cfg)
{
    NEED_CTX();
    if (!cfg) return fail(ctx, MCB200_EINVAL, "null config");
    if (cfg->nGrids < 1 || cfg->nbins < 3 || cfg->nStars < 0 || cfg->nAngleBins < 0)
        return fail(ctx, MCB200_EINVAL, "bad sizes in config");
    if (cfg->lgPlaneIonization && cfg->lgSymmetricXYZ)
        return fail(ctx, MCB200_EINVAL, "lgSymmetricXYZ and lgPlaneIonization flags both raised (photon_mod.f90:2675-2678)");
    if (cfg->totAngleBinsTheta < 1 || cfg->totAngleBinsPhi < 1 || !(cfg->dTheta > 0.f) || !(cfg->dPhi > 0.f))
        return fail(ctx, MCB200_EINVAL, "bad angle bins in config");
    ctx->cfg = *cfg;
    ctx->grids.clear();
    ctx->grids.resize(cfg->nGrids);
    ctx->haveCfg = true;
    ctx->gridsDirty = true;
    ctx->pending = false;
    ctx->exchanged = false;
    return MCB200_OK;
}

int mcb200_set_grid(mcb200_ctx *ctx, int32_t iG, int32_t nx, int32_t ny, int32_t nz, int32_t nCells,
                    int32_t motherP, const float *xAxis, const float *yAxis, const float *zAxis,
                    const int32_t *active)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g) return fail(ctx, MCB200_EINVAL, "grid index %d out of range (set_config first)", iG);
    if (nx < 2 || ny < 2 || nz < 2 || nCells < 0 || !xAxis || !yAxis || !zAxis || !active)
        return fail(ctx, MCB200_EINVAL, "bad grid arguments");
    if (iG > 1 && motherP != 1) return fail(ctx, MCB200_EUNSUPPORTED, "nested sub-grids (motherP != 1) are not supported (photon_mod.f90:2551-2554)");
    g->nx = nx; g->ny = ny; g->nz = nz; g->nCells = nCells; g->motherP = motherP;
    g->hx.assign(xAxis, xAxis + nx); g->hy.assign(yAxis, yAxis + ny); g->hz.assign(zAxis, zAxis + nz);
    for (auto *ax : {&g->hx, &g->hy, &g->hz})
        for (size_t i = 1; i < ax->size(); ++i)
            if (!((*ax)[i] > (*ax)[i - 1])) return fail(ctx, MCB200_EINVAL, "grid %d: axes must be strictly ascending", iG);
    // operand range of the transport's division (div_rn, transport_core.cuh): coordinates in cm,
    // zero or 1e-10 <= |x| <= 1e27, spacings >= 1e-10
    for (auto *ax : {&g->hx, &g->hy, &g->hz})
        for (size_t i = 0; i < ax->size(); ++i) {
            float v = std::fabs((*ax)[i]);
            if (!(v == 0.f || (v >= 1.e-10f && v <= 1.e27f)) || (i > 0 && !((*ax)[i] - (*ax)[i - 1] >= 1.e-10f)))
                return fail(ctx, MCB200_EUNSUPPORTED, "grid %d: axis coordinates must be 0 or within 1e-10..1e27 cm, spacings >= 1e-10 cm", iG);
        }
    size_t nTot = (size_t)nx * ny * nz;
    if (nTot >= ((size_t)1 << 31) || nx > 32000 || ny > 32000 || nz > 32000) return fail(ctx, MCB200_EINVAL, "grid %d too large for 32-bit cell indexing", iG);
    g->hactive.assign(active, active + nTot);
    // validate and detect the dense numbering (k fastest, grid_mod.f90:1227-1262)
    int dense = 1;
    for (int z = 0; z < nz && dense >= 0; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                int a = active[(size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * (size_t)z)];
                if (a > nCells) return fail(ctx, MCB200_EINVAL, "grid %d: active id %d > nCells", iG, a);
                if (a < 0 && (iG != 1 || -a > ctx->cfg.nGrids || -a < 2))
                    return fail(ctx, MCB200_EINVAL, "grid %d: bad sub-grid pointer %d in active", iG, a);
                if (a != 1 + z + nz * (y + ny * x)) dense = 0;
            }
    g->dense = dense;
    g->geo[0] = (g->hx[nx - 1] - g->hx[nx - 2]) / 2.f;
    g->geo[1] = (g->hy[ny - 1] - g->hy[ny - 2]) / 2.f;
    g->geo[2] = (g->hz[nz - 1] - g->hz[nz - 2]) / 2.f;
    // path-length quantum: smallest cell width / 2^24, rounded down to a power of two -- but not below the
    // largest cell width / 2^33, so that on strongly graded axes (benchmarks/dust/2D: spacing ratio 10^5) a
    // (cell, nu) element still holds > 6x10^7 diagonal crossings of the widest cell before its sum reaches 2^63
    double wmin = 1e300, wmax = 0.0;
    for (auto *ax : {&g->hx, &g->hy, &g->hz})
        for (size_t i = 1; i < ax->size(); ++i) {
            const double w = (double)(*ax)[i] - (double)(*ax)[i - 1];
            wmin = std::fmin(wmin, w);
            wmax = std::fmax(wmax, w);
        }
    wmin *= 0.5;
    wmax *= 0.5;
    {
        const int eFine = (int)std::floor(std::log2(wmin)) - 24, eCap = (int)std::floor(std::log2(wmax)) - 33;
        g->lenExp = eFine > eCap ? eFine : eCap;
    }
    cudaStream_t s = ctx->stream;
    CU(g->xAxis.upload(g->hx.data(), nx, s)); CU(g->yAxis.upload(g->hy.data(), ny, s)); CU(g->zAxis.upload(g->hz.data(), nz, s));
    auto wx = make_walls(g->hx), wy = make_walls(g->hy), wz = make_walls(g->hz);
    CU(g->xWall.upload(wx.data(), wx.size(), s)); CU(g->yWall.upload(wy.data(), wy.size(), s)); CU(g->zWall.upload(wz.data(), wz.size(), s));
    CU(g->active.upload(g->hactive.data(), nTot, s));
    // dV(0:nCells) in 1e45 cm^3 (photon_mod.f90:1469-1512)
    bool sym = ctx->cfg.lgSymmetricXYZ != 0;
    auto dxs = make_widths(g->hx, sym), dys = make_widths(g->hy, sym), dzs = make_widths(g->hz, sym);
    std::vector<float> dV((size_t)nCells + 1, 1.f);
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                int a = active[(size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * (size_t)z)];
                if (a > 0) {
                    volatile float v = dxs[x] * dys[y];
                    dV[a] = v * dzs[z];
                }
            }
    CU(g->dV.upload(dV.data(), dV.size(), s));
    CU(cudaStreamSynchronize(s));
    g->set = true;
    g->haveOpacity = false;
    g->havePdf = false;
    // (re)size estimator storage lazily
    g->JsteQ.release(); g->Jste.release(); g->escQ.release(); g->esc.release();
    g->JdifQ.release(); g->Jdif.release(); g->lineQ.release(); g->linePk.release();
    ctx->gridsDirty = true;
    return MCB200_OK;
}

int mcb200_set_spectra(mcb200_ctx *ctx, const float *nuArray, const float *gSca, const float *inSpectrumProbDen)
{
    NEED_CTX();
    if (!ctx->haveCfg) return fail(ctx, MCB200_ESTATE, "set_config first");
    if (!nuArray || !inSpectrumProbDen) return fail(ctx, MCB200_EINVAL, "null spectra");
    int nb = ctx->cfg.nbins, ns = ctx->cfg.nStars + 1;
    ctx->hNu.assign(nuArray, nuArray + nb);
    CU(ctx->nuArray.upload(nuArray, nb, ctx->stream));
    if (gSca) CU(ctx->gSca.upload(gSca, nb, ctx->stream));
    else if (ctx->cfg.lgDust && !ctx->cfg.lgIsotropic) return fail(ctx, MCB200_EINVAL, "gSca required with dust");
    // Fortran (0:nStars, nbins), star fastest -> rows [s][nu]; rows s>=1 must be non-decreasing
    std::vector<float> rows((size_t)ns * nb);
    for (int s = 0; s < ns; ++s)
        for (int i = 0; i < nb; ++i) rows[(size_t)s * nb + i] = inSpectrumProbDen[(size_t)s + (size_t)ns * i];
    // Uniforms are < 1, so only min(cdf,1) matters to getNu2's scan: setProbDen forces just
    // the entries >= max to 1 (continuum_mod.f90:454-467) and a float32 running sum that
    // overshoots 1 legitimately leaves e.g. [.., 1.0000001, 1, 1] behind.
    for (int s = 1; s < ns; ++s)
        for (int i = 1; i < nb; ++i)
            if (!(std::fmin(rows[(size_t)s * nb + i], 1.f) >= std::fmin(rows[(size_t)s * nb + i - 1], 1.f)))
                return fail(ctx, MCB200_ETABLE, "inSpectrumProbDen(%d,:) is not non-decreasing at bin %d", s, i + 1);
    CU(ctx->starCdf.upload(rows.data(), rows.size(), ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->haveSpectra = true;
    return MCB200_OK;
}

int mcb200_set_stars(mcb200_ctx *ctx, const float *starPosition, const int32_t *starIndeces)
{
    NEED_CTX();
    if (!ctx->haveCfg) return fail(ctx, MCB200_ESTATE, "set_config first");
    int ns = ctx->cfg.nStars;
    if (ns > 0 && (!starPosition || !starIndeces)) return fail(ctx, MCB200_EINVAL, "null stars");
    std::vector<int> idx((size_t)4 * (ns > 0 ? ns : 1), 1), cells(ns > 0 ? ns : 1, 0);
    std::vector<float> pos((size_t)3 * (ns > 0 ? ns : 1), 0.f);
    for (int i = 0; i < ns; ++i) {
        for (int k = 0; k < 4; ++k) idx[4 * i + k] = starIndeces[(size_t)i + (size_t)ns * k];
        for (int k = 0; k < 3; ++k) pos[3 * i + k] = starPosition[3 * i + k];
        int gP = idx[4 * i + 3];
        GridState *g = grid_of(ctx, gP);
        if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "star %d: grid %d not set", i + 1, gP);
        int x = idx[4 * i], y = idx[4 * i + 1], z = idx[4 * i + 2];
        if (x < 1 || x > g->nx || y < 1 || y > g->ny || z < 1 || z > g->nz)
            return fail(ctx, MCB200_EINVAL, "star %d: indices outside grid %d", i + 1, gP);
        cells[i] = g->hactive[(size_t)(x - 1) + (size_t)g->nx * ((size_t)(y - 1) + (size_t)g->ny * (size_t)(z - 1))];
    }
    ctx->hStarPos = pos; ctx->hStarIdx = idx;
    CU(ctx->starPos.upload(pos.data(), pos.size(), ctx->stream));
    CU(ctx->starIdx.upload(idx.data(), idx.size(), ctx->stream));
    CU(ctx->starCell.upload(cells.data(), cells.size(), ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->haveStars = true;
    return MCB200_OK;
}

int mcb200_set_viewpoints(mcb200_ctx *ctx, const int32_t *pT, const int32_t *pP, const float *vT, const float *vP)
{
    NEED_CTX();
    if (!ctx->haveCfg) return fail(ctx, MCB200_ESTATE, "set_config first");
    if (!pT || !pP || !vT || !vP) return fail(ctx, MCB200_EINVAL, "null viewpoint tables");
    const mcb200_config &c = ctx->cfg;
    for (int i = 0; i <= c.totAngleBinsTheta; ++i) if (pT[i] < 0 || pT[i] > c.nAngleBins) return fail(ctx, MCB200_EINVAL, "viewPointPtheta out of range");
    for (int i = 0; i <= c.totAngleBinsPhi; ++i) if (pP[i] < 0 || pP[i] > c.nAngleBins) return fail(ctx, MCB200_EINVAL, "viewPointPphi out of range");
    CU(ctx->vpPtheta.upload(pT, c.totAngleBinsTheta + 1, ctx->stream));
    CU(ctx->vpPphi.upload(pP, c.totAngleBinsPhi + 1, ctx->stream));
    CU(ctx->vpTheta.upload(vT, c.nAngleBins + 1, ctx->stream));
    CU(ctx->vpPhi.upload(vP, c.nAngleBins + 1, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->haveView = true;
    return MCB200_OK;
}

int mcb200_set_dust_species(mcb200_ctx *ctx, const int32_t *nSpeciesPart, const float *grainAbun,
                            const int32_t *dustComPoint, const float *TdustSublime, int32_t nSpecies)
{
    NEED_CTX();
    if (!ctx->haveCfg) return fail(ctx, MCB200_ESTATE, "set_config first");
    const mcb200_config &c = ctx->cfg;
    if (!nSpeciesPart || !grainAbun || !dustComPoint || !TdustSublime || nSpecies < 1 || c.nDustComp < 1 || c.nSpeciesMax < 1)
        return fail(ctx, MCB200_EINVAL, "bad dust species tables");
    ctx->nSpeciesPart.assign(nSpeciesPart, nSpeciesPart + c.nDustComp);
    ctx->dustComPoint.assign(dustComPoint, dustComPoint + c.nDustComp);
    ctx->grainAbun.assign(grainAbun, grainAbun + (size_t)c.nDustComp * c.nSpeciesMax);
    ctx->TdustSublime.assign(TdustSublime, TdustSublime + nSpecies);
    for (int k = 0; k < c.nDustComp; ++k)
        if (nSpeciesPart[k] < 0 || nSpeciesPart[k] > c.nSpeciesMax || dustComPoint[k] < 1 || dustComPoint[k] - 1 + nSpeciesPart[k] > nSpecies)
            return fail(ctx, MCB200_EINVAL, "dust component %d inconsistent", k + 1);
    ctx->haveDustSpecies = true;
    return MCB200_OK;
}

int mcb200_set_opacity(mcb200_ctx *ctx, int32_t iG, const float *opacity, const float *scaOpac)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    if (!opacity) return fail(ctx, MCB200_EINVAL, "null opacity");
    size_t ts = tsize(ctx, *g);
    bool re = (g->opacity.n != ts) || (scaOpac && g->scaOpac.n != ts);
    CU(g->opacity.upload(opacity, ts, ctx->stream));
    if (scaOpac) CU(g->scaOpac.upload(scaOpac, ts, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    g->haveOpacity = true;
    if (re) ctx->gridsDirty = true;
    return MCB200_OK;
}

int mcb200_set_pdfs(mcb200_ctx *ctx, int32_t iG, const float *recPDF, const float *dustPDF,
                    const float *totalLines, const float *linePDF)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    const mcb200_config &c = ctx->cfg;
    const float *src = c.lgGas ? recPDF : dustPDF;
    if (!src) return fail(ctx, MCB200_EINVAL, "%s required", c.lgGas ? "recPDF" : "dustPDF");
    if (c.lgGas && !totalLines) return fail(ctx, MCB200_EINVAL, "totalLines required with gas");
    if (c.lgDebug && c.lgGas && !linePDF) return fail(ctx, MCB200_EINVAL, "linePDF required in debug mode");
    size_t ts = tsize(ctx, *g);
    int nRows = g->nCells + 1;
    bool re = (g->pdfT.n != ts);
    if (ctx->asyncPdfs) {
        // Enqueue upload + transpose + check on the copy stream and return: the tables are
        // first needed by the re-emissions of wave 1, so the (PCIe-bound) upload overlaps the
        // stellar wave.  The caller's buffer must stay valid (and should be pinned) until the
        // next transport call returns; the monotonicity verdict is reported by that call.
        cudaStream_t cs = ctx->copyStream;
        // small (possibly pageable -> synchronous) tables first, on the main stream
        if (c.lgGas) {
            if (g->totalLines.n != (size_t)nRows) re = true;
            CU(g->totalLines.upload(totalLines, nRows, ctx->stream));
        }
        if (c.lgDebug && c.lgGas) {
            if (g->linePDF.n != lsize(ctx, *g)) re = true;
            CU(g->linePDF.upload(linePDF, lsize(ctx, *g), ctx->stream));
        }
        CU(cudaStreamSynchronize(ctx->stream));          // earlier users of pdfT are done
        CU(ctx->pdfBad.alloc(c.nGrids));
        if (!ctx->pdfPending) CU(ctx->pdfBad.zero(cs));
        CU(g->pdfT.alloc(ts));
        { int rcs = upload_table_slabs(ctx, *g, src, cs); if (rcs) return rcs; }
        CU(launch_transpose_pdf(g->stage.p, g->pdfT.p, nRows, c.nbins, cs));
        CU(launch_check_monotone(g->pdfT.p, nRows, c.nbins, ctx->pdfBad.p + (iG - 1), cs));
        CU(cudaEventRecord(ctx->pdfReady, cs));
        ctx->pdfPending = true;
        g->havePdf = true;
        if (re) ctx->gridsDirty = true;
        return MCB200_OK;
    }
    { int rcs = upload_table_slabs(ctx, *g, src, ctx->stream); if (rcs) return rcs; }
    CU(g->pdfT.alloc(ts));
    CU(launch_transpose_pdf(g->stage.p, g->pdfT.p, nRows, c.nbins, ctx->stream));
    CU(ctx->flag.alloc(1)); CU(ctx->flag.zero(ctx->stream));
    CU(launch_check_monotone(g->pdfT.p, nRows, c.nbins, ctx->flag.p, ctx->stream));
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    g->stage.release();
    if (bad) { g->havePdf = false; return fail(ctx, MCB200_ETABLE, "grid %d: re-emission PDF is not non-decreasing along nu", iG); }
    if (c.lgGas) {
        if (g->totalLines.n != (size_t)nRows) re = true;
        CU(g->totalLines.upload(totalLines, nRows, ctx->stream));
    }
    if (c.lgDebug && c.lgGas) {
        if (g->linePDF.n != lsize(ctx, *g)) re = true;
        CU(g->linePDF.upload(linePDF, lsize(ctx, *g), ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    g->havePdf = true;
    if (re) ctx->gridsDirty = true;
    return MCB200_OK;
}

int mcb200_set_dust_state(mcb200_ctx *ctx, int32_t iG, const float *Tdust, const int32_t *dustAbunIndex)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    if (!ctx->haveDustSpecies) return fail(ctx, MCB200_ESTATE, "set_dust_species first");
    const mcb200_config &c = ctx->cfg;
    if (!Tdust) return fail(ctx, MCB200_EINVAL, "null Tdust");
    if (c.lgMultiDustChemistry && !dustAbunIndex) return fail(ctx, MCB200_EINVAL, "dustAbunIndex required with multiChemistry");
    // sublimation test of photon_mod.f90:1722-1748 evaluated once per cell
    std::vector<unsigned char> can((size_t)g->nCells + 1, 0);
    size_t s0 = (size_t)c.nSpeciesMax + 1, s1 = (size_t)c.nSizes + 1;
    for (int cell = 1; cell <= g->nCells; ++cell) {
        int comp = c.lgMultiDustChemistry ? dustAbunIndex[cell] : 1;
        if (comp < 1 || comp > c.nDustComp) continue;
        int nSp = ctx->nSpeciesPart[comp - 1];
        for (int nS = 1; nS <= nSp; ++nS) {
            float ab = ctx->grainAbun[(size_t)(comp - 1) + (size_t)c.nDustComp * (nS - 1)];
            float Td = Tdust[(size_t)nS + s0 * (0 + s1 * (size_t)cell)];
            if (ab > 0.f && Td < ctx->TdustSublime[ctx->dustComPoint[comp - 1] - 1 + nS - 1]) { can[cell] = 1; break; }
        }
    }
    bool re = (g->canScatter.n != can.size());
    CU(g->canScatter.upload(can.data(), can.size(), ctx->stream));
    if (ctx->haveDustTables) {
        // the dust closure updates this state in place on the device
        CU(g->Tdust.upload(Tdust, s0 * s1 * ((size_t)g->nCells + 1), ctx->stream));
        if (c.lgMultiDustChemistry) CU(g->dustAbun.upload(dustAbunIndex, (size_t)g->nCells + 1, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (re) ctx->gridsDirty = true;
    return MCB200_OK;
}

int mcb200_set_dust_tables(mcb200_ctx *ctx, const float *widFlx, const float *grainWeight,
                           const int32_t *dustAbsXsecP, int32_t nSpeciesTot, const float *dustEmIntegral,
                           int32_t nTemps)
{
    NEED_CTX();
    if (!ctx->haveCfg || !ctx->haveSpectra) return fail(ctx, MCB200_ESTATE, "set_config and set_spectra first");
    if (!ctx->haveDustSpecies) return fail(ctx, MCB200_ESTATE, "set_dust_species first");
    if (!ctx->xSec.p) return fail(ctx, MCB200_ESTATE, "mcb200_set_xsec first");
    const mcb200_config &c = ctx->cfg;
    if (!widFlx || !grainWeight || !dustAbsXsecP || !dustEmIntegral || nSpeciesTot < 1 || nTemps < 2)
        return fail(ctx, MCB200_EINVAL, "bad dust tables");
    if (nSpeciesTot < c.nSpeciesMax) return fail(ctx, MCB200_EINVAL, "nSpecies %d < nSpeciesMax %d", nSpeciesTot, c.nSpeciesMax);
    if ((size_t)nSpeciesTot != ctx->TdustSublime.size())
        return fail(ctx, MCB200_EINVAL, "nSpecies %d differs from set_dust_species (%d)", nSpeciesTot, (int)ctx->TdustSublime.size());
    size_t nP = (size_t)nSpeciesTot * c.nSizes;
    for (size_t k = 0; k < nP; ++k)
        if (dustAbsXsecP[k] < 1 || (int64_t)dustAbsXsecP[k] + c.nbins - 1 > (int64_t)ctx->xSec.n)
            return fail(ctx, MCB200_EINVAL, "dustAbsXsecP entry %d out of xSecArray", (int)k + 1);
    // dustEmIntegral(nSpecies, nSizes, nTemps) -> contiguous temperature rows; must ascend for locate
    std::vector<float> emT(nP * (size_t)nTemps);
    for (size_t k = 0; k < nP; ++k)
        for (int t = 0; t < nTemps; ++t) {
            float v = dustEmIntegral[k + nP * (size_t)t];
            emT[k * (size_t)nTemps + t] = v;
            if (t > 0 && !(v >= emT[k * (size_t)nTemps + t - 1]))
                return fail(ctx, MCB200_ETABLE, "dustEmIntegral(%d,%d,:) is not non-decreasing", (int)(k % nSpeciesTot) + 1, (int)(k / nSpeciesTot) + 1);
        }
    CU(ctx->widFlx.upload(widFlx, c.nbins, ctx->stream));
    CU(ctx->grainWeight.upload(grainWeight, c.nSizes, ctx->stream));
    CU(ctx->absP.upload(dustAbsXsecP, nP, ctx->stream));
    CU(ctx->emT.upload(emT.data(), emT.size(), ctx->stream));
    CU(ctx->dGrainAbun.upload(ctx->grainAbun.data(), ctx->grainAbun.size(), ctx->stream));
    CU(ctx->dSublime.upload(ctx->TdustSublime.data(), ctx->TdustSublime.size(), ctx->stream));
    CU(ctx->dSpeciesPart.upload(ctx->nSpeciesPart.data(), ctx->nSpeciesPart.size(), ctx->stream));
    CU(ctx->dComPoint.upload(ctx->dustComPoint.data(), ctx->dustComPoint.size(), ctx->stream));
    CU(ctx->nConv.alloc(1));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->nSpeciesTot = nSpeciesTot;
    ctx->nTemps = nTemps;
    ctx->haveDustTables = true;
    return MCB200_OK;
}

namespace {
int dust_args(mcb200_ctx *ctx, GridState *g, DustArgs &A)
{
    const mcb200_config &c = ctx->cfg;
    if (!ctx->haveDustTables) return fail(ctx, MCB200_ESTATE, "mcb200_set_dust_tables first");
    if (c.lgGas || !c.lgDust) return fail(ctx, MCB200_EUNSUPPORTED, "the dust closure is the dust-only branch (lgDust and not lgGas)");
    if (!g->Tdust.p) return fail(ctx, MCB200_ESTATE, "mcb200_set_dust_state (after set_dust_tables) first");
    A = DustArgs{};
    A.nCells = g->nCells; A.nb = c.nbins; A.nSpeciesMax = c.nSpeciesMax; A.nSizes = c.nSizes;
    A.nDustComp = c.nDustComp; A.nSpeciesTot = ctx->nSpeciesTot; A.nTemps = ctx->nTemps;
    A.multiChem = c.lgMultiDustChemistry; A.lgDebug = c.lgDebug; A.sym = c.lgSymmetricXYZ;
    A.nuArray = ctx->nuArray.p; A.widFlx = ctx->widFlx.p; A.xSec = ctx->xSec.p; A.absP = ctx->absP.p;
    A.nSpeciesPart = ctx->dSpeciesPart.p; A.dustComPoint = ctx->dComPoint.p; A.dustAbunIndex = g->dustAbun.p;
    A.grainAbun = ctx->dGrainAbun.p; A.grainWeight = ctx->grainWeight.p; A.TdustSublime = ctx->dSublime.p;
    A.emT = ctx->emT.p; A.Tdust = g->Tdust.p; A.canScatter = g->canScatter.p; A.nConv = ctx->nConv.p;
    return MCB200_OK;
}
}  // namespace

int mcb200_dust_update(mcb200_ctx *ctx, int32_t iG, float XHILimit, float *Tdust, int32_t *lgConverged,
                       int64_t *nConverged)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    DustArgs A;
    int rc = dust_args(ctx, g, A);
    if (rc) return rc;
    if (ctx->pending) return fail(ctx, MCB200_ESTATE, "tallies pending: call mcb200_reduce first");
    rc = ensure_estimators(ctx, *g);
    if (rc) return rc;
    // iterateMC zeroes grid%lgConverged at the start of every iteration (iteration_mod.f90:87);
    // a cell no packet crossed is skipped by updateCell (update_mod.f90:104-149) and stays 0
    CU(g->lgConverged.alloc((size_t)g->nCells + 1));
    CU(g->lgConverged.zero(ctx->stream));
    CU(ctx->nConv.zero(ctx->stream));
    A.Jste = g->Jste.p; A.Jdif = g->Jdif.p; A.lgConverged = g->lgConverged.p; A.XHILimit = XHILimit;
    if (A.lgDebug && !A.Jdif) return fail(ctx, MCB200_ESTATE, "Jdif missing in debug mode");
    CU(launch_dust_update(A, ctx->stream));
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, ctx->nConv.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (Tdust) CU(cudaMemcpyAsync(Tdust, g->Tdust.p, g->Tdust.n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (lgConverged) CU(cudaMemcpyAsync(lgConverged, g->lgConverged.p, g->lgConverged.n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (nConverged) *nConverged = (int64_t)n;
    return MCB200_OK;
}

int mcb200_dust_pdf(mcb200_ctx *ctx, int32_t iG, float *dustPDF)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    DustArgs A;
    int rc = dust_args(ctx, g, A);
    if (rc) return rc;
    size_t ts = tsize(ctx, *g);
    bool re = (g->pdfT.n != ts);
    if (ctx->pdfPending) { CU(cudaStreamSynchronize(ctx->copyStream)); ctx->pdfPending = false; }
    CU(g->pdfT.alloc(ts));
    A.pdfT = g->pdfT.p;
    CU(launch_dust_pdf(A, ctx->numSMs, ctx->stream));
    if (dustPDF) {
        // back to the reference layout (0:nCells, nbins), cell index fastest
        CU(g->stage.alloc(ts));
        CU(launch_transpose_pdf(g->pdfT.p, g->stage.p, ctx->cfg.nbins, g->nCells + 1, ctx->stream));
        CU(cudaMemcpyAsync(dustPDF, g->stage.p, ts * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        g->stage.release();
    }
    CU(cudaStreamSynchronize(ctx->stream));
    g->havePdf = true;
    if (re) ctx->gridsDirty = true;
    return MCB200_OK;
}

int mcb200_zero_estimators(mcb200_ctx *ctx)
{
    NEED_CTX();
    if (!ctx->haveCfg) return fail(ctx, MCB200_ESTATE, "set_config first");
    if (ctx->sideStream) CU(cudaStreamSynchronize(ctx->sideStream));   // pushes / clears of an exchange that was never folded
    {
        int rc = ensure_sed(ctx);
        if (rc) return rc;
        CU(ctx->sedQ.zero(ctx->stream));
        std::fill(ctx->sed.begin(), ctx->sed.end(), 0.f);
        std::fill(ctx->sedCount.begin(), ctx->sedCount.end(), 0ll);
        ctx->sedReady = false;
    }
    for (auto &g : ctx->grids) {
        if (!g.set) continue;
        int rc = ensure_estimators(ctx, g);
        if (rc) return rc;
        CU(g.JsteQ.zero(ctx->stream)); CU(g.Jste.zero(ctx->stream));
        CU(g.escQ.zero(ctx->stream));  CU(g.esc.zero(ctx->stream));
        CU(g.JdifQ.zero(ctx->stream)); CU(g.Jdif.zero(ctx->stream));
        CU(g.lineQ.zero(ctx->stream)); CU(g.linePk.zero(ctx->stream));
        CU(g.nuTouched.zero(ctx->stream));
        CU(g.JsteQ2.zero(ctx->stream)); CU(g.escQ2.zero(ctx->stream)); CU(g.nuTouched2.zero(ctx->stream));
        g.folded.clear();
        g.jShards.clear();
    }
    ctx->pending2 = false;
    CU(ctx->planeDist.zero(ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->pending = false;
    ctx->exchanged = false;
    return MCB200_OK;
}

int mcb200_transport(mcb200_ctx *ctx, int32_t iStar, int64_t nPacketsGlobal, float deltaE, mcb200_counters *counters)
{
    NEED_CTX();
    if (iStar < 1) return fail(ctx, MCB200_EINVAL, "iStar must be >= 1 (use mcb200_transport_diffuse for iStar=0)");
    return run_transport(ctx, iStar, 0, nullptr, nPacketsGlobal, deltaE, counters);
}

int mcb200_transport_diffuse(mcb200_ctx *ctx, int32_t gpLoc, const int32_t *cellLoc, int64_t nPacketsGlobal,
                             float deltaE, mcb200_counters *counters)
{
    NEED_CTX();
    return run_transport(ctx, 0, gpLoc, cellLoc, nPacketsGlobal, deltaE, counters);
}

int mcb200_set_res_line_packets(mcb200_ctx *ctx, int32_t iG, const int32_t *resLinePackets)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    if (!resLinePackets) { g->resLinePackets.clear(); return MCB200_OK; }
    g->resLinePackets.assign(resLinePackets, resLinePackets + g->nCells + 1);
    for (int v : g->resLinePackets) if (v < 0) return fail(ctx, MCB200_EINVAL, "negative resLinePackets");
    return MCB200_OK;
}

int mcb200_transport_reslines(mcb200_ctx *ctx, int32_t iStar, float deltaE, mcb200_counters *counters)
{
    NEED_CTX();
    if (iStar < 1) return fail(ctx, MCB200_EINVAL, "iStar must be >= 1");
    return run_transport(ctx, iStar, 0, nullptr, 0, deltaE, counters, true);
}

int mcb200_tally_buffer(mcb200_ctx *ctx, int32_t iG, int32_t which, void **devPtr, int64_t *count)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set || !devPtr || !count) return fail(ctx, MCB200_EINVAL, "bad tally_buffer arguments");
    int rc = ensure_estimators(ctx, *g);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    if (which >= 16) {                   // second tally set
        rc = ensure_second_set(ctx, *g);
        if (rc) return rc;
        CU(cudaStreamSynchronize(ctx->stream));
        if (which == 16) { *devPtr = g->JsteQ2.p; *count = (int64_t)g->JsteQ2.n; }
        else if (which == 17) { *devPtr = g->escQ2.p; *count = (int64_t)g->escQ2.n; }
        else if (which == 20) { *devPtr = g->nuTouched2.p; *count = (int64_t)g->nuTouched2.n; }
        else return fail(ctx, MCB200_EINVAL, "bad tally selector %d", which);
        return MCB200_OK;
    }
    if (which == 0 || which == 2) {
        DevBuf<unsigned long long> *b = which == 0 ? &g->JsteQ : &g->JdifQ;
        *devPtr = b->p; *count = (int64_t)b->n;
    } else if (which == 1 || which == 3) {
        DevBuf<unsigned int> *b = which == 1 ? &g->escQ : &g->lineQ;
        *devPtr = b->p; *count = (int64_t)b->n;
    } else if (which == 4) {
        *devPtr = g->nuTouched.p; *count = (int64_t)g->nuTouched.n;
    } else if (which == 5) {
        *devPtr = ctx->planeDist.p; *count = (int64_t)ctx->planeDist.n;
    } else if (which == 6) {
        rc = ensure_sed(ctx);
        if (rc) return rc;
        CU(cudaStreamSynchronize(ctx->stream));
        *devPtr = ctx->sedQ.p; *count = (int64_t)ctx->sedQ.n;
    } else {
        return fail(ctx, MCB200_EINVAL, "bad tally selector %d", which);
    }
    return MCB200_OK;
}

int mcb200_reduce(mcb200_ctx *ctx)
{
    NEED_CTX();
    auto t0 = std::chrono::steady_clock::now();
    trace_mark(ctx, "reduce: start");
    int rc = fold_pending(ctx);
    ctx->lastPhaseMs[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    trace_mark(ctx, "reduce: done");
    if (ctx->traceOn == 1 && ctx->rank == 0 && !ctx->xtrace.empty()) {
        fprintf(stderr, "[mcb200 trace]");
        for (auto &m : ctx->xtrace) fprintf(stderr, " | %s %.3f", m.first, m.second);
        fprintf(stderr, " | device: pushes %.3f ms, merge %.3f ms\n", ctx->lastPhaseMs[3], ctx->lastPhaseMs[2]);
        ctx->xtrace.clear();
    }
    return rc;
}

int mcb200_reduce_range(mcb200_ctx *ctx, int32_t iG, int32_t nu0, int32_t nu1)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    const int nb = ctx->cfg.nbins;
    if (nu0 < 0 || nu1 > nb || nu1 < nu0) return fail(ctx, MCB200_EINVAL, "bad plane range %d..%d", nu0, nu1);
    if (!ctx->pending) return MCB200_OK;
    if (ctx->pending2) return fail(ctx, MCB200_ESTATE, "second tally set pending: use mcb200_reduce");
    int rc = ensure_sed(ctx);
    if (rc) return rc;
    int launches = 0;
    if (!ctx->sedReady)                  // the SED counts of these planes, before the fold clears them
        for (int ang = 0; ang <= ctx->cfg.nAngleBins; ++ang) {
            CU(launch_sed_sum(g->escQ.p, (size_t)g->nCells + 1, nu0 + (nb + 1) * ang, nu1 - nu0 + 1, ctx->sedQ.p, ctx->stream));
            ++launches;
        }
    rc = fold_planes(ctx, *g, nu0, nu1, &launches);
    if (rc) return rc;
    if (g->folded.empty()) g->folded.assign(nb + 1, 0);
    for (int nu = nu0; nu <= nu1; ++nu) g->folded[nu] = 1;
    ctx->rangeFoldLaunches += launches;
    return MCB200_OK;
}

int mcb200_comm_unique_id(mcb200_ctx *ctx, void *id128)
{
    NEED_CTX();
    if (!id128) return fail(ctx, MCB200_EINVAL, "null id buffer");
    NcclApi &N = nccl_api();
    if (!N.handle) return fail(ctx, MCB200_ECOMM, "NCCL unavailable: %s", N.why.c_str());
    NcclId id;
    NC(N.GetUniqueId(&id));
    std::memcpy(id128, id.internal, sizeof(id.internal));
    return MCB200_OK;
}

int mcb200_comm_init(mcb200_ctx *ctx, const void *id128)
{
    NEED_CTX();
    if (!id128) return fail(ctx, MCB200_EINVAL, "null id buffer");
    if (ctx->comm) return fail(ctx, MCB200_ESTATE, "communicator already initialised");
    NcclApi &N = nccl_api();
    if (!N.handle) return fail(ctx, MCB200_ECOMM, "NCCL unavailable: %s", N.why.c_str());
    NcclId id;
    std::memcpy(id.internal, id128, sizeof(id.internal));
    NC(N.CommInitRank(&ctx->comm, ctx->nranks, id, ctx->rank));
    return MCB200_OK;
}

int mcb200_comm_destroy(mcb200_ctx *ctx)
{
    NEED_CTX();
    if (!ctx->comm) return MCB200_OK;
    bool mapped = false;
    for (auto &g : ctx->grids) mapped = mapped || !g.peerQ.empty();
    if (mapped) {
        // nobody unmaps (or, afterwards, frees) a buffer a peer may still be using: every rank is
        // past its last peer access before the mappings go, and every mapping is gone before any
        // rank can go on to free its buffers
        int rc = comm_barrier(ctx);
        if (rc) return rc;
        CU(cudaStreamSynchronize(ctx->stream));
        for (auto &g : ctx->grids) p2p_close(g, ctx->rank);
        rc = comm_barrier(ctx);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    NC(nccl_api().CommDestroy(ctx->comm));
    ctx->comm = nullptr;
    ctx->p2pState = 0;
    return MCB200_OK;
}

int mcb200_exchange_path(mcb200_ctx *ctx, int32_t *path, char *why, int64_t whyLen, double *phaseMs)
{
    NEED_CTX();
    if (path) *path = ctx->lastExchangePath;
    if (phaseMs) for (int i = 0; i < 4; ++i) phaseMs[i] = ctx->lastPhaseMs[i];
    if (why && whyLen > 0) {
        std::strncpy(why, ctx->lastP2PWhy.c_str(), (size_t)whyLen - 1);
        why[whyLen - 1] = 0;
    }
    return MCB200_OK;
}

int mcb200_exchange(mcb200_ctx *ctx)
{
    NEED_CTX();
    if (!ctx->haveCfg) return fail(ctx, MCB200_ESTATE, "set_config first");
    if (ctx->solo) return fail(ctx, MCB200_ESTATE, "option solo is set: this context runs as a single rank");
    if (!ctx->pending || (ctx->nranks == 1 && !ctx->comm)) return MCB200_OK;
    if (!ctx->comm) return fail(ctx, MCB200_ESTATE, "no communicator: call mcb200_comm_init (or all-reduce the buffers of mcb200_tally_buffer yourself)");
    if (ctx->exchanged) return fail(ctx, MCB200_ESTATE, "pending tallies already exchanged: call mcb200_reduce");
    auto t0 = std::chrono::steady_clock::now();
    int rc = comm_exchange(ctx);
    ctx->lastPhaseMs[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc == MCB200_OK) ctx->exchanged = true;
    return rc;
}

int mcb200_nccl_info(int32_t *version, char *path, int64_t pathLen)
{
    NcclApi &N = nccl_api();
    if (version) *version = 0;
    if (path && pathLen > 0) path[0] = 0;
    if (!N.handle) return MCB200_ECOMM;
    int v = 0;
    if (N.GetVersion(&v) != 0) return MCB200_ECOMM;
    if (version) *version = v;
    Dl_info di;
    if (path && pathLen > 0 && dladdr(reinterpret_cast<void *>(N.GetVersion), &di) && di.dli_fname) {
        std::strncpy(path, di.dli_fname, (size_t)pathLen - 1);
        path[pathLen - 1] = 0;
    }
    return MCB200_OK;
}

int mcb200_exchange_info(mcb200_ctx *ctx, int64_t *bytesSent, int32_t *sparseGrids, int32_t *ncclVersion)
{
    NEED_CTX();
    if (bytesSent) *bytesSent = ctx->lastExchangeBytes;
    if (sparseGrids) *sparseGrids = ctx->lastExchangeSparse;
    if (ncclVersion) {
        *ncclVersion = 0;
        NcclApi &N = nccl_api();
        if (N.handle && N.GetVersion) { int v = 0; if (N.GetVersion(&v) == 0) *ncclVersion = v; }
    }
    return MCB200_OK;
}

int mcb200_escaped_compact(mcb200_ctx *ctx, int32_t iG, int32_t set, void **devList, int64_t *nEntries)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set || !devList || !nEntries) return fail(ctx, MCB200_EINVAL, "bad escaped_compact arguments");
    if (!ctx->pending) return fail(ctx, MCB200_ESTATE, "no pending tallies");
    int rc = esc_compact(ctx, *g, set, nEntries);
    if (rc) return rc;
    *devList = ctx->escList.p;
    return MCB200_OK;
}

int mcb200_escaped_scatter(mcb200_ctx *ctx, int32_t iG, int32_t set, const void *devList, int64_t nEntries)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set || (nEntries > 0 && !devList) || nEntries < 0) return fail(ctx, MCB200_EINVAL, "bad escaped_scatter arguments");
    unsigned int *Q = set == 1 ? g->escQ2.p : g->escQ.p;
    if (!Q) return fail(ctx, MCB200_ESTATE, "tally set %d not allocated", set);
    size_t total = set == 1 ? g->escQ2.n : g->escQ.n;
    CU(launch_esc_scatter(Q, total, reinterpret_cast<const unsigned long long *>(devList), (unsigned long long)nEntries,
                          ctx->numSMs * 8, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MCB200_OK;
}

int mcb200_fetch_sed(mcb200_ctx *ctx, float *SED, int64_t *counts)
{
    NEED_CTX();
    if (!ctx->haveCfg) return fail(ctx, MCB200_ESTATE, "set_config first");
    if (ctx->pending) return fail(ctx, MCB200_ESTATE, "tallies pending: call mcb200_reduce first");
    int rc = ensure_sed(ctx);
    if (rc) return rc;
    // internal rows are nu = 0..nbins (row 0 = packets that escaped before getting a frequency
    // bin is never written); the reference's SED(1:nbins, 0:nAngleBins) drops row 0
    const int nb = ctx->cfg.nbins, nA = ctx->cfg.nAngleBins;
    for (int a = 0; a <= nA; ++a)
        for (int f = 1; f <= nb; ++f) {
            size_t src = (size_t)f + (size_t)(nb + 1) * a, dst = (size_t)(f - 1) + (size_t)nb * a;
            if (SED) SED[dst] = ctx->sed[src];
            if (counts) counts[dst] = ctx->sedCount[src];
        }
    return MCB200_OK;
}

int mcb200_fetch_contcube(mcb200_ctx *ctx, int32_t iG, float *contI)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set || !contI) return fail(ctx, MCB200_EINVAL, "bad fetch_contcube arguments");
    if (ctx->pending) return fail(ctx, MCB200_ESTATE, "tallies pending: call mcb200_reduce first");
    int rc = ensure_estimators(ctx, *g);
    if (rc) return rc;
    const size_t nR = (size_t)g->nCells + 1;
    const int nA = ctx->cfg.nAngleBins + 1;
    CU(g->contI.alloc(nR * (size_t)nA));
    CU(launch_contcube(g->esc.p, nR, ctx->cfg.nbins, nA, g->contI.p, ctx->stream));
    CU(cudaMemcpyAsync(contI, g->contI.p, nR * (size_t)nA * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MCB200_OK;
}

int mcb200_fetch_estimators_cells(mcb200_ctx *ctx, int32_t iG, int32_t firstCell, int32_t cellStride, float *Jste,
                                   float *Jdif, int64_t *nCellsOut)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    if (ctx->pending) return fail(ctx, MCB200_ESTATE, "tallies pending: call mcb200_reduce first");
    if (firstCell < 1 || cellStride < 1) return fail(ctx, MCB200_EINVAL, "firstCell >= 1 and cellStride >= 1 required");
    int rc = ensure_estimators(ctx, *g);
    if (rc) return rc;
    const int nb = ctx->cfg.nbins;
    const size_t nR = (size_t)g->nCells + 1;
    const int64_t nMine = firstCell > g->nCells ? 0 : ((int64_t)g->nCells - firstCell) / cellStride + 1;
    if (nCellsOut) *nCellsOut = nMine;
    if (nMine == 0) return MCB200_OK;
    if (Jdif && !g->Jdif.p) return fail(ctx, MCB200_ESTATE, "Jdif only exists in debug mode");
    CU(g->cellStage.alloc((size_t)nMine * (size_t)nb));      // its own buffer: no reallocation against the PDF staging
    for (int which = 0; which < 2; ++which) {
        float *dst = which ? Jdif : Jste;
        if (!dst) continue;
        CU(launch_gather_cells(which ? g->Jdif.p : g->Jste.p, nR, nb, firstCell, cellStride, (int)nMine, g->cellStage.p, ctx->stream));
        CU(cudaMemcpyAsync(dst, g->cellStage.p, (size_t)nMine * (size_t)nb * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return MCB200_OK;
}

// position-sensitive 64-bit checksum of a float32 array (bit patterns): sum of bits(i) * (2*i + 1)
// mod 2^64 -- equal arrays give equal sums, and a changed, moved or missing element changes it
__global__ void __launch_bounds__(256) checksum_kernel(const unsigned int *__restrict__ x, size_t n, unsigned long long *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long s = 0;
    for (; i < n; i += stride) s += (unsigned long long)x[i] * (2ull * (unsigned long long)i + 1ull);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

int mcb200_checksum(mcb200_ctx *ctx, int32_t iG, int32_t which, uint64_t *sum)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set || !sum) return fail(ctx, MCB200_EINVAL, "bad checksum arguments");
    if (ctx->pending) return fail(ctx, MCB200_ESTATE, "tallies pending: call mcb200_reduce first");
    int rc = ensure_estimators(ctx, *g);
    if (rc) return rc;
    DevBuf<float> *b = which == 0 ? &g->Jste : which == 1 ? &g->esc : which == 2 ? &g->Jdif : which == 3 ? &g->linePk : nullptr;
    if (!b) return fail(ctx, MCB200_EINVAL, "bad checksum selector %d", which);
    CU(ctx->nConv.alloc(1)); CU(ctx->nConv.zero(ctx->stream));
    if (b->n) checksum_kernel<<<ctx->numSMs * 8, 256, 0, ctx->stream>>>(reinterpret_cast<const unsigned int *>(b->p), b->n, ctx->nConv.p);
    CU(cudaGetLastError());
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, ctx->nConv.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *sum = (uint64_t)v;
    return MCB200_OK;
}

int mcb200_fetch_estimators(mcb200_ctx *ctx, int32_t iG, float *Jste, float *escapedPackets, float *Jdif, float *linePackets)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    if (ctx->pending) return fail(ctx, MCB200_ESTATE, "tallies pending: call mcb200_reduce first");
    int rc = ensure_estimators(ctx, *g);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    if (Jste) CU(cudaMemcpy(Jste, g->Jste.p, g->Jste.n * sizeof(float), cudaMemcpyDeviceToHost));
    if (escapedPackets) CU(cudaMemcpy(escapedPackets, g->esc.p, g->esc.n * sizeof(float), cudaMemcpyDeviceToHost));
    if (Jdif) {
        if (!g->Jdif.p) return fail(ctx, MCB200_ESTATE, "Jdif only exists in debug mode");
        CU(cudaMemcpy(Jdif, g->Jdif.p, g->Jdif.n * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (linePackets) {
        if (!g->linePk.p) return fail(ctx, MCB200_ESTATE, "linePackets only exists in debug mode");
        CU(cudaMemcpy(linePackets, g->linePk.p, g->linePk.n * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return MCB200_OK;
}

// Non-zero entries of a float array as (index << 32 | value bits), tile by tile: a block takes a
// tile of 4096 consecutive elements (16 per thread, so the tile's entries come out in index
// order), counts its non-zeros with a block scan and reserves its slice of the list with one
// atomic.  Tiles land in the list in any order, entries inside a tile ascend: the host writes
// that follow touch one 16 KB window at a time.  HBM-bound: the array is read once.
constexpr int kSparsePerThread = 16;
__global__ void __launch_bounds__(256) sparse_f32_kernel(const float *__restrict__ E, unsigned long long len,
                                                         unsigned long long *__restrict__ list,
                                                         unsigned long long *__restrict__ count, unsigned long long capacity)
{
    __shared__ unsigned int warpTot[8];
    __shared__ unsigned long long tileBase;
    const unsigned int lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const unsigned long long tile = 256ull * kSparsePerThread;
    for (unsigned long long t0 = (unsigned long long)blockIdx.x * tile; t0 < len; t0 += (unsigned long long)gridDim.x * tile) {
        const unsigned long long i0 = t0 + (unsigned long long)threadIdx.x * kSparsePerThread;
        unsigned int bits[kSparsePerThread];
        unsigned int nz = 0;
        if (i0 + kSparsePerThread <= len) {
            const uint4 *p = reinterpret_cast<const uint4 *>(E + i0);      // i0 is a multiple of 16: aligned
#pragma unroll
            for (int k = 0; k < kSparsePerThread / 4; ++k) {
                uint4 v = p[k];
                bits[4 * k] = v.x; bits[4 * k + 1] = v.y; bits[4 * k + 2] = v.z; bits[4 * k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kSparsePerThread; ++k) bits[k] = i0 + k < len ? __float_as_uint(E[i0 + k]) : 0u;
        }
#pragma unroll
        for (int k = 0; k < kSparsePerThread; ++k) nz += bits[k] != 0u;
        // exclusive scan of nz over the block
        unsigned int incl = nz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (unsigned int)d) incl += t;
        }
        if (lane == 31) warpTot[w] = incl;
        __syncthreads();
        unsigned int before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { unsigned int t = warpTot[k]; if ((unsigned int)k < w) before += t; total += t; }
        if (threadIdx.x == 0) tileBase = total ? atomicAdd(count, (unsigned long long)total) : 0ull;
        __syncthreads();
        if (nz) {
            unsigned long long pos = tileBase + before + (incl - nz);
#pragma unroll
            for (int k = 0; k < kSparsePerThread; ++k) {
                if (bits[k] != 0u) {
                    if (pos < capacity) list[pos] = ((i0 + k) << 32) | (unsigned long long)bits[k];
                    ++pos;
                }
            }
        }
        __syncthreads();                       // warpTot / tileBase are reused by the next tile
    }
}

namespace {
// Jste (optional): its dense copy is started once the device part of the sparse path is done, so
// that it crosses PCIe while the host threads scatter the escapedPackets entries
int fetch_sparse_impl(mcb200_ctx *ctx, int32_t iG, float *Jste, float *escapedPackets, int32_t clearPrevious,
                      int64_t *nNonZero)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    if (!escapedPackets) return fail(ctx, MCB200_EINVAL, "escapedPackets is NULL");
    if (ctx->pending) return fail(ctx, MCB200_ESTATE, "tallies pending: call mcb200_reduce first");
    int rc = ensure_estimators(ctx, *g);
    if (rc) return rc;
    const size_t len = g->esc.n;
    // The host side is a scatter of ~10^6-10^7 single floats into a multi-GB array: every store
    // misses the cache, so it is spread over a few host threads (disjoint entries, no sharing).
    auto parallel_for = [](size_t n, auto &&body) {
        unsigned int hw = std::thread::hardware_concurrency();
        size_t nt = n < (1u << 16) ? 1 : (hw >= 16 ? 8 : hw >= 4 ? hw / 2 : 1);
        if (nt <= 1) { body((size_t)0, n); return; }
        std::vector<std::thread> th;
        size_t chunk = (n + nt - 1) / nt;
        for (size_t t = 0; t < nt; ++t) {
            size_t a = t * chunk, b = a + chunk < n ? a + chunk : n;
            if (a < b) th.emplace_back([&body, a, b] { body(a, b); });
        }
        for (auto &x : th) x.join();
    };
    const size_t cap = len / 64 + 4096;         // 1.6 % of the entries; beyond that the dense copy is as good
    unsigned long long n = ~0ull;
    if (len < (1ull << 32)) {
        CU(ctx->sparseList.alloc(cap));
        CU(ctx->sparseCount.alloc(1));
        CU(ctx->sparseCount.zero(ctx->stream));
        int dev = 0, sms = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        sparse_f32_kernel<<<sms * 8, 256, 0, ctx->stream>>>(g->esc.p, (unsigned long long)len, ctx->sparseList.p,
                                                            ctx->sparseCount.p, (unsigned long long)cap);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(&n, ctx->sparseCount.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    const bool withJ = Jste != nullptr && ctx->copyStream != nullptr;
    if (Jste && !withJ) CU(cudaMemcpy(Jste, g->Jste.p, g->Jste.n * sizeof(float), cudaMemcpyDeviceToHost));
    if (n > cap) {
        // too dense (or an index would not fit 32 bits): the plain copy; every entry is written
        CU(cudaStreamSynchronize(ctx->stream));
        if (withJ) CU(cudaMemcpy(Jste, g->Jste.p, g->Jste.n * sizeof(float), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(escapedPackets, g->esc.p, len * sizeof(float), cudaMemcpyDeviceToHost));
        g->sparsePrev.clear();
        g->sparsePrevPtr = escapedPackets;
        g->sparsePrevDense = true;
        if (nNonZero) *nNonZero = -1;
        return MCB200_OK;
    }
    if (n > ctx->sparseHostCap) {
        if (ctx->sparseHost) cudaFreeHost(ctx->sparseHost);
        ctx->sparseHost = nullptr;
        ctx->sparseHostCap = 0;
        size_t want = (size_t)n + (size_t)n / 4 + 4096;
        if (want > cap) want = cap;
        CU(cudaHostAlloc((void **)&ctx->sparseHost, want * 8, cudaHostAllocDefault));
        ctx->sparseHostCap = want;
    }
    if (n) CU(cudaMemcpy(ctx->sparseHost, ctx->sparseList.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (withJ) CU(cudaMemcpyAsync(Jste, g->Jste.p, g->Jste.n * sizeof(float), cudaMemcpyDeviceToHost, ctx->copyStream));
    // from here on the host works while Jste crosses PCIe
    if (clearPrevious && g->sparsePrevPtr == escapedPackets) {
        if (g->sparsePrevDense) {
            parallel_for(len, [&](size_t a, size_t b) { memset(escapedPackets + a, 0, (b - a) * sizeof(float)); });
        } else {
            const unsigned int *prev = g->sparsePrev.data();
            parallel_for(g->sparsePrev.size(), [&](size_t a, size_t b) { for (size_t k = a; k < b; ++k) escapedPackets[prev[k]] = 0.f; });
        }
    }
    g->sparsePrev.clear();
    g->sparsePrevPtr = escapedPackets;
    g->sparsePrevDense = false;
    g->sparsePrev.resize((size_t)n);
    {
        const unsigned long long *src = ctx->sparseHost;
        unsigned int *prev = g->sparsePrev.data();
        parallel_for((size_t)n, [&](size_t a, size_t b) {
            for (size_t k = a; k < b; ++k) {
                unsigned long long e = src[k];
                unsigned int idx = (unsigned int)(e >> 32), bits = (unsigned int)e;
                float v;
                memcpy(&v, &bits, 4);
                escapedPackets[idx] = v;
                prev[k] = idx;
            }
        });
    }
    if (withJ) CU(cudaStreamSynchronize(ctx->copyStream));
    if (nNonZero) *nNonZero = (int64_t)n;
    return MCB200_OK;
}
}  // namespace

int mcb200_fetch_escaped_sparse(mcb200_ctx *ctx, int32_t iG, float *escapedPackets, int32_t clearPrevious,
                                int64_t *nNonZero)
{
    return fetch_sparse_impl(ctx, iG, nullptr, escapedPackets, clearPrevious, nNonZero);
}

int mcb200_fetch_estimators_sparse(mcb200_ctx *ctx, int32_t iG, float *Jste, float *escapedPackets,
                                   int32_t clearPrevious, int64_t *nNonZero)
{
    if (ctx && !Jste) return fail(ctx, MCB200_EINVAL, "Jste is NULL");
    return fetch_sparse_impl(ctx, iG, Jste, escapedPackets, clearPrevious, nNonZero);
}

int mcb200_fetch_tallies(mcb200_ctx *ctx, int32_t iG, int64_t *JsteQ, int64_t *escapedQ, int64_t *JdifQ, int64_t *linePacketsQ)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    int rc = ensure_estimators(ctx, *g);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    if (JsteQ) CU(cudaMemcpy(JsteQ, g->JsteQ.p, g->JsteQ.n * 8, cudaMemcpyDeviceToHost));
    auto widen = [&](DevBuf<unsigned int> &b, int64_t *dst) -> cudaError_t {
        std::vector<unsigned int> tmp(b.n);
        cudaError_t e = cudaMemcpy(tmp.data(), b.p, b.n * sizeof(unsigned int), cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < b.n; ++i) dst[i] = (int64_t)tmp[i];
        return e;
    };
    if (escapedQ) CU(widen(g->escQ, escapedQ));
    if (JdifQ && g->JdifQ.p) CU(cudaMemcpy(JdifQ, g->JdifQ.p, g->JdifQ.n * 8, cudaMemcpyDeviceToHost));
    if (linePacketsQ && g->lineQ.p) CU(widen(g->lineQ, linePacketsQ));
    return MCB200_OK;
}

int mcb200_len_unit(mcb200_ctx *ctx, int32_t iG, double *lenUnit)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set || !lenUnit) return fail(ctx, MCB200_EINVAL, "bad len_unit arguments");
    *lenUnit = std::ldexp(1.0, g->lenExp);
    return MCB200_OK;
}

int mcb200_fetch_plane_distribution(mcb200_ctx *ctx, int32_t *planeIonDistribution)
{
    NEED_CTX();
    if (!planeIonDistribution || !ctx->planeDist.p) return fail(ctx, MCB200_ESTATE, "no plane-ionisation run yet");
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(planeIonDistribution, ctx->planeDist.p, ctx->planeDist.n * sizeof(int), cudaMemcpyDeviceToHost));
    return MCB200_OK;
}

int mcb200_fetch_qphot_counts(mcb200_ctx *ctx, int64_t *counts)
{
    NEED_CTX();
    if (!counts || !ctx->qphot.p) return fail(ctx, MCB200_ESTATE, "no transport call yet");
    CU(cudaMemcpy(counts, ctx->qphot.p, ctx->qphot.n * 8, cudaMemcpyDeviceToHost));
    return MCB200_OK;
}

int mcb200_fetch_fates(mcb200_ctx *ctx, int32_t *fates, int64_t nPackets)
{
    NEED_CTX();
    if (!fates || !ctx->fates.p || (size_t)nPackets * 4 > ctx->fates.n) return fail(ctx, MCB200_ESTATE, "no fate trace available");
    CU(cudaMemcpy(fates, ctx->fates.p, (size_t)nPackets * 4 * sizeof(int), cudaMemcpyDeviceToHost));
    return MCB200_OK;
}

int mcb200_set_option(mcb200_ctx *ctx, const char *name, int64_t value)
{
    NEED_CTX();
    if (!name) return MCB200_EINVAL;
    if (!strcmp(name, "trace")) { ctx->trace = value != 0; return MCB200_OK; }
    if (!strcmp(name, "blocks_per_sm")) { ctx->blocksPerSM = (int)value; return MCB200_OK; }
    if (!strcmp(name, "seed")) { ctx->seed = (uint64_t)value; return MCB200_OK; }
    if (!strcmp(name, "order")) { ctx->orderMode = (int)value; return MCB200_OK; }
    if (!strcmp(name, "wavefront")) { ctx->waveMode = (int)value; return MCB200_OK; }
    if (!strcmp(name, "step_budget")) { ctx->stepBudget = (int)value; return MCB200_OK; }
    if (!strcmp(name, "tail")) { ctx->tailThreshold = value; return MCB200_OK; }
    if (!strcmp(name, "wave0_order")) { ctx->wave0Order = value < 0 ? 0 : (value > 2 ? 2 : (int)value); return MCB200_OK; }
    if (!strcmp(name, "wave0_exact")) { ctx->wave0Exact = value != 0; return MCB200_OK; }
    if (!strcmp(name, "wave0_blocks")) { ctx->wave0Blocks = value < 0 ? 0 : (int)value; return MCB200_OK; }   // 0 = full grid
    if (!strcmp(name, "fly_batch")) { ctx->flyBatch = value < 1 ? 1 : (value > 32 ? 32 : (int)value); return MCB200_OK; }
    if (!strcmp(name, "async_pdfs")) { ctx->asyncPdfs = value != 0; return MCB200_OK; }
    if (!strcmp(name, "sed_local")) {
        if (ctx->pending) return fail(ctx, MCB200_ESTATE, "sed_local cannot change while tallies are pending");
        ctx->sedLocal = value != 0;
        return MCB200_OK;
    }
    if (!strcmp(name, "tally_set")) {
        if (value != 0 && value != 1) return fail(ctx, MCB200_EINVAL, "tally_set must be 0 or 1");
        if (ctx->tallySet != (int)value) { ctx->tallySet = (int)value; ctx->gridsDirty = true; }
        return MCB200_OK;
    }
    if (!strcmp(name, "parts")) { ctx->partCount = value < 1 ? 1 : (int)value; ctx->partIndex = 0; return MCB200_OK; }
    if (!strcmp(name, "part")) {
        if (value < 0 || value >= ctx->partCount) return fail(ctx, MCB200_EINVAL, "part out of range");
        ctx->partIndex = (int)value; return MCB200_OK;
    }
    if (!strcmp(name, "exchange_dense")) { ctx->exchangeDense = value != 0; return MCB200_OK; }
    if (!strcmp(name, "exchange_allreduce")) { ctx->exchangeAllReduce = value != 0; return MCB200_OK; }
    if (!strcmp(name, "esc_compact")) { ctx->escCompact = value != 0; return MCB200_OK; }
    if (!strcmp(name, "epoch")) { ctx->epoch = value; return MCB200_OK; }
    if (!strcmp(name, "pdf_slabs")) { ctx->pdfSlabs = value != 0; return MCB200_OK; }
    if (!strcmp(name, "exchange_pack")) { ctx->p2pPack = value != 0; return MCB200_OK; }
    if (!strcmp(name, "exchange_push_blocks")) {
        if (value < 1 || value > 64) return fail(ctx, MCB200_EINVAL, "exchange_push_blocks must be 1..64");
        ctx->p2pPushBlocks = (int)value;
        return MCB200_OK;
    }
    if (!strcmp(name, "exchange_push")) { ctx->p2pPush = (int)value; return MCB200_OK; }
    if (!strcmp(name, "exchange_p2p")) { ctx->p2pMode = (int)value; if (ctx->p2pState < 0) ctx->p2pState = 0; return MCB200_OK; }
    if (!strcmp(name, "solo")) {
        // 1: this rank behaves as rank 0 of 1 (transports every packet of a call itself and folds at
        // once) until the option is cleared -- the N-rank answer checked against the 1-rank answer
        // on the same context (bench.py: nrank_parity)
        if (ctx->pending) return fail(ctx, MCB200_ESTATE, "option solo: tallies pending, call mcb200_reduce first");
        if (value && !ctx->solo) { ctx->soloRank = ctx->rank; ctx->soloNranks = ctx->nranks; ctx->rank = 0; ctx->nranks = 1; ctx->solo = true; }
        else if (!value && ctx->solo) { ctx->rank = ctx->soloRank; ctx->nranks = ctx->soloNranks; ctx->solo = false; }
        return MCB200_OK;
    }
    if (!strcmp(name, "defer_fold")) { ctx->deferFold = value != 0; return MCB200_OK; }
    if (!strcmp(name, "agg_steps")) { ctx->aggSteps = (int)value; return MCB200_OK; }
    if (!strcmp(name, "batch")) { ctx->batch = (int)value; return MCB200_OK; }
    return fail(ctx, MCB200_EINVAL, "unknown option %s", name);
}

int mcb200_test_detmath(mcb200_ctx *ctx, int32_t which, const float *in, float *out, int64_t n)
{
    NEED_CTX();
    DevBuf<float> a, b;
    CU(a.upload(in, (size_t)n, ctx->stream));
    CU(b.alloc((size_t)n));
    detmath_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(which, a.p, b.p, (long long)n);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, b.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MCB200_OK;
}

// Speed of light of the transport's access pattern ("atomic roofline"): every cell
// crossing is one 4-byte read of opacity(cell, nu) and one 64-bit reduction into JsteQ(cell, nu)
// at an address the lanes of a warp do not share.  This kernel issues exactly those two requests
// per iteration at uniformly random addresses inside two windows and nothing else, so its rate
// is what the memory system can give that pattern; bench.py reports the FLY kernel against it.
__global__ void __launch_bounds__(256) access_peak_kernel(unsigned long long *q, const float *tab, unsigned long long nq,
                                                          unsigned long long nt, int iters, int mode, float *sink)
{
    unsigned int s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        unsigned int h = s ^ (s >> 15);
        if (mode & 1) atomicAdd(&q[((unsigned long long)h * nq) >> 32], 1ull);                 // RED.E.ADD.64
        if (mode & 2) acc += __ldg(&tab[((unsigned long long)(h * 2246822519u) * nt) >> 32]);  // LDG.32
    }
    if (acc == 123.456f) *sink = acc;
}

int mcb200_test_access_peak(mcb200_ctx *ctx, int32_t mode, int64_t redWindowBytes, int64_t loadWindowBytes,
                            int64_t opsTotal, double *opsPerSecond)
{
    NEED_CTX();
    if (!(mode >= 1 && mode <= 3) || redWindowBytes < 8 || loadWindowBytes < 4 || opsTotal < 1 || !opsPerSecond)
        return fail(ctx, MCB200_EINVAL, "bad access-peak arguments");
    DevBuf<unsigned long long> q;
    DevBuf<float> t, sink;
    const unsigned long long nq = (unsigned long long)redWindowBytes / 8, nt = (unsigned long long)loadWindowBytes / 4;
    CU(q.alloc((size_t)nq));
    CU(t.alloc((size_t)nt));
    CU(sink.alloc(1));
    CU(q.zero(ctx->stream));
    CU(t.zero(ctx->stream));
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256;                 // 2048 threads per SM: full occupancy
    int iters = (int)((opsTotal + (int64_t)blocks * threads - 1) / ((int64_t)blocks * threads));
    if (iters < 1) iters = 1;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best = 0.f;
    for (int rep = 0; rep < 4; ++rep) {                        // first repetition warms up; best of the rest
        CU(cudaEventRecord(e0, ctx->stream));
        access_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(q.p, t.p, nq, nt, iters, mode, sink.p);
        CU(cudaGetLastError());
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && (best == 0.f || ms < best)) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *opsPerSecond = (double)blocks * threads * iters / ((double)best * 1e-3);
    return MCB200_OK;
}

// measurement hook: the push kernels of the peer-memory merge with every "peer" buffer local (one GPU):
// mode 0 = 64-bit push kernel, 1 = packed push kernel; nElems = elements of the whole exchanged range.
int mcb200_test_push_kernels(mcb200_ctx *ctx, int32_t mode, int32_t nranksSim, int64_t nElems, double *ms)
{
    NEED_CTX();
    if (!ms || nranksSim < 2 || nranksSim > 16 || nElems < 1) return fail(ctx, MCB200_EINVAL, "bad test_push_kernels arguments");
    const int world = nranksSim, rank = 0;
    const size_t count = (size_t)nElems / (size_t)world, slotStride = (count + 1024) / 256 * 256, flagStride = slotStride / (size_t)p2p_pack_block() + 2;
    DevBuf<unsigned long long> Q, recv;
    CU(Q.alloc((size_t)nElems));
    const size_t words = slotStride * (size_t)(world - 1) + (flagStride * (size_t)(world - 1) + 7) / 8 + 1;
    CU(recv.alloc(words * (size_t)(world - 1)));          // one receive buffer per simulated peer
    cudaStream_t s = ctx->stream;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CU(cudaMemsetAsync(Q.p, 0, (size_t)nElems * 8, s));
        CU(cudaMemsetAsync(Q.p, 1, (size_t)nElems * 4, s));     // low words only: the usual case (no block flagged) ...
        CU(cudaMemsetAsync(Q.p, 1, (size_t)nElems / 4, s));     // ... but for the first 1/32 of the range
        CU(cudaEventRecord(ctx->ev0, s));
        if (mode == 1) {
            P2PPackedPeers PP{};
            PP.nranks = world; PP.rank = rank;
            for (int r = 1; r < world; ++r) {
                unsigned long long *base = recv.p + (size_t)(r - 1) * words;
                PP.lo[r] = (unsigned int *)base;
                PP.hi[r] = PP.lo[r] + (size_t)(world - 1) * slotStride;
                PP.flag[r] = (unsigned char *)(base + (size_t)(world - 1) * slotStride);
            }
            CU(launch_p2p_push_packed(PP, Q.p, 0, count, 0, slotStride, flagStride, ctx->numSMs * 16, s));
            CU(cudaMemsetAsync(Q.p + count, 0, (size_t)(world - 1) * count * 8, s));
        } else {
            P2PPush PP{};
            PP.nranks = world; PP.rank = rank;
            for (int r = 1; r < world; ++r) PP.recv[r] = recv.p + (size_t)(r - 1) * words;
            CU(launch_p2p_push(PP, Q.p, 0, count, slotStride, 0, ctx->numSMs * 16, s));
            CU(cudaMemsetAsync(Q.p + count, 0, (size_t)(world - 1) * count * 8, s));
        }
        CU(cudaEventRecord(ctx->ev1, s));
        CU(cudaEventSynchronize(ctx->ev1));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1));
        best = t < best ? t : best;
    }
    *ms = best;
    return MCB200_OK;
}

int mcb200_test_uniforms(mcb200_ctx *ctx, uint64_t seed, uint64_t pid, uint32_t stream, int32_t n, float *out)
{
    NEED_CTX();
    DevBuf<float> b;
    CU(b.alloc((size_t)n));
    uniforms_kernel<<<1, 32, 0, ctx->stream>>>(seed, pid, stream, n, b.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, b.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MCB200_OK;
}


int mcb200_set_xsec(mcb200_ctx *ctx, const float *xSecArray, int64_t nXsec)
{
    NEED_CTX();
    if (!xSecArray || nXsec < 1) return fail(ctx, MCB200_EINVAL, "bad xSecArray");
    CU(ctx->xSec.upload(xSecArray, (size_t)nXsec, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->hXsec.assign(xSecArray, xSecArray + nXsec);
    return MCB200_OK;
}

int mcb200_photo_integrals(mcb200_ctx *ctx, int32_t iG, int32_t nBands, const int32_t *bandOff,
                           const int32_t *bandLow, const int32_t *bandHigh, float *nPhotoSte, float *heatSte,
                           float *nPhotoDif, float *heatDif)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    if (!ctx->xSec.p || !ctx->haveSpectra) return fail(ctx, MCB200_ESTATE, "mcb200_set_xsec and set_spectra first");
    if (ctx->pending) return fail(ctx, MCB200_ESTATE, "tallies pending: call mcb200_reduce first");
    if (nBands < 1 || !bandOff || !bandLow || !bandHigh) return fail(ctx, MCB200_EINVAL, "bad band list");
    if ((nPhotoDif || heatDif) && !ctx->cfg.lgDebug) return fail(ctx, MCB200_ESTATE, "Jdif only exists in debug mode");
    int rc = ensure_estimators(ctx, *g);
    if (rc) return rc;
    const int nb = ctx->cfg.nbins;
    const int64_t nXs = (int64_t)ctx->hXsec.size();
    std::vector<int> hi(nBands), heatHigh(nBands), nuStart(nb + 1, 0), nuBand;
    for (int b = 0; b < nBands; ++b) {
        int lo = bandLow[b];
        hi[b] = bandHigh[b] < nb ? bandHigh[b] : nb;
        if (lo < 1 || lo > nb) return fail(ctx, MCB200_EINVAL, "band %d: first bin %d out of range", b + 1, lo);
        if (hi[b] >= lo && (bandOff[b] < 1 || (int64_t)bandOff[b] + (hi[b] - lo) > nXs))
            return fail(ctx, MCB200_EINVAL, "band %d: xSecArray index out of range", b + 1);
        // thermBalance leaves the frequency loop at the first cross-section below 1e-35 (:1186)
        heatHigh[b] = hi[b];
        for (int j = lo; j <= hi[b]; ++j)
            if (ctx->hXsec[(size_t)bandOff[b] - 1 + (j - lo)] < 1.e-35f) { heatHigh[b] = j - 1; break; }
    }
    for (int nu = 1; nu <= nb; ++nu) {
        for (int b = 0; b < nBands; ++b)
            if (nu >= bandLow[b] && nu <= hi[b]) nuBand.push_back(b);
        nuStart[nu] = (int)nuBand.size();
    }
    if (nuBand.empty()) nuBand.push_back(0);
    cudaStream_t s = ctx->stream;
    DevBuf<int> dStart, dBand, dOff, dLow, dHeat;
    DevBuf<float> dRate, dHeatOut;
    size_t nR = (size_t)g->nCells + 1, outN = nR * (size_t)nBands;
    CU(dStart.upload(nuStart.data(), nuStart.size(), s)); CU(dBand.upload(nuBand.data(), nuBand.size(), s));
    CU(dOff.upload(bandOff, nBands, s)); CU(dLow.upload(bandLow, nBands, s)); CU(dHeat.upload(heatHigh.data(), nBands, s));
    CU(dRate.alloc(outN)); CU(dHeatOut.alloc(outN));
    PhotoArgs A{};
    A.nCells = g->nCells; A.nb = nb; A.sym = ctx->cfg.lgSymmetricXYZ;
    A.nuStart = dStart.p; A.nuBand = dBand.p; A.off = dOff.p; A.low = dLow.p; A.heatHigh = dHeat.p;
    A.xSec = ctx->xSec.p; A.nuArray = ctx->nuArray.p; A.nPhoto = dRate.p; A.heat = dHeatOut.p;
    const int perLaunch = 96;                 // 2 * 96 * 128 * 4 B = 96 KB of shared memory
    for (int pass = 0; pass < 2; ++pass) {
        float *outR = pass ? nPhotoDif : nPhotoSte, *outH = pass ? heatDif : heatSte;
        if (!outR && !outH) continue;
        A.J = pass ? g->Jdif.p : g->Jste.p;
        if (!A.J) return fail(ctx, MCB200_ESTATE, "estimator missing");
        for (int b0 = 0; b0 < nBands; b0 += perLaunch) {
            A.b0 = b0; A.nB = nBands - b0 < perLaunch ? nBands - b0 : perLaunch;
            CU(launch_photo(A, s));
        }
        if (outR) CU(cudaMemcpyAsync(outR, dRate.p, outN * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (outH) CU(cudaMemcpyAsync(outH, dHeatOut.p, outN * sizeof(float), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    return MCB200_OK;
}

int mcb200_assemble_opacity(mcb200_ctx *ctx, int32_t iG, int32_t nBands, const int32_t *bandSpecies,
                            const int32_t *bandOff, const int32_t *bandLow, const int32_t *bandHigh,
                            int32_t nSpeciesDen, const float *den, const float *ff1, const float *Ndust,
                            const float *Tdust, const int32_t *dustAbunIndex, const float *grainWeight,
                            const int32_t *dustScaXsecP, const int32_t *dustAbsXsecP, int32_t nSpeciesTot)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    if (!ctx->xSec.p) return fail(ctx, MCB200_ESTATE, "mcb200_set_xsec first");
    const mcb200_config &c = ctx->cfg;
    int nb = c.nbins, nRows = g->nCells + 1;
    if (nBands < 0 || nSpeciesDen < 0 || (nBands > 0 && (!bandSpecies || !bandOff || !bandLow || !bandHigh || !den)))
        return fail(ctx, MCB200_EINVAL, "bad band list");
    int64_t nXs = (int64_t)ctx->xSec.n;
    // CSR over nu of the bands covering each bin, in the reference's band order
    // (inOpacity, ionization_mod.f90:448-482: i = nuLowP .. max(nuLowP, min(nuHighP, nbins)))
    std::vector<int> nuStart(nb + 1, 0), spec, xs;
    for (int nu = 1; nu <= nb; ++nu) {
        for (int b = 0; b < nBands; ++b) {
            int lo = bandLow[b], up = bandHigh[b] < nb ? bandHigh[b] : nb;
            if (up < lo) up = lo;
            if (nu < lo || nu > up) continue;
            int sp = bandSpecies[b];
            int64_t xi = (int64_t)nu + bandOff[b];
            if (sp < 1 || sp > nSpeciesDen) return fail(ctx, MCB200_EINVAL, "band %d: species column out of range", b + 1);
            if (xi < 1 || xi > nXs) return fail(ctx, MCB200_EINVAL, "band %d: xSecArray index out of range", b + 1);
            spec.push_back(sp - 1);
            xs.push_back((int)xi);
        }
        nuStart[nu] = (int)spec.size();
    }
    size_t ts = tsize(ctx, *g);
    bool re = (g->opacity.n != ts);
    // device temporaries live in the context: cudaMalloc / cudaFree per call cost more than K1
    DevBuf<float> &dDen = ctx->opDen, &dFf = ctx->opFf, &dNd = ctx->opNd, &dCoef = ctx->opCoef;
    DevBuf<int> &dStart = ctx->opStart, &dSpec = ctx->opSpec, &dXs = ctx->opXs, &dComp = ctx->opComp,
                &dTermOn = ctx->opTermOn, &dScaP = ctx->opScaP, &dAbsP = ctx->opAbsP;
    DevBuf<unsigned char> &dOn = ctx->opOn;
    cudaStream_t s = ctx->stream;
    const bool tr = ctx->trace;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    auto t0 = now();
    OpacityArgs A{};
    A.nRows = nRows; A.nb = nb; A.nSpeciesDen = nSpeciesDen;
    if (nSpeciesDen > 0) CU(dDen.upload(den, (size_t)nRows * nSpeciesDen, s));
    if (ff1) CU(dFf.upload(ff1, nRows, s));
    CU(dStart.upload(nuStart.data(), nuStart.size(), s));
    if (!spec.empty()) { CU(dSpec.upload(spec.data(), spec.size(), s)); CU(dXs.upload(xs.data(), xs.size(), s)); }
    A.den = dDen.p; A.ff1 = dFf.p; A.xSec = ctx->xSec.p;
    A.nuStart = dStart.p; A.nuBandSpecies = dSpec.p; A.nuBandXs = dXs.p;
    CU(g->opacity.alloc(ts));
    A.opacity = g->opacity.p;
    std::vector<unsigned char> on;
    std::vector<float> coef;
    std::vector<int> termOn, scaP, absP, comps;
    if (c.lgDust && Ndust) {
        if (!ctx->haveDustSpecies) return fail(ctx, MCB200_ESTATE, "set_dust_species first");
        if (!grainWeight || !dustScaXsecP || !dustAbsXsecP || nSpeciesTot < 1 || c.nSizes < 1)
            return fail(ctx, MCB200_EINVAL, "bad dust arguments");
        // Tdust = NULL: take the sublimation mask from the device-resident dust state
        // (mcb200_set_dust_state / mcb200_dust_update), which then never leaves the device
        const bool devT = Tdust == nullptr;
        if (devT && (!ctx->haveDustTables || g->Tdust.n != ((size_t)c.nSpeciesMax + 1) * ((size_t)c.nSizes + 1) * (size_t)nRows ||
                     nSpeciesTot != ctx->nSpeciesTot))
            return fail(ctx, MCB200_ESTATE, "Tdust = NULL needs the device dust state: mcb200_set_dust_tables and mcb200_set_dust_state first");
        if (c.lgMultiDustChemistry && !dustAbunIndex) return fail(ctx, MCB200_EINVAL, "dustAbunIndex required");
        int nT = nSpeciesTot * c.nSizes, nC = c.nDustComp;
        coef.assign((size_t)nC * nT, 0.f); termOn.assign((size_t)nC * nT, 0);
        scaP.resize(nT); absP.resize(nT);
        for (int sg = 1; sg <= nSpeciesTot; ++sg)
            for (int ai = 1; ai <= c.nSizes; ++ai) {
                int t = (sg - 1) * c.nSizes + (ai - 1);
                scaP[t] = dustScaXsecP[(size_t)(sg - 1) + (size_t)nSpeciesTot * (ai - 1)];
                absP[t] = dustAbsXsecP[(size_t)(sg - 1) + (size_t)nSpeciesTot * (ai - 1)];
                if (scaP[t] < 1 || scaP[t] + nb - 1 > nXs || absP[t] < 1 || absP[t] + nb - 1 > nXs)
                    return fail(ctx, MCB200_EINVAL, "dust cross-section pointer out of range");
                for (int k = 1; k <= nC; ++k) {
                    int dcp = ctx->dustComPoint[k - 1], nS = sg - dcp + 1;
                    if (nS < 1 || nS > ctx->nSpeciesPart[k - 1]) continue;
                    termOn[(size_t)(k - 1) * nT + t] = 1;
                    volatile float cf = ctx->grainAbun[(size_t)(k - 1) + (size_t)nC * (nS - 1)] * grainWeight[ai - 1];
                    coef[(size_t)(k - 1) * nT + t] = cf;
                }
            }
        // sublimation mask Tdust(nS,ai,cell) < TdustSublime(dcp-1+nS), iteration_mod.f90:189
        if (!devT) on.assign((size_t)nT * nRows, 0);
        comps.assign(nRows, 0);
        size_t s0 = (size_t)c.nSpeciesMax + 1, s1 = (size_t)c.nSizes + 1;
        for (int cell = 1; cell < nRows; ++cell) {
            int k = c.lgMultiDustChemistry ? dustAbunIndex[cell] : 1;
            if (k < 1 || k > nC) { comps[cell] = -1; continue; }
            comps[cell] = k - 1;
            if (devT) continue;
            int dcp = ctx->dustComPoint[k - 1];
            for (int nS = 1; nS <= ctx->nSpeciesPart[k - 1]; ++nS)
                for (int ai = 1; ai <= c.nSizes; ++ai) {
                    int sg = dcp - 1 + nS, t = (sg - 1) * c.nSizes + (ai - 1);
                    float Td = Tdust[(size_t)nS + s0 * ((size_t)ai + s1 * (size_t)cell)];
                    on[(size_t)t * nRows + cell] = Td < ctx->TdustSublime[sg - 1] ? 1 : 0;
                }
        }
        auto t1 = now();
        if (tr) fprintf(stderr, "[mcb200] assemble_opacity: gas uploads + dust host mask %.1f ms\n", ms(t0, t1));
        CU(dNd.upload(Ndust, nRows, s));
        CU(dCoef.upload(coef.data(), coef.size(), s)); CU(dTermOn.upload(termOn.data(), termOn.size(), s));
        CU(dScaP.upload(scaP.data(), nT, s)); CU(dAbsP.upload(absP.data(), nT, s));
        CU(dComp.upload(comps.data(), nRows, s));
        if (devT) {
            CU(dOn.alloc((size_t)nT * nRows));
            CU(launch_dust_mask(g->Tdust.p, dComp.p, ctx->dComPoint.p, ctx->dSpeciesPart.p, ctx->dSublime.p, nRows, nSpeciesTot,
                                c.nSizes, c.nSpeciesMax + 1, c.nSizes + 1, dOn.p, s));
        } else {
            CU(dOn.upload(on.data(), on.size(), s));
        }
        if (g->scaOpac.n != ts) re = true;
        CU(g->scaOpac.alloc(ts)); CU(g->absOpac.alloc(ts));
        A.nDustTerms = nT; A.Ndust = dNd.p; A.dustOn = dOn.p; A.dustCoef = dCoef.p;
        A.dustCompOfCell = dComp.p; A.dustTermOn = dTermOn.p; A.dustScaP = dScaP.p; A.dustAbsP = dAbsP.p;
        A.scaOpac = g->scaOpac.p; A.absOpac = g->absOpac.p;
    } else if (c.lgDust) {
        // lgEquivalentTau first iteration (iteration_mod.f90:169-171): dust opacities stay zero
        if (g->scaOpac.n != ts) re = true;
        CU(g->scaOpac.alloc(ts)); CU(g->absOpac.alloc(ts));
        CU(g->scaOpac.zero(s)); CU(g->absOpac.zero(s));
    }
    auto t2 = now();
    CU(launch_opacity(A, s));
    CU(cudaStreamSynchronize(s));
    if (tr) fprintf(stderr, "[mcb200] assemble_opacity: total host prep %.1f ms, kernel + pending copies %.1f ms\n", ms(t0, t2), ms(t2, now()));
    g->haveOpacity = true;
    if (re) ctx->gridsDirty = true;
    return MCB200_OK;
}

int mcb200_get_opacity_rows(mcb200_ctx *ctx, int32_t iG, int32_t nWanted, const int32_t *cells, float *rows)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set) return fail(ctx, MCB200_ESTATE, "grid %d not set", iG);
    if (!g->haveOpacity || !g->opacity.p) return fail(ctx, MCB200_ESTATE, "grid %d has no opacity yet", iG);
    if (nWanted < 0 || (nWanted > 0 && (!cells || !rows))) return fail(ctx, MCB200_EINVAL, "bad get_opacity_rows arguments");
    if (nWanted == 0) return MCB200_OK;
    for (int r = 0; r < nWanted; ++r)
        if (cells[r] < 0 || cells[r] > g->nCells) return fail(ctx, MCB200_EINVAL, "cell %d out of range 0..%d", cells[r], g->nCells);
    const int nb = ctx->cfg.nbins;
    DevBuf<int> dCells;
    DevBuf<float> dOut;
    CU(dCells.upload(cells, (size_t)nWanted, ctx->stream));
    CU(dOut.alloc((size_t)nWanted * nb));
    CU(launch_gather_rows(g->opacity.p, (size_t)g->nCells + 1, nb, dCells.p, nWanted, dOut.p, ctx->stream));
    CU(cudaMemcpyAsync(rows, dOut.p, (size_t)nWanted * nb * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MCB200_OK;
}

int mcb200_get_opacity(mcb200_ctx *ctx, int32_t iG, float *opacity, float *scaOpac, float *absOpac)
{
    NEED_CTX();
    GridState *g = grid_of(ctx, iG);
    if (!g || !g->set || !g->haveOpacity) return fail(ctx, MCB200_ESTATE, "opacity of grid %d not set", iG);
    CU(cudaStreamSynchronize(ctx->stream));
    if (opacity) CU(cudaMemcpy(opacity, g->opacity.p, g->opacity.n * sizeof(float), cudaMemcpyDeviceToHost));
    if (scaOpac) {
        if (!g->scaOpac.p) return fail(ctx, MCB200_ESTATE, "no scaOpac on device");
        CU(cudaMemcpy(scaOpac, g->scaOpac.p, g->scaOpac.n * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (absOpac) {
        if (!g->absOpac.p) return fail(ctx, MCB200_ESTATE, "no absOpac on device (host-assembled opacity)");
        CU(cudaMemcpy(absOpac, g->absOpac.p, g->absOpac.n * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return MCB200_OK;
}

}  // extern "C"
