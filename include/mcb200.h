/*
 * mcb200.h -- C ABI of libmocassin_b200.so: the B200 (sm_100a) implementation of
 * mocassin's energy-packet transport hot path.
 *
 * The reference (rwesson/mocassin, Fortran 90 + MPI) has no FFI for this path; the
 * seam is a set of Fortran calls inside the Lucy iteration `iterateMC`
 * (source/iteration_mod.f90).  Each entry point below names the reference
 * interface it replaces.  Signatures are plain C (pointers + sizes, no C++/torch
 * types) so the reference's Fortran driver binds them with ISO_C_BINDING
 * (fortran/mcb200_mod.f90).
 *
 * Conventions
 *  - All host arrays stay owned by the caller; the library copies in/out.
 *  - Arrays use the reference's own layouts (column major, see each function), and
 *    indices stored inside arrays (active, starIndeces) are 1-based, so a Fortran
 *    caller passes c_loc(array) unchanged.
 *  - Every function returns 0 on success or a negative MCB200_E* code;
 *    mcb200_last_error() gives the message.  The reference's behaviour on the same
 *    conditions is `print*; stop` (e.g. photon_mod.f90:826-832); the Fortran shim
 *    turns a non-zero status into exactly that.
 *  - One host thread per context; all calls are synchronous on return unless noted.
 *  - There is no CPU fallback: without a CUDA device mcb200_create fails.
 */
#ifndef MCB200_H
#define MCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB200_OK            0
#define MCB200_ENODEV       -1   /* no CUDA device / CUDA runtime error            */
#define MCB200_EINVAL       -2   /* bad argument                                    */
#define MCB200_ESTATE       -3   /* call sequence error (e.g. transport before set) */
#define MCB200_ENOMEM       -4   /* device allocation failed                        */
#define MCB200_EPACKET      -5   /* a packet hit one of the reference's `stop`s     */
#define MCB200_EUNSUPPORTED -6   /* nested sub-grids / lg1D                          */
#define MCB200_ETABLE       -7   /* CDF table not monotone / bad range              */
#define MCB200_ECOMM        -8   /* NCCL unavailable or a collective failed         */

typedef struct mcb200_ctx mcb200_ctx;   /* opaque */

/* Global flags and scalars of common_mod read by photon_mod (set once, after
 * setStarPosition, mocassin.f90:83).  Logical flags are 0/1. */
typedef struct mcb200_config {
    int32_t nGrids;                /* common_mod nGrids                                 */
    int32_t nbins;                 /* number of frequency bins                          */
    int32_t nStars;
    int32_t nAngleBins;            /* viewing angles (`inclination` keyword)            */
    int32_t totAngleBinsTheta;     /* common_mod.f90:399 (10)                           */
    int32_t totAngleBinsPhi;       /* common_mod.f90:400 (20, or 1: grid_mod.f90:416-429)*/
    int32_t nLines;                /* size of linePackets/linePDF 2nd dim (debug only)  */
    int32_t lgDust, lgGas, lgSymmetricXYZ, lgIsotropic, lgPlaneIonization, lgDebug,
            lgMultistars, lgMultiDustChemistry;
    int32_t nSpeciesMax, nSizes, nDustComp;   /* Tdust / grainAbun extents              */
    float dTheta, dPhi;            /* grid_mod.f90:416-431                              */
    float R_out;                   /* outer radius [cm], 0 = unset                      */
    float ionEdge1;                /* ionEdge(1) [Ryd]                                  */
} mcb200_config;

/* Counters of one transport call (photon_mod module variables Qphot, absInt,
 * scaInt, trapped, photon_mod.f90:16,126,1702,1720,1803), as exact integers. */
typedef struct mcb200_counters {
    int64_t nPackets;              /* packets this rank transported                     */
    int64_t nAbs, nSca;            /* absInt, scaInt                                    */
    int64_t trapped;               /* recursionLimit hits                               */
    int64_t nLinePackets;          /* packets that left as non-ionising line packets    */
    int64_t nDropped;              /* safeLimit / outer-wall returns (no tally)         */
    int64_t nSegments;             /* trips of the cell-crossing loop :1194             */
    int64_t nFlights;              /* pathSegment calls                                 */
    int64_t nEscaped;              /* escape tallies                                    */
    int64_t nEarlyEscaped;         /* of which at energyPacketRun :370                  */
    double  Qphot;                 /* sum deltaE/(2.1799153e-11*nu), nu>1 Ryd (:859-861)*/
    double  kernel_ms;             /* device time of the transport kernel (CUDA events)  */
    double  total_ms;              /* device time transport kernel + fold epilogue        */
    int64_t nLaunches;             /* kernels launched by this call (transport + fold)    */
    int64_t nWaves;                /* wave-front schedule: waves (0 = persistent kernel)  */
    double  fly_ms;                /* wave-front schedule: device time of the cell-crossing (FLY) kernels alone */
} mcb200_counters;

/* ---- lifecycle ---------------------------------------------------------------- */

/* Create a context on CUDA device `device` for rank `rank` of `nranks`
 * (replaces nothing in the reference; called where mocassin.f90:54-56 has
 * mpi_comm_rank/size).  `seed` keys the Philox4x32-10 packet streams (the reference
 * reseeds from the wall clock, photon_mod.f90:68-87). */
int mcb200_create(mcb200_ctx **ctx, int32_t device, int32_t rank, int32_t nranks, uint64_t seed);
int mcb200_destroy(mcb200_ctx *ctx);                 /* before mpi_finalize         */
const char *mcb200_last_error(const mcb200_ctx *ctx);
int mcb200_set_config(mcb200_ctx *ctx, const mcb200_config *cfg);

/* ---- static inputs -------------------------------------------------------------- */

/* Geometry of grid iG (1-based): grid_type members nx,ny,nz,nCells,motherP,
 * xAxis,yAxis,zAxis,active (common_mod.f90:241-302) as built by fillGrid /
 * setMotherGrid / setSubGrids (grid_mod.f90:488,898,1804).  active(nx,ny,nz) int32,
 * x fastest.  geoCorr is recomputed as in grid_mod.f90:809-812. */
int mcb200_set_grid(mcb200_ctx *ctx, int32_t iG, int32_t nx, int32_t ny, int32_t nz,
                    int32_t nCells, int32_t motherP, const float *xAxis, const float *yAxis,
                    const float *zAxis, const int32_t *active);

/* Frequency-indexed globals: nuArray(nbins), gSca(nbins) (may be NULL without dust),
 * inSpectrumProbDen(0:nStars,nbins) (Fortran layout, star index fastest),
 * deltaE is passed per transport call. */
int mcb200_set_spectra(mcb200_ctx *ctx, const float *nuArray, const float *gSca,
                       const float *inSpectrumProbDen);

/* starPosition(nStars) as x,y,z triplets [cm] and starIndeces(nStars,4) (Fortran
 * layout, star index fastest) from setStarPosition (grid_mod.f90:3569-3648). */
int mcb200_set_stars(mcb200_ctx *ctx, const float *starPosition, const int32_t *starIndeces);

/* Viewing-angle tables of initCartesianGrid (grid_mod.f90:433-468):
 * viewPointPtheta(0:totAngleBinsTheta), viewPointPphi(0:totAngleBinsPhi),
 * viewPointTheta(0:nAngleBins), viewPointPhi(0:nAngleBins).  Only needed when
 * nAngleBins>0. */
int mcb200_set_viewpoints(mcb200_ctx *ctx, const int32_t *viewPointPtheta,
                          const int32_t *viewPointPphi, const float *viewPointTheta,
                          const float *viewPointPhi);

/* Dust species tables used by the sublimation test at scattering
 * (photon_mod.f90:1722-1748): nSpeciesPart(nDustComp), grainAbun(nDustComp,
 * nSpeciesMax), dustComPoint(nDustComp), TdustSublime(nSpecies). */
int mcb200_set_dust_species(mcb200_ctx *ctx, const int32_t *nSpeciesPart, const float *grainAbun,
                            const int32_t *dustComPoint, const float *TdustSublime,
                            int32_t nSpecies);

/* ---- per-iteration inputs --------------------------------------------------------- */

/* Host-assembled opacities of grid iG: opacity, scaOpac (0:nCells,nbins)
 * (scaOpac may be NULL without dust).  Use instead of mcb200_assemble_opacity when
 * the host keeps ionizationDriver (iteration_mod.f90:117-227). */
int mcb200_set_opacity(mcb200_ctx *ctx, int32_t iG, const float *opacity, const float *scaOpac);

/* Device opacity assembly (K1), replaces the ionizationDriver loop + dust add of
 * iteration_mod.f90:117-227 for grid iG:
 *   opacity(c,nu) = ff1(c)*[nu==1] + sum_b den(c, bandSpecies(b)) * xSec(nu + bandOff(b))
 *                   for bandLow(b) <= nu <= min(bandHigh(b),nbins)
 *                 + scaOpac(c,nu) + absOpac(c,nu)
 *   scaOpac/absOpac(c,nu) = sum_{s,a: Tdust(s,a,c)<TdustSublime(s)} grainAbun*grainWeight(a)
 *                            *Ndust(c)*xSec(dustSca/AbsXsecP(s,a)+nu-1)
 * The band list is the flattened (species,shell) list of addOpacity
 * (ionization_mod.f90:396-443): bandSpecies(b) indexes columns of den
 * (0:nCells, nSpeciesDen), i.e. density(elem,ion)=ionDen*elemAbun*Hden of
 * ionization_mod.f90:65-80.  ff1 (0:nCells) is the bin-1 free-free opacity of
 * addOpacity :369-393 (the only bin the reference ever fills), computed by the host's
 * BoltGaunt so its cell-order-dependent Gaunt-factor cache is preserved. All band
 * indices are 1-based Fortran values; xSecArray is passed once with its length.
 * Tdust = NULL: the sublimation test reads the dust temperatures the library holds on the
 * device (mcb200_set_dust_state, updated in place by mcb200_dust_update), so a dust-only Lucy
 * iteration -- mcb200_dust_pdf, mcb200_transport, mcb200_dust_update, mcb200_assemble_opacity
 * with nBands = 0 -- runs with nothing but counters crossing PCIe, sublimation included
 * (needs mcb200_set_dust_tables; MCB200_ESTATE otherwise). */
int mcb200_set_xsec(mcb200_ctx *ctx, const float *xSecArray, int64_t nXsec);
int mcb200_assemble_opacity(mcb200_ctx *ctx, int32_t iG,
                            int32_t nBands, const int32_t *bandSpecies, const int32_t *bandOff,
                            const int32_t *bandLow, const int32_t *bandHigh,
                            int32_t nSpeciesDen, const float *den, const float *ff1,
                            /* dust part, all NULL/0 without dust */
                            const float *Ndust, const float *Tdust, const int32_t *dustAbunIndex,
                            const float *grainWeight, const int32_t *dustScaXsecP,
                            const int32_t *dustAbsXsecP, int32_t nSpeciesTot);
/* rows(1:nWanted, 1:nbins) = opacity(cells(r), :) of grid iG for a short list of cells (0..nCells):
 * what writeTauNu / integratePathTauNu (output_mod.f90:2384-2505, pathIntegration_mod.f90:241-470)
 * read along their rays from the origin every iteration.  With the opacities assembled on the
 * device (mcb200_assemble_opacity) this replaces a download of the whole table (5 GB at 128^3 x
 * 600) by a few hundred KB; mocassin_b200/output.py: tau_nu, write_tau_nu is the host part. */
int mcb200_get_opacity_rows(mcb200_ctx *ctx, int32_t iG, int32_t nWanted, const int32_t *cells, float *rows);

/* Read back opacity / scaOpac / absOpac (0:nCells,nbins) of grid iG (any may be NULL). */
int mcb200_get_opacity(mcb200_ctx *ctx, int32_t iG, float *opacity, float *scaOpac, float *absOpac);

/* Re-emission tables of grid iG built by emissionDriver (iteration_mod.f90:279-424):
 * recPDF or dustPDF (0:nCells,nbins), totalLines(0:nCells) (NULL for dust-only),
 * linePDF (0:nCells,nLines) (debug only, else NULL).  Tables must be non-decreasing
 * along nu (they are cumulative sums); MCB200_ETABLE otherwise.  With option
 * "async_pdfs"=1 the call only enqueues upload+transpose on a copy stream (the buffers must
 * stay valid, ideally pinned, until the next transport call returns): the upload then
 * overlaps the stellar wave and MCB200_ETABLE is reported by that transport call. */
int mcb200_set_pdfs(mcb200_ctx *ctx, int32_t iG, const float *recPDF, const float *dustPDF,
                    const float *totalLines, const float *linePDF);

/* Dust temperatures Tdust(0:nSpeciesMax,0:nSizes,0:nCells) and dustAbunIndex(0:nCells)
 * of grid iG, read by the sublimation test at scattering. */
int mcb200_set_dust_state(mcb200_ctx *ctx, int32_t iG, const float *Tdust,
                          const int32_t *dustAbunIndex);

/* ---- dust-only closure of the iteration ("next" rows) ------------------
 * With these three calls a dust-only model (lgDust and not lgGas) iterates without the
 * estimators or the emission PDFs leaving the device:
 *     transport -> [reduce] -> mcb200_dust_update -> mcb200_dust_pdf -> transport ...
 *
 * mcb200_set_dust_tables: once, after set_dust_species and set_xsec and BEFORE
 *   set_dust_state (which then keeps a device copy of Tdust).  widFlx(nbins),
 *   grainWeight(nSizes), dustAbsXsecP(nSpecies,nSizes) (1-based offsets into xSecArray),
 *   dustEmIntegral(nSpecies,nSizes,nTemps) (dust_mod / xSec_mod; grid_mod.f90 tables).
 *
 * mcb200_dust_update: replaces the dust-only branch of updateCell (update_mod.f90:308-334)
 *   and getDustT (:1836-1945) for all cells of grid iG.  Reads the device-resident Jste
 *   (Jdif in debug mode) with the host scaling of iteration_mod.f90:705-724 applied on the
 *   fly (*1e-9, /8 for symmetricXYZ), updates the device Tdust and the sublimation flags of
 *   the next transport.  Optional outputs: Tdust(0:nSpeciesMax,0:nSizes,0:nCells),
 *   lgConverged(0:nCells), and the number of converged cells.  A cell no packet crossed (no
 *   Jste(cell,:) > 0, nor Jdif in debug mode) is left alone, as updateCell returns first thing
 *   for it (update_mod.f90:104-149): its Tdust stays, its lgConverged is the 0 iterateMC gives
 *   every cell at the start of an iteration (iteration_mod.f90:87).  A grain whose absorption
 *   integral is below dustEmIntegral(.,.,1) gets 1 K (the reference's lgTalk branch; without
 *   lgTalk the reference reads dustEmIntegral(.,.,0), out of bounds).  In multi-rank runs every
 *   rank holds the reduced Jste and updates all cells itself: no exchange (the reference
 *   splits cells over ranks and all-reduces TdustTemp, iteration_mod.f90:797-870).
 *
 * mcb200_dust_pdf: replaces setDustPDF (emission_mod.f90:1313-1387; quantum heating
 *   excluded) for all cells: builds the re-emission CDF rows of grid iG from the device
 *   Tdust, in place of mcb200_set_pdfs.  dustPDF (0:nCells,nbins), if not NULL, receives a
 *   copy in the reference's layout. */
int mcb200_set_dust_tables(mcb200_ctx *ctx, const float *widFlx, const float *grainWeight,
                           const int32_t *dustAbsXsecP, int32_t nSpecies,
                           const float *dustEmIntegral, int32_t nTemps);
int mcb200_dust_update(mcb200_ctx *ctx, int32_t iG, float XHILimit, float *Tdust,
                       int32_t *lgConverged, int64_t *nConverged);
int mcb200_dust_pdf(mcb200_ctx *ctx, int32_t iG, float *dustPDF);

/* ---- the hot path ------------------------------------------------------------------ */

/* Zero Jste, Jdif, linePackets, escapedPackets of all grids
 * (iteration_mod.f90:458-472). */
int mcb200_zero_estimators(mcb200_ctx *ctx);

/* Transport.  Replaces `call energyPacketDriver(iStar, load, grid)` for all ranks
 * (iteration_mod.f90:474-496): nPacketsGlobal is nPhotons(iStar); this rank takes
 * the reference's share, load=int(N/nranks), +1 if rank<mod(N,nranks), with
 * contiguous global packet ids, so results do not depend on nranks.
 * deltaE is deltaE(iStar).  The path-length and escape tallies of this call are
 * folded into the float32 estimators (Jste += L*deltaE/dV etc.) on return when
 * nranks==1; with nranks>1 the integer tallies stay pending until
 * mcb200_reduce (so the cross-rank sum is exact and order independent). */
int mcb200_transport(mcb200_ctx *ctx, int32_t iStar, int64_t nPacketsGlobal, float deltaE,
                     mcb200_counters *counters);

/* Extra diffuse source, replaces `call energyPacketDriver(iStar=0, n=load, grid,
 * gpLoc, cellLoc)` (iteration_mod.f90:498-550); deltaE is
 * LdiffuseLoc(cell)/NphotonsDiffuseLoc (photon_mod.f90:64-66). */
int mcb200_transport_diffuse(mcb200_ctx *ctx, int32_t gpLoc, const int32_t *cellLoc,
                             int64_t nPacketsGlobal, float deltaE, mcb200_counters *counters);

/* Resonance-line packet transfer, the second half of energyPacketDriver
 * (photon_mod.f90:180-266; the host decides when it runs: lgDust, convPercent >=
 * resLinesTransfer, not the first iteration).  resLinePackets(0:nCells) of grid iG is the
 * table emission_mod fills (emission_mod.f90:244); mcb200_transport_reslines then starts
 * that many "diffuse" packets at the centre of every cell this rank owns under the
 * reference's round-robin rule mod(iCell-(taskid+1),numtasks)==0, tallying with
 * deltaE(iStar) exactly like the loop it replaces. */
int mcb200_set_res_line_packets(mcb200_ctx *ctx, int32_t iG, const int32_t *resLinePackets);
int mcb200_transport_reslines(mcb200_ctx *ctx, int32_t iStar, float deltaE, mcb200_counters *counters);

/* Device pointers and element counts of the pending integer tallies of grid iG so
 * the caller's communicator can sum them across ranks in place (NCCL allreduce, sum)
 * -- replaces MPI_ALLREDUCE at iteration_mod.f90:627,649,653,659.  which: 0 JsteQ,
 * 2 JdifQ (int64 fixed-point path lengths); 1 escapedQ, 3 linePacketsQ (uint32 packet
 * counts: the global packet count of one call must stay below 2^32); 4 nuTouched
 * (int32[nbins+1], 1 = a packet was emitted in that frequency bin since the last fold:
 * only those nu-planes of the tallies can be non-zero, so an exchange may max-reduce the
 * flags first and then sum only the flagged planes -- plane nu of JsteQ is the contiguous
 * run [(nu-1)*(nCells+1), nu*(nCells+1)), plane (nu,ang) of escapedQ starts at
 * (nCells+1)*(nu + (nbins+1)*ang)); 5 planeIonDistribution (int32); 6 sedQ (uint64
 * (nbins+1)*(nAngleBins+1), iG ignored; see mcb200_fetch_sed, option "sed_local");
 * 16, 17, 20 = buffers 0, 1, 4 of the second tally set (option "tally_set"). */
int mcb200_tally_buffer(mcb200_ctx *ctx, int32_t iG, int32_t which, void **devPtr, int64_t *count);
/* Fold only the nu-planes [nu0, nu1] of grid iG (J planes nu >= 1, escape-count planes nu0..nu1 of
 * every angle) and return without waiting: lets the caller fold the planes whose all-reduce has
 * finished while later planes are still being exchanged.  mcb200_reduce afterwards folds whatever
 * is left and closes the call.  Not available while a second tally set is pending. */
int mcb200_reduce_range(mcb200_ctx *ctx, int32_t iG, int32_t nu0, int32_t nu1);

/* Sparse form of the escapedQ exchange.  escapedPackets is indexed by the cell a packet was
 * last emitted or scattered in, so only a few per cent of its entries are non-zero after a
 * call: instead of all-reducing the dense array (5 GB at 128^3 x 600) each rank
 *   1. mcb200_escaped_compact: moves its non-zero (index, count) pairs of grid iG into a device
 *      list (uint64 index, uint64 count per entry; the array itself is cleared),
 *   2. all-gathers the lists (NCCL all_gather, or MPI_Allgatherv on a CUDA-aware MPI),
 *   3. mcb200_escaped_scatter: adds every rank's list (its own included) back into the array.
 * Integer adds: the result equals the dense all-reduce bit for bit.  set = 0, or 1 for the
 * second tally set. */
int mcb200_escaped_compact(mcb200_ctx *ctx, int32_t iG, int32_t set, void **devList, int64_t *nEntries);
int mcb200_escaped_scatter(mcb200_ctx *ctx, int32_t iG, int32_t set, const void *devList, int64_t nEntries);
/* The exchange done by the library itself: a NCCL communicator owned by the context, for a
 * host (the Fortran/MPI reference) that has no device-aware collective of its own.  Replaces
 * the MPI_ALLREDUCE block iteration_mod.f90:564,627,649,653,659 together with mcb200_reduce.
 *   rank 0:     mcb200_comm_unique_id(ctx, id)         128-byte ncclUniqueId
 *   host:       MPI_BCAST(id, 128, MPI_BYTE, 0, ...)   the only host-side message
 *   every rank: mcb200_comm_init(ctx, id)              rank / nranks as given to mcb200_create
 *   per source: mcb200_transport(...); mcb200_exchange(ctx); mcb200_reduce(ctx);
 * mcb200_exchange sums the pending integer tallies of every grid over the ranks on the library
 * stream: max-reduce of the nuTouched flags, then only the flagged nu-planes.  JsteQ (and JdifQ in
 * debug mode) are REDUCE-SCATTERED: every rank receives the global sums of 1/nranks of each
 * touched range; mcb200_reduce then folds that share only and all-gathers the float32 estimator
 * in place -- 12 instead of the 16 bytes per element an int64 all-reduce moves, and the fold
 * divided by nranks (option "exchange_allreduce"=1: the plain all-reduce + full fold).
 * linePacketsQ (debug), planeIonDistribution: all-reduced; the escape counts: all-gathered sparse
 * (index, count) lists when those are shorter than the dense planes (with option "sed_local":
 * the (nu, angle) counts of buffer 6 instead).  Integer sums and one fold per element: the
 * folded estimators are bit-identical on every rank and for every rank count.  A second tally
 * set (option "tally_set") is merged first.  No-op for nranks = 1 or when nothing is pending;
 * MCB200_ESTATE if the pending tallies were exchanged already, or if a transport call follows an
 * exchange without mcb200_reduce in between (either would count the other ranks' packets twice).
 * NCCL is bound at run time (dlopen, RTLD_LOCAL, of $MCB200_NCCL_LIB, libnccl.so.2, libnccl.so): the
 * library has no link-time dependency on it; MCB200_ECOMM if it cannot be found.
 * mcb200_exchange_info: bytes this rank handed to NCCL in the last exchange, the number of
 * grids whose escape counts went sparse, and ncclGetVersion (0 if NCCL is not loaded); any
 * pointer may be NULL. */
int mcb200_comm_unique_id(mcb200_ctx *ctx, void *id128);
int mcb200_comm_init(mcb200_ctx *ctx, const void *id128);
int mcb200_comm_destroy(mcb200_ctx *ctx);
int mcb200_exchange(mcb200_ctx *ctx);
int mcb200_exchange_info(mcb200_ctx *ctx, int64_t *bytesSent, int32_t *sparseGrids, int32_t *ncclVersion);
/* How the last mcb200_exchange merged the J tallies: 1 = NCCL all-reduce (option
 * "exchange_allreduce"), 2 = NCCL reduce-scatter + fold of the share + all-gather of the float32
 * estimator, 3 = the fused peer-memory merge: every rank's JsteQ, Jste and a receive buffer are
 * mapped into every process (cudaIpc over NVLink / NVSwitch); mcb200_exchange pushes each peer's
 * share of this rank's partial sums into the peer's receive buffer (packed: low 32 bits, high 32
 * bits of flagged 256-element blocks only, nothing of all-zero blocks; posted peer stores, beside
 * the escape-count exchange), and in mcb200_reduce ONE kernel per rank adds the integers of its
 * share, folds once and stores the float32 result into all ranks' Jste -- reduce-scatter, fold and
 * all-gather as two passes of peer stores, no intermediate buffers.  Path 3 is taken when the
 * buffers can be mapped (same node, peer access; not in debug mode), else 2; `why` receives the
 * reason peer memory was not used.  Option "exchange_p2p": -1 auto (default), 0 never, 1 required
 * (MCB200_ECOMM if unavailable).  All paths leave bit-identical estimators.  phaseMs (nullable, 4
 * doubles): host time of the last mcb200_exchange, host time of the last mcb200_reduce, device time
 * of the J merge inside that fold (barrier + merge kernel + barrier, or fold + all-gather), and of
 * the pushes.  MCB200_TRACE_EXCHANGE=1 in the environment: rank 0 prints the host-clock timeline of
 * every exchange + merge to stderr. */
int mcb200_exchange_path(mcb200_ctx *ctx, int32_t *path, char *why, int64_t whyLen, double *phaseMs);
/* Which NCCL the library binds (needs no context and no device): ncclGetVersion and the file the
 * symbols came from.  A process that also hosts another NCCL user (PyTorch) must bind the SAME
 * copy: the dynamic loader keeps one object per SONAME, so whichever libnccl.so.2 is opened first
 * serves both -- point $MCB200_NCCL_LIB at the newer one (mocassin_b200/_lib.py does, for torch's
 * bundled copy).  MCB200_ECOMM if NCCL cannot be loaded. */
int mcb200_nccl_info(int32_t *version, char *path, int64_t pathLen);

/* After the allreduce: fold the (now global) integer tallies of the last transport
 * call into the float32 estimators. No-op when nothing is pending. */
int mcb200_reduce(mcb200_ctx *ctx);

/* SED(1:nbins, 0:nAngleBins) = sum over grids and cells of escapedPackets(i, freq, imu):
 * the reduction at the head of writeSED (output_mod.f90:2561-2568), done on the device so the
 * (cell, nu, angle) array never has to leave the GPU for the SED.  Raw sums in the same sense
 * as mcb200_fetch_estimators (before the host's /8 of iteration_mod.f90:719 and writeSED's own
 * *8, *4 and Jy conversion, which stay on the host).  Each transport call contributes
 * float(packets escaped in (freq, imu)) * deltaE -- the integer count is exact and independent
 * of order and rank count; the reference adds deltaE packet by packet in float32.
 * counts (nullable) receives the cumulative integer counts.  Zeroed by mcb200_zero_estimators.
 * Multi-rank: with option "sed_local"=1 the per-call counts are taken from this rank's own
 * escapes at the end of mcb200_transport and exposed as tally buffer 6 (uint64,
 * (nbins+1)*(nAngleBins+1)); the caller all-reduces that instead of buffer 1, and
 * escapedPackets then stays rank-local (only its sum over cells, the SED, is global). */
int mcb200_fetch_sed(mcb200_ctx *ctx, float *SED, int64_t *counts);

/* contI(0:nCells, 0:nAngleBins) = sum over freq = 1..nbins of escapedPackets(cell, freq, imu): the
 * reduction inside writeContCube (output_mod.f90:2762-2772), done on the device from the folded
 * float32 estimator in the reference's order (running float32 sum, freq ascending), so it equals
 * the reference's loop over the array mcb200_fetch_estimators returns bit for bit -- and that array
 * (5 GB at 128^3 x 600) need not be downloaded for output/contCube.out.  Raw sums: the host's /8 of
 * iteration_mod.f90:719 (exact in binary) and writeContCube's /dTheta, /dPhi, /(4 Pi) and file
 * layout stay on the host (mocassin_b200/output.py: write_cont_cube). */
int mcb200_fetch_contcube(mcb200_ctx *ctx, int32_t iG, float *contI);

/* Photo-rate pre-integration for updateCell: instead of shipping Jste
 * (5 GB at 128^3 x 600) to the host solver, integrate it on the device against each ion's
 * outer-shell cross-section.  Band b = (bandOff: 1-based index in xSecArray of the
 * cross-section at bin bandLow; bandLow = IPnuP; bandHigh = highNuP, clipped to nbins) -- the
 * (elem, ion) loop of update_mod.f90:170-213 flattened by the host.  Outputs
 * (0:nCells, nBands), cell index fastest, any may be NULL:
 *   nPhotoSte(cell,b) = 1e-20 + sum_{j, Jste>0} Jste*phXSec/(hcRyd*nuArray(j))   (:240-245)
 *   heatSte(cell,b)   = sum_{j until phXSec<1e-35, Jste>0} phXSec*Jste*(nu(j)-nu(IP))/nu(j)
 *                       (thermBalance :1196-1201, before the ionDen*elemAbun factor)
 * and the same from Jdif in debug mode.  Jste is the device-resident estimator with the host
 * scaling of iteration_mod.f90:705-724 applied on the fly.  Needs mcb200_set_xsec. */
int mcb200_photo_integrals(mcb200_ctx *ctx, int32_t iG, int32_t nBands, const int32_t *bandOff,
                           const int32_t *bandLow, const int32_t *bandHigh, float *nPhotoSte,
                           float *heatSte, float *nPhotoDif, float *heatDif);

/* Copy the raw estimator sums of grid iG into the caller's arrays, laid out as
 * Jste(0:nCells,nbins), escapedPackets(0:nCells,0:nbins,0:nAngleBins),
 * Jdif(0:nCells,nbins), linePackets(0:nCells,nLines) (NULL = skip).  These are the
 * values the reference holds after its MPI_ALLREDUCE block (iteration_mod.f90:
 * 583-703), *before* the host's own scaling (:705-724), which therefore stays
 * unchanged.  Row 0 of Jste/Jdif (the inactive-cell sink, never read by the
 * reference) is left zero. */
int mcb200_fetch_estimators(mcb200_ctx *ctx, int32_t iG, float *Jste, float *escapedPackets,
                            float *Jdif, float *linePackets);

/* The estimator rows of the cells ONE rank works on.  After the merge the reference's ranks share
 * the cells round robin -- rank r updates the cells with mod(iCell-(r+1), numtasks) == 0
 * (iteration_mod.f90:832) -- so a rank only ever reads Jste(iCell, :) of its own cells.
 * Jste (and Jdif in debug mode; either may be NULL) receive the compact array (1:nMine, 1:nbins),
 * cell index fastest, row j = cell firstCell + (j-1)*cellStride (1-based; pass taskid+1 and
 * numtasks): nMine*nbins*4 bytes cross PCIe instead of (nCells+1)*nbins*4.  The host indexes it with
 * (iCell - firstCell)/cellStride + 1 where it indexed grid%Jste(iCell, :).
 * nCellsOut (nullable) = nMine. */
int mcb200_fetch_estimators_cells(mcb200_ctx *ctx, int32_t iG, int32_t firstCell, int32_t cellStride, float *Jste,
                                  float *Jdif, int64_t *nCellsOut);

/* 64-bit checksum of a device-resident float32 estimator of grid iG, as mcb200_fetch_estimators
 * would return it (which: 0 Jste, 1 escapedPackets, 2 Jdif, 3 linePackets): the sum over the
 * elements of bits(i) * (2*i + 1) mod 2^64.  Lets every rank of a multi-GPU run show that it holds
 * the same estimators as every other rank, and as a single-rank run of the same packets (bench.py
 * "nrank_parity"; the arrays the reference compares after MPI_ALLREDUCE, iteration_mod.f90:583-703),
 * without 10 GB per rank crossing PCIe. */
int mcb200_checksum(mcb200_ctx *ctx, int32_t iG, int32_t which, uint64_t *sum);

/* escapedPackets of grid iG, sparsely.  escapedPackets(cell, nu, angle) is indexed by the cell a
 * packet was last emitted or scattered in, so in a large grid almost all of it is zero (0.5 % of
 * the 1.26e9 entries at 128^3 x 600): the library compacts the non-zero entries on the device,
 * moves only those over PCIe (48 MB instead of 5 GB) and writes them into the caller's array.
 * Contract: the array is zero where this call does not write -- iterateMC zeroes
 * grid%escapedPackets before the packet loop (iteration_mod.f90:466-470), so a Fortran host gets
 * exactly what mcb200_fetch_estimators would give.  clearPrevious != 0: entries written by the
 * previous call into the same array are zeroed first (for a caller that reuses the array without
 * zeroing it).  Falls back to the dense copy when more than 1/64 of the entries are non-zero.
 * nNonZero (optional): entries written, -1 after a dense fallback. */
int mcb200_fetch_escaped_sparse(mcb200_ctx *ctx, int32_t iG, float *escapedPackets, int32_t clearPrevious,
                                int64_t *nNonZero);
/* Jste (dense, as mcb200_fetch_estimators) and escapedPackets (sparse, as above) in one call: the
 * Jste copy crosses PCIe while the host threads write the escapedPackets entries.  Jste should be
 * page-locked (cudaHostRegister / pinned) for the two to overlap; pageable memory works, serially. */
int mcb200_fetch_estimators_sparse(mcb200_ctx *ctx, int32_t iG, float *Jste, float *escapedPackets,
                                   int32_t clearPrevious, int64_t *nNonZero);

/* Diagnostics: raw integer tallies (same shapes as above, int64) and the
 * path-length unit [cm] of grid iG's fixed-point J tally: 2^e, e = max(floor(log2(smallest cell
 * half-width)) - 24, floor(log2(widest cell half-width)) - 33), chosen by mcb200_set_grid. */
int mcb200_fetch_tallies(mcb200_ctx *ctx, int32_t iG, int64_t *JsteQ, int64_t *escapedQ,
                         int64_t *JdifQ, int64_t *linePacketsQ);
int mcb200_len_unit(mcb200_ctx *ctx, int32_t iG, double *lenUnit);
/* planeIonDistribution(grid(1)%nx, grid(1)%nz) of plane-parallel runs (packets emitted per
 * (x,z) of the y=0 face, photon_mod.f90:643-646; written to planeIonDistribution.out by
 * iteration_mod.f90:553-581); accumulated since the last mcb200_zero_estimators.  For
 * nranks>1 sum it across ranks (mcb200_tally_buffer which=5, int32). */
int mcb200_fetch_plane_distribution(mcb200_ctx *ctx, int32_t *planeIonDistribution);
/* Number of stellar emissions per frequency bin with nu>1 Ryd of the last transport
 * call (Qphot = sum counts*deltaE/(2.1799153e-11*nu)). */
int mcb200_fetch_qphot_counts(mcb200_ctx *ctx, int64_t *counts);
/* Per-packet fate records of the last call (4 int32 each: segments, generations,
 * last nuP, fate); enable with mcb200_set_option("trace",1) before transport. */
int mcb200_fetch_fates(mcb200_ctx *ctx, int32_t *fates, int64_t nPackets);
/* Tuning and diagnostic options; none of them changes a result bit (tallies are order-independent
 * integers, every packet owns its random stream).
 *   "seed"          Philox seed (also the `seed` of mcb200_create)
 *   "epoch"         advances the Philox key: key = seed + epoch * 0x9E3779B97F4A7C15 (mod 2^64).  The reference draws
 *                   fresh random numbers in every Lucy iteration; a host reproduces that by setting epoch to its
 *                   iteration counter (nIterateMC) before the packet loop -- otherwise every iteration replays the
 *                   histories of the first and the Monte Carlo noise freezes.  Within one epoch every source has its
 *                   own streams: a star's packets are keyed (packet index, iStar); the extra diffuse source of
 *                   mcb200_transport_diffuse by (packet index, grid, emitting cell), so the per-cell calls of
 *                   iteration_mod.f90:498-550 draw independent histories; resonance-line packets by (cell, index).
 *                   (This option and "seed" DO change results; the options below do not.)
 *   "trace"         1: keep per-packet fate records (mcb200_fetch_fates), print phase timings
 *   "wavefront"     -1 auto (wave-front pipeline for >= 2^17 packets per call), 0 persistent kernel, 1 wave front
 *   "order"         persistent kernel: -1 auto / 0 / 1 process packets in order of their first frequency bin
 *   "batch"         persistent kernel: lanes that must wait for a rare phase before it runs (12)
 *   "agg_steps"     persistent kernel: warp-aggregate the tallies of the first n crossings (0 = off)
 *   "blocks_per_sm" CTAs per SM of the transport kernels (0 = occupancy default)
 *   "step_budget"   wave front: cell crossings per flight per wave (96)
 *   "fly_batch"     wave front: idle lanes of a warp that trigger its store/claim pass (8)
 *   "tail"          wave front: alive packets below which the persistent kernel finishes the call (32768)
 *   "wave0_order"   wave front: 0 off, 1 auto, 2 always: emit wave 0 in first-frequency order into the FLY array
 *   "wave0_blocks"  CTAs per SM of that emission (4; 0 = full grid); "wave0_exact" 1: exact slots (slower)
 *   "async_pdfs"    1: mcb200_set_pdfs only enqueues the upload (see there)
 *   "sed_local"     1: per-rank SED counts, see mcb200_fetch_sed
 *   "tally_set", "parts", "part"   second tally set / sub-ranges of a rank's share, for overlapping the
 *                   exchange of one half with the transport of the other (PacketEngine.energyPacketDriverOverlapped)
 *   "exchange_dense" 1: mcb200_exchange all-reduces the escape counts densely whatever the list lengths
 *   "exchange_allreduce" 1: mcb200_exchange all-reduces the J planes and every rank folds all of them
 *   "pdf_slabs"     1: mcb200_set_pdfs uploads only this rank's 1/nranks slab of nu-planes of the (identical on
 *                   every rank) table over PCIe and all-gathers the slabs over NVLink (needs mcb200_comm_init)
 *   "exchange_p2p"  -1 auto / 0 / 1: fused peer-memory merge of the J tallies (mcb200_exchange_path)
 *   "exchange_pack" peer-memory merge, push variants: 1 (default) the partial sums travel packed -- the low 32 bits of
 *                   every element, the high 32 bits only of the 256-element blocks in which one is non-zero, nothing
 *                   of blocks that are zero altogether, a flag byte per block -- at most 4 instead of 8 bytes per
 *                   element on the links; 0 plain 64-bit pushes
 *   "exchange_push_blocks" CTAs per SM of the push kernels (default 4)
 *   "exchange_push" peer-memory merge: 1 (default) every rank pushes each peer's share of its partial sums into the
 *                   peer's receive buffer (device-to-device copies, posted writes), the owner sums locally; 2 the
 *                   same with a kernel that stores to all peers at once instead of copy after copy; 0 the owner
 *                   pulls the partial sums with peer loads inside the merge kernel
 *   "solo"          1: this rank acts as rank 0 of 1 until cleared (N-rank vs 1-rank check on one context)
 *   "defer_fold"    1: a single rank leaves its tallies pending after mcb200_transport, as a multi-rank run
 *                   does, until mcb200_reduce (lets one GPU walk the mcb200_exchange path) */
int mcb200_set_option(mcb200_ctx *ctx, const char *name, int64_t value);

/* unit-test hooks: run device primitives on n inputs (host arrays in/out). */
int mcb200_test_detmath(mcb200_ctx *ctx, int32_t which, const float *in, float *out, int64_t n);
int mcb200_test_uniforms(mcb200_ctx *ctx, uint64_t seed, uint64_t pid, uint32_t stream,
                         int32_t n, float *out);

/* measurement hook: device time of the push kernels of the peer-memory merge with all peer buffers local (one GPU);
 * mode 0 = 64-bit push kernel, 1 = packed push kernel; nElems = elements of the exchanged range, nranksSim = ranks. */
int mcb200_test_push_kernels(mcb200_ctx *ctx, int32_t mode, int32_t nranksSim, int64_t nElems, double *ms);

/* measurement hook (bench.py, SURVEY.md 8d "atomic roofline"): rate at which the device serves
 * the transport's per-crossing access pattern and nothing else -- mode bit 0: one 64-bit
 * reduction, bit 1: one 4-byte read, per iteration, at uniformly random addresses inside windows
 * of the given sizes (a nu-plane of JsteQ / opacity at 128^3 is 16.8 / 8.4 MB: L2 resident; use
 * GBs for the DRAM regime).  Returns iterations per second (best of 3 timed launches). */
int mcb200_test_access_peak(mcb200_ctx *ctx, int32_t mode, int64_t redWindowBytes, int64_t loadWindowBytes,
                            int64_t opsTotal, double *opsPerSecond);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_H */
