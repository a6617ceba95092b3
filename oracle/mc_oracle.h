/*
 * oracle/mc_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Data contract of the CPU oracle: a plain-C restatement of the energy-packet
 * transport of the reference (source/photon_mod.f90:26-2974) and of the per-cell
 * opacity assembly (source/ionization_mod.f90:349-484, iteration_mod.f90:166-227).
 *
 * PIN: the reference ships no golden vectors / tests for this path
 * and cannot be compiled here (no Fortran compiler).  It is *run* all the same: oracle/f90ref
 * translates photon_mod.f90 (and ionization_mod/emission_mod/update_mod/output_mod routines
 * either side of it) statement by statement into an executable module under oracle/_ref/, and
 * this oracle is checked against that -- bit for bit, every float32 tally element and every
 * packet history -- in tests/test_reference_pin.py, live where /root/reference exists and
 * through the golden vectors tests/golden/ref_*.npz everywhere.  Two restatement errors were
 * found and fixed that way (getNu2's "+1" rule on linePDF rows; updateCell's no-hit return).
 * Bound from outside, because the Fortran standard leaves them to the processor:
 * RANDOM_NUMBER (Philox here) and LOG/SIN/COS/ACOS/ATAN/EXP (detmath here); with the
 * platform's libm instead of detmath a few per cent of the packet histories differ by a
 * branch flipped at the last ulp (test_platform_libm_changes_few_histories).
 *
 * Array layouts follow the Fortran reference (column major, 1-based unless noted):
 *   active(nx,ny,nz)               -> active[(x-1) + nx*((y-1) + ny*(z-1))]
 *   T(0:nCells, 1:nbins)           -> T[(nu-1)*(nCells+1) + cell]
 *   escapedPackets(0:nCells,0:nbins,0:nAngleBins)
 *                                  -> E[cell + (nCells+1)*(nu + (nbins+1)*ang)]
 *   Tdust(0:nSpeciesMax,0:nSizes,0:nCells)
 *                                  -> Td[nS + (nSpeciesMax+1)*(ai + (nSizes+1)*cell)]
 *   linePackets/linePDF(0:nCells,1:nLines) like T with nLines
 */
#ifndef MC_ORACLE_H
#define MC_ORACLE_H

#include <stdint.h>

typedef struct OrGrid {
    int32_t nx, ny, nz, nCells, motherP;
    float geoCorrX, geoCorrY, geoCorrZ;
    float invLenUnit;              /* 1/unit of the fixed-point path-length tally (power of two) */
    const float *xAxis, *yAxis, *zAxis;
    const int32_t *active;
    const float *opacity, *scaOpac;
    const float *recPDF, *dustPDF, *linePDF, *totalLines;
    const float *Tdust;
    const int32_t *dustAbunIndex;
    /* outputs, faithful float32 sequential accumulation (may be NULL) */
    float *Jste, *Jdif, *escapedPackets, *linePackets;
    /* outputs, order-independent integer tallies (may be NULL):
     *   JsteQ/JdifQ: sum of llrintf(pathlength * invLenUnit)
     *   escapedQ/linePacketsQ: packet counts */
    int64_t *JsteQ, *JdifQ, *escapedQ, *linePacketsQ;
    /* extra packets per cell of the resonance-line transfer (0:nCells), may be NULL */
    const int32_t *resLinePackets;
    /* number of path segments added to every (cell, nu) element of Jste / Jdif (may be NULL): the
     * n in the per-element error bound of a float32 running sum against the fixed-point tally */
    int32_t *JsteN, *JdifN;
} OrGrid;

typedef struct OrParams {
    int32_t nGrids, nbins, nAngleBins, totAngleBinsTheta, totAngleBinsPhi, nLines, nStars;
    int32_t lgDust, lgGas, lgSymmetricXYZ, lgIsotropic, lgPlaneIonization, lgDebug,
            lgMultistars, lgMultiDustChemistry;
    int32_t nSpeciesMax, nSizes, nDustComp;
    float dTheta, dPhi, R_out, ionEdge1;
    const float *nuArray;          /* 1:nbins */
    const float *gSca;             /* 1:nbins */
    const int32_t *viewPointPtheta; /* 0:totAngleBinsTheta */
    const int32_t *viewPointPphi;   /* 0:totAngleBinsPhi */
    const float *viewPointTheta;   /* 0:nAngleBins */
    const float *viewPointPhi;     /* 0:nAngleBins */
    const float *starPosition;     /* [3*(i-1)+k], cm */
    const int32_t *starIndeces;    /* [4*(i-1)+k] = xP,yP,zP,grid (1-based values) */
    const float *deltaE;           /* 0:nStars */
    const float *inSpectrumProbDen;/* [(s)*nbins + (nu-1)], s=0..nStars */
    const int32_t *nSpeciesPart;   /* 1:nDustComp */
    const float *grainAbun;        /* (nDustComp,nSpeciesMax) column major */
    const int32_t *dustComPoint;   /* 1:nDustComp */
    const float *TdustSublime;     /* 1:nSpecies */
    int32_t *planeIonDistribution; /* (grid(1)%nx, grid(1)%nz) packets emitted per (x,z) of the y=0 face
                                      in plane-parallel mode (photon_mod.f90:643-646), may be NULL */
} OrParams;

typedef struct OrCounters {
    float Qphot;                   /* faithful fp32 sequential sum (photon_mod.f90:859-861) */
    float absInt, scaInt;          /* faithful fp32 counters (photon_mod.f90:1702,1720,1803) */
    int64_t nAbs, nSca;            /* exact event counts */
    int64_t trapped;               /* photon_mod.f90:126 */
    int64_t nLinePackets;          /* packets that ended as non-ionising line packets */
    int64_t nDropped;              /* packets dropped by safeLimit / wall return (no tally) */
    int64_t nSegments;             /* trips of the cell-crossing loop (photon_mod.f90:1194) */
    int64_t nFlights;              /* calls of pathSegment */
    int64_t nEscaped;              /* escape tallies */
    int64_t nEarlyEscaped;         /* escapes at energyPacketRun :370 (no dust, nu<ionEdge) */
} OrCounters;

/* per-packet fate record, 4 int32 per packet (optional):
 * [0] total segments, [1] generations (energyPacketRun calls), [2] last nuP,
 * [3] fate: 1 escaped, 2 line packet, 3 dropped, 4 trapped(recursion limit), 5 early escape */

#ifdef __cplusplus
extern "C" {
#endif

/* Transport n packets with global ids [firstId, firstId+n) of source iStar
 * (iStar>=1 stellar; iStar==0 extra diffuse source in cell cellLoc of grid gpLoc).
 * qphotCounts (nbins int64, may be NULL) receives the number of stellar emissions
 * per frequency bin with nu>1 Ryd.  Returns 0, or a negative code for the
 * reference's "print; stop" conditions. */
int oracle_transport(const OrParams *P, OrGrid *grids, int32_t iStar,
                     int64_t firstId, int64_t n, uint64_t seed,
                     int32_t gpLoc, const int32_t *cellLoc,
                     OrCounters *C, int64_t *qphotCounts, int32_t *fate);

/* Same packets, nThreads host threads, integer tallies only (atomic adds);
 * used as the CPU baseline in bench.py. */
int oracle_transport_mt(const OrParams *P, OrGrid *grids, int32_t iStar,
                        int64_t firstId, int64_t n, uint64_t seed,
                        int32_t nThreads, OrCounters *C, int64_t *qphotCounts);

/* Resonance-line packet transfer (photon_mod.f90:180-266): for every cell owned by `rank`
 * (mod(iCell-(rank+1),nranks)==0, iCell counting all cells of all grids in loop order)
 * resLinePackets(cell) "diffuse" packets start at the cell centre.  The Philox stream of
 * packet j (global enumeration order over grids/cells/iPhot) is keyed by 2^40+j, so the
 * result does not depend on nranks.  Returns the number of packets run via *nRun. */
int oracle_transport_reslines(const OrParams *P, OrGrid *grids, int32_t iStar, uint64_t seed,
                              int32_t rank, int32_t nranks, OrCounters *C, int64_t *nRun);

/* Opacity assembly for every active cell of one grid: restatement of ionizationDriver's
 * density computation (ionization_mod.f90:65-80), addOpacity/putOpacity/inOpacity
 * (:349-484; the free-free term only ever exists in bin 1, :369-393, and is passed in as
 * ff1 because it depends on the host's BoltGaunt cache) and of the dust loop of
 * iteration_mod.f90:166-227.  Layouts: ionDen(0:nCells,nElementsUsed,nstages),
 * elemAbun(nAbComp,30), elementP(30,30,7,3), nShells(30,30), all Fortran column major;
 * abIndex/Hden/ff1/Ndust/dustAbunIndex are (0:nCells). */
typedef struct OrOpacityIn {
    int32_t nCells, nbins, nstages, nElementsUsed, nAbComp;
    const int32_t *lgElementOn;   /* 1:30 */
    const int32_t *elementXref;   /* 1:30 */
    const float *ionDen, *elemAbun, *Hden, *ff1;
    const int32_t *abIndex;
    const float *xSecArray;
    int32_t HlevXSecP1, HlevNuP1, HeISingXSecP1, HeIlevNuP1, HeIIXSecP1, HeIIlevNuP1;
    const int32_t *elementP, *nShells;
    /* dust (lgDust): */
    int32_t lgDust, lgMultiDustChemistry, nSpeciesMax, nSizes, nDustComp, nSpeciesTot;
    const int32_t *nSpeciesPart, *dustComPoint, *dustAbunIndex, *dustScaXsecP, *dustAbsXsecP;
    const float *grainAbun, *grainWeight, *TdustSublime, *Tdust, *Ndust;
} OrOpacityIn;
void oracle_opacity(const OrOpacityIn *in, float *opacity, float *scaOpac, float *absOpac);

/* Dust-only closure of the Lucy iteration ("next" rows of SURVEY.md 8f):
 *  oracle_dust_pdf:    setDustPDF (emission_mod.f90:1313-1387, non-quantum-heating branch)
 *                      for every active cell: dustPDF(cell,:) = normalised running sum of
 *                      Cabs * B_nu(Tdust) * widFlx * grainWeight * grainAbun
 *  oracle_dust_update: the dust-only branch of updateCell (update_mod.f90:308-334) with
 *                      getDustT (:1836-1945): Tdust from sum_nu Cabs*J/pi by inverse lookup in
 *                      dustEmIntegral, weighted means, lgConverged from |dT/T| <= XHILimit.
 *                      A cell with no Jste(cell,:) > 0 (no Jdif > 0 either in debug mode) is
 *                      left untouched (updateCell's lgHit test, :104-149): its Tdust and its
 *                      lgConverged entry keep the values they came in with.
 *                      Jste is the array the host holds at that point, i.e. after the scaling
 *                      of iteration_mod.f90:705-724.  getDustT indexes dustAbsXsecP and
 *                      dustEmIntegral with the component-local species number (sic).  A cell
 *                      whose absorption integral lies below dustEmIntegral(.,.,1) makes the
 *                      reference read dustEmIntegral(.,.,0) (out of bounds, :1915-1919) unless
 *                      lgTalk is set; here it gets the lgTalk value, 1 K.
 * getFlux is continuum_mod.f90:359-416 ('blackbody') with detmath's exp. */
typedef struct OrDustIn {
    int32_t nCells, nbins, nSpeciesMax, nSizes, nDustComp, nSpeciesTot, nTemps;
    int32_t lgMultiDustChemistry, lgDebug;
    const float *nuArray, *widFlx, *xSecArray;
    const int32_t *dustAbsXsecP;      /* (nSpeciesTot, nSizes) */
    const int32_t *nSpeciesPart, *dustComPoint, *dustAbunIndex;
    const float *grainAbun, *grainWeight, *TdustSublime;
    const float *dustEmIntegral;      /* (nSpeciesTot, nSizes, nTemps) */
} OrDustIn;
void oracle_dust_pdf(const OrDustIn *in, const float *Tdust, float *dustPDF);
void oracle_dust_update(const OrDustIn *in, const float *Jste, const float *Jdif, float XHILimit,
                        float *Tdust, int32_t *lgConverged);
float oracle_get_flux(float energy, float temperature);

/* Photo-rate pre-integration for updateCell: per band b (one ion's outer
 * shell: 1-based offset off into xSecArray, frequency range low..high) and per cell
 *   nPhoto(cell,b) = 1e-20 + sum_{j=low..high, J(cell,j)>0} J*x/(hcRyd*nu(j)), x<1e-35 -> 0
 *                                                     (update_mod.f90:170-262)
 *   heat(cell,b)   = sum_{j=low.., stop at the first x<1e-35, J>0} x*J*(nu(j)-nu(low))/nu(j)
 *                                                     (thermBalance, update_mod.f90:1160-1214)
 * J is the host-scaled Jste (or Jdif).  Outputs (0:nCells, nBands), cell index fastest. */
void oracle_photo_integrals(int32_t nCells, int32_t nbins, int32_t nBands, const int32_t *off,
                            const int32_t *low, const int32_t *high, const float *xSecArray,
                            const float *nuArray, const float *J, float *nPhoto, float *heat);

/* unit-test hooks */
void oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                   uint32_t k0, uint32_t k1, uint32_t *out4);
void oracle_uniforms(uint64_t seed, uint64_t pid, uint32_t stream, int32_t n, float *out);
int32_t oracle_locate(const float *xa, int32_t n, float x);
int32_t oracle_getnu2(const float *probDen, int64_t stride, int32_t nbins,
                      uint64_t seed, uint64_t pid, uint32_t stream);
void oracle_random_unit_vector(uint64_t seed, uint64_t pid, uint32_t stream, float *out3);
int32_t oracle_hg(float g, const float *vin, uint64_t seed, uint64_t pid, uint32_t stream,
                  float *vout);
void oracle_detmath(int32_t which, const float *in, float *out, int64_t n);
void oracle_detmath_d(int32_t which, const double *in, double *out, int64_t n);
int32_t oracle_escape_bins(const OrParams *P, const float *dir, int32_t *idirT, int32_t *idirP);
float oracle_cell_volume(const OrParams *P, const OrGrid *g, int32_t xP, int32_t yP, int32_t zP);

#ifdef __cplusplus
}
#endif
#endif
