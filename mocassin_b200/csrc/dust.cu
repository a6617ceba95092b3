// dust.cu -- dust-only closure of the Lucy iteration on the device (SURVEY.md 8f "next" rows).
//
//  K5 dust_update_kernel : getDustT (update_mod.f90:1836-1945) + the dust-only branch of
//                          updateCell (:308-334) for every cell, reading the folded Jste that
//                          is already resident after mcb200_transport / mcb200_reduce.
//  K8 photo_kernel       : photo-ionisation and heating integrals of updateCell / thermBalance
//                          (update_mod.f90:170-262, :1160-1214) per (cell, ion band).
//  K6 dust_pdf_kernel    : setDustPDF (emission_mod.f90:1313-1387, non-quantum-heating branch)
//                          written straight into the nu-contiguous PDF rows the transport samples.
//
// Both are float32 restatements with the reference's operation order (sequential sums over
// nu, products left to right), -fmad=false, and the deterministic exp of detmath.cuh, so the
// results are bit-identical to oracle_dust_update / oracle_dust_pdf.
#include "dust.h"
#include "detmath.cuh"
#include "locate.cuh"

namespace mcb {

// getFlux, continuum_mod.f90:359-416, cShape 'blackbody'
__device__ __forceinline__ float get_flux(float energy, float temperature)
{
    const float hPlanck = 6.6262e-27f, hcRyd_k = 157893.94f;
    float constant = 0.5250229f / hPlanck;
    if (hcRyd_k * energy / temperature > 86.f) {
        float pre = constant * energy * energy * energy;
        return (float)((double)pre * dm_exp_d((double)(-hcRyd_k * energy / temperature)));
    }
    float denominator = dm_expf(hcRyd_k * energy / temperature) - 1.f;
    if (denominator <= 0.f) return 3.32154e-6f * energy * energy * temperature / hPlanck;
    return constant * energy * energy * energy / denominator;
}

__device__ __forceinline__ size_t tdust_at(const DustArgs &A, int nS, int ai, int cell)
{
    return (size_t)nS + (size_t)(A.nSpeciesMax + 1) * ((size_t)ai + (size_t)(A.nSizes + 1) * (size_t)cell);
}

// ---------------------------------------------------------------------------------------
// K5: one thread per cell.  The absorption integrals of kPairs (species,size) pairs are
// accumulated together so that the Jste row of the cell is read ceil(nPairs/kPairs) times
// (coalesced across the cells of a warp; the xSec reads are warp-uniform broadcasts).
// HBM-bound: algorithmic bytes per cell = 4*nbins (Jste) + 4*(nSpeciesMax+1)*(nSizes+1) (Tdust).
// ---------------------------------------------------------------------------------------
constexpr int kPairs = 8;

__global__ void __launch_bounds__(128) dust_update_kernel(const DustArgs A)
{
    int cell = blockIdx.x * blockDim.x + threadIdx.x + 1;
    bool live = cell <= A.nCells;
    int conv = 0;
    if (live) {
        const size_t nR = (size_t)A.nCells + 1;
        // updateCell returns first thing for a cell no packet crossed (update_mod.f90:104-149):
        // Tdust and the sublimation flag keep their previous values, lgConverged the 0 that
        // iterateMC gave it at the start of the iteration (iteration_mod.f90:87)
        bool hit = false;
        for (int i = 1; i <= A.nb && !hit; ++i) {
            size_t o = (size_t)(i - 1) * nR + cell;
            float J = A.Jste[o] * 1.e-9f;
            if (A.sym) J = J / 8.f;
            hit = J > 0.f;
            if (A.lgDebug && !hit) {
                float Jd = A.Jdif[o] * 1.e-9f;
                if (A.sym) Jd = Jd / 8.f;
                hit = Jd > 0.f;
            }
        }
        if (!hit) {
            conv = A.lgConverged[cell];
            live = false;
        }
    }
    if (live) {
        const size_t nR = (size_t)A.nCells + 1;
        const float Pi = 3.141592654f;
        int nspU = A.multiChem ? __ldg(&A.dustAbunIndex[cell]) : 1;
        bool comp = nspU >= 1 && nspU <= A.nDustComp;
        int nSp = comp ? __ldg(&A.nSpeciesPart[nspU - 1]) : 0;
        float XOldHI = A.Tdust[tdust_at(A, 0, 0, cell)];
        for (int nS = 0; nS <= A.nSpeciesMax; ++nS)
            for (int ai = 0; ai <= A.nSizes; ++ai) A.Tdust[tdust_at(A, nS, ai, cell)] = 0.f;
        const int nPairs = nSp * A.nSizes;          // pair p = (nS-1)*nSizes + (ai-1), nS-major
        float sumS = 0.f, tot = 0.f;
        for (int p0 = 0; p0 < nPairs; p0 += kPairs) {
            float acc[kPairs];
            int ap[kPairs];
#pragma unroll
            for (int k = 0; k < kPairs; ++k) {
                acc[k] = 0.f;
                int p = p0 + k < nPairs ? p0 + k : nPairs - 1;
                int nS = p / A.nSizes + 1, ai = p % A.nSizes + 1;
                // sic: getDustT indexes dustAbsXsecP with the component-local species number
                ap[k] = __ldg(&A.absP[(size_t)(nS - 1) + (size_t)A.nSpeciesTot * (size_t)(ai - 1)]);
            }
            for (int i = 1; i <= A.nb; ++i) {
                size_t o = (size_t)(i - 1) * nR + cell;
                float J = A.Jste[o] * 1.e-9f;                      // iteration_mod.f90:706
                if (A.sym) J = J / 8.f;                            // :718
                float rf;
                if (A.lgDebug) {
                    float Jd = A.Jdif[o] * 1.e-9f;
                    if (A.sym) Jd = Jd / 8.f;
                    rf = (J + Jd) / Pi;
                } else rf = J / Pi;
#pragma unroll
                for (int k = 0; k < kPairs; ++k) acc[k] = acc[k] + __ldg(&A.xSec[ap[k] + i - 2]) * rf;
            }
#pragma unroll
            for (int k = 0; k < kPairs; ++k) {
                int p = p0 + k;
                if (p >= nPairs) break;
                int nS = p / A.nSizes + 1, ai = p % A.nSizes + 1;
                const float *row = A.emT + ((size_t)(nS - 1) + (size_t)A.nSpeciesTot * (size_t)(ai - 1)) * (size_t)A.nTemps;
                float dustAbsIntegral = acc[k];
                int iT = locate_axis(row, A.nTemps, dustAbsIntegral);
                float T;
                if (iT <= 0) T = 1.f;
                else if (iT >= A.nTemps) T = (float)A.nTemps;
                else T = (float)iT + (dustAbsIntegral - __ldg(&row[iT - 1])) * ((float)(iT + 1) - (float)iT) /
                                         (__ldg(&row[iT]) - __ldg(&row[iT - 1]));
                A.Tdust[tdust_at(A, nS, ai, cell)] = T;
                sumS = sumS + T * __ldg(&A.grainWeight[ai - 1]);
                if (ai == A.nSizes) {
                    A.Tdust[tdust_at(A, nS, 0, cell)] = sumS;
                    tot = tot + sumS * __ldg(&A.grainAbun[(size_t)(nspU - 1) + (size_t)A.nDustComp * (size_t)(nS - 1)]);
                    sumS = 0.f;
                }
            }
        }
        A.Tdust[tdust_at(A, 0, 0, cell)] = tot;
        float deltaXHI = (tot - XOldHI) / XOldHI;
        conv = fabsf(deltaXHI) <= A.XHILimit ? 1 : 0;
        A.lgConverged[cell] = conv;
        // sublimation test of photon_mod.f90:1722-1748 for the next transport
        unsigned char can = 0;
        if (comp) {
            int dcp = __ldg(&A.dustComPoint[nspU - 1]);
            for (int nS = 1; nS <= nSp; ++nS) {
                float ab = __ldg(&A.grainAbun[(size_t)(nspU - 1) + (size_t)A.nDustComp * (size_t)(nS - 1)]);
                float Td = A.Tdust[tdust_at(A, nS, 0, cell)];
                if (ab > 0.f && Td < __ldg(&A.TdustSublime[dcp - 1 + nS - 1])) { can = 1; break; }
            }
        }
        A.canScatter[cell] = can;
    }
    unsigned m = __ballot_sync(0xffffffffu, conv);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(A.nConv, (unsigned long long)__popc(m));
}

// ---------------------------------------------------------------------------------------
// K6: one warp per cell.  Lanes stride over nu for the (species,size) sums, lane 0 runs the
// sequential prefix sum in shared memory (float addition order is the reference's), lanes
// normalise and store the row.  Compute-bound on the double-precision exp of getFlux:
// nbins * nSpeciesPart * nSizes evaluations per cell.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dust_pdf_kernel(const DustArgs A)
{
    extern __shared__ float rows[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float *row = rows + (size_t)w * A.nb;
    for (int cell = blockIdx.x * wpb + w; cell <= A.nCells; cell += gridDim.x * wpb) {
        float *out = A.pdfT + (size_t)cell * A.nb;
        if (cell == 0) {
            for (int i = lane; i < A.nb; i += 32) out[i] = 0.f;
            continue;
        }
        int nspE = A.multiChem ? __ldg(&A.dustAbunIndex[cell]) : 1;
        if (nspE < 1 || nspE > A.nDustComp) {
            for (int i = lane; i < A.nb; i += 32) out[i] = 0.f;
            continue;
        }
        int dcp = __ldg(&A.dustComPoint[nspE - 1]);
        int nSp = __ldg(&A.nSpeciesPart[nspE - 1]);
        for (int i = lane; i < A.nb; i += 32) row[i] = 0.f;
        for (int n = 1; n <= nSp; ++n) {
            float sub = __ldg(&A.TdustSublime[dcp - 1 + n - 1]);
            float ga = __ldg(&A.grainAbun[(size_t)(nspE - 1) + (size_t)A.nDustComp * (size_t)(n - 1)]);
            for (int ai = 1; ai <= A.nSizes; ++ai) {
                float treal = A.Tdust[tdust_at(A, n, ai, cell)];
                if (!(treal > 0.f && treal < sub)) continue;
                int ap = __ldg(&A.absP[(size_t)(n + dcp - 1 - 1) + (size_t)A.nSpeciesTot * (size_t)(ai - 1)]);
                float gw = __ldg(&A.grainWeight[ai - 1]);
                for (int i = lane; i < A.nb; i += 32) {
                    float bb = get_flux(__ldg(&A.nuArray[i]), treal);
                    row[i] = row[i] + __ldg(&A.xSec[ap + i - 1]) * bb * __ldg(&A.widFlx[i]) * gw * ga;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            float s = row[0];
            for (int i = 1; i < A.nb; ++i) { s = s + row[i]; row[i] = s; }
        }
        __syncwarp();
        float last = row[A.nb - 1];
        for (int i = lane; i < A.nb; i += 32) out[i] = (i == A.nb - 1) ? 1.f : row[i] / last;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// K8: nPhoto(cell,b) = 1e-20 + sum_j J x /(hcRyd nu_j);  heat(cell,b) = sum_j x J (nu_j - nu_IP)/nu_j.
// One CTA = 128 cells; nu is the outer loop so every Jste element is read once (coalesced
// over cells); the bands covering a bin come from a CSR in band order, each band's running
// sums live in shared memory, so each sum is formed in the reference's frequency order.
// HBM-bound: 4 B per (cell, nu) read + 8 B per (cell, band) written.
// ---------------------------------------------------------------------------------------
constexpr int kPhotoTile = 128;

__global__ void __launch_bounds__(kPhotoTile) photo_kernel(const PhotoArgs A)
{
    extern __shared__ float acc[];               // [2][nB][kPhotoTile]
    const float hcRyd = 2.1799153e-11f;
    const int t = threadIdx.x;
    const int cell = blockIdx.x * kPhotoTile + t;        // 0..nCells
    const bool live = cell <= A.nCells;
    const size_t nR = (size_t)A.nCells + 1;
    float *accR = acc, *accH = acc + (size_t)A.nB * kPhotoTile;
    for (int b = 0; b < A.nB; ++b) { accR[b * kPhotoTile + t] = 1.e-20f; accH[b * kPhotoTile + t] = 0.f; }
    for (int j = 1; j <= A.nb; ++j) {
        int k0 = __ldg(&A.nuStart[j - 1]), k1 = __ldg(&A.nuStart[j]);
        if (k0 == k1) continue;
        float J = 0.f;
        if (live) {
            J = A.J[(size_t)(j - 1) * nR + cell] * 1.e-9f;     // iteration_mod.f90:706
            if (A.sym) J = J / 8.f;                              // :718
        }
        if (!(J > 0.f)) continue;
        float nu = __ldg(&A.nuArray[j - 1]);
        for (int k = k0; k < k1; ++k) {
            int b = __ldg(&A.nuBand[k]);
            if (b < A.b0 || b >= A.b0 + A.nB) continue;
            int lo = __ldg(&A.low[b]);
            float x = __ldg(&A.xSec[__ldg(&A.off[b]) + (j - lo) - 1]);
            if (x < 1.e-35f) x = 0.f;
            int lb = b - A.b0;
            accR[lb * kPhotoTile + t] = accR[lb * kPhotoTile + t] + J * x / (hcRyd * nu);
            if (j <= __ldg(&A.heatHigh[b]))
                accH[lb * kPhotoTile + t] = accH[lb * kPhotoTile + t] + x * J * (nu - __ldg(&A.nuArray[lo - 1])) / nu;
        }
    }
    if (live)
        for (int b = 0; b < A.nB; ++b) {
            if (A.nPhoto) A.nPhoto[(size_t)(A.b0 + b) * nR + cell] = accR[b * kPhotoTile + t];
            if (A.heat) A.heat[(size_t)(A.b0 + b) * nR + cell] = accH[b * kPhotoTile + t];
        }
}

cudaError_t launch_photo(const PhotoArgs &A, cudaStream_t s)
{
    size_t smem = (size_t)2 * A.nB * kPhotoTile * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(photo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    unsigned blocks = (unsigned)((A.nCells + 1 + kPhotoTile - 1) / kPhotoTile);
    photo_kernel<<<blocks, kPhotoTile, smem, s>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_dust_update(const DustArgs &A, cudaStream_t s)
{
    unsigned blocks = (unsigned)((A.nCells + 127) / 128);
    dust_update_kernel<<<blocks, 128, 0, s>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_dust_pdf(const DustArgs &A, int numSMs, cudaStream_t s)
{
    const int threads = 256;
    size_t smem = (size_t)(threads / 32) * A.nb * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(dust_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int perSM = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, dust_pdf_kernel, threads, smem);
    if (e != cudaSuccess) return e;
    if (perSM < 1) perSM = 1;
    long long want = ((long long)A.nCells + 1 + threads / 32 - 1) / (threads / 32);
    long long cap = (long long)numSMs * perSM;
    unsigned blocks = (unsigned)(want < cap ? want : cap);
    if (blocks < 1) blocks = 1;
    dust_pdf_kernel<<<blocks, threads, smem, s>>>(A);
    return cudaGetLastError();
}

}  // namespace mcb
