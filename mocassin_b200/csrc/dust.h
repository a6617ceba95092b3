// dust.h -- arguments of the dust-closure kernels (dust.cu), filled by capi.cu.
#pragma once
#include <cuda_runtime.h>

namespace mcb {

struct DustArgs {
    int nCells, nb, nSpeciesMax, nSizes, nDustComp, nSpeciesTot, nTemps;
    int multiChem, lgDebug, sym;
    const float *nuArray, *widFlx, *xSec;
    const int *absP;                 // dustAbsXsecP(nSpeciesTot, nSizes), reference layout
    const int *nSpeciesPart, *dustComPoint, *dustAbunIndex;
    const float *grainAbun, *grainWeight, *TdustSublime;
    const float *emT;                // dustEmIntegral re-laid as [(nS-1) + nSpeciesTot*(ai-1)][T]
    float *Tdust;                    // (0:nSpeciesMax, 0:nSizes, 0:nCells)
    const float *Jste, *Jdif;        // folded raw sums, (0:nCells, nbins)
    float *pdfT;                     // (0:nCells)[nbins]
    int *lgConverged;                // (0:nCells)
    unsigned char *canScatter;       // (0:nCells)
    float XHILimit;
    unsigned long long *nConv;
};

// photo-rate pre-integration (K8, dust.cu): bands [b0, b0+nB) of a CSR over nu
struct PhotoArgs {
    int nCells, nb, sym;
    int b0, nB;                      // bands handled by this launch
    const int *nuStart, *nuBand;     // CSR: bands covering bin nu are nuBand[nuStart[nu-1] .. nuStart[nu])
    const int *off, *low, *heatHigh; // per band: 1-based xSecArray offset, first bin, last bin of the heating sum
    const float *xSec, *nuArray;
    const float *J;                  // folded raw sums (0:nCells, nbins)
    float *nPhoto, *heat;            // (0:nCells, nBands)
};
cudaError_t launch_photo(const PhotoArgs &A, cudaStream_t s);

cudaError_t launch_dust_update(const DustArgs &A, cudaStream_t s);
cudaError_t launch_dust_pdf(const DustArgs &A, int numSMs, cudaStream_t s);

}  // namespace mcb
