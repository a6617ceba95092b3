/* host_example.c -- the C ABI of libmocassin_b200.so used from plain C, with no Python and no
 * torch in the process: what the Fortran host does through fortran/mcb200_mod.f90, spelled out.
 *
 * A gas-only uniform cube (n^3 cells, grey opacity kappa, star at the centre, non-symmetric):
 * every packet is absorbed and re-emitted until it leaves, so the energy that escapes must
 * equal the luminosity, and in the optically thin limit the path-length estimator integrates to
 * L x <chord>:  sum_cells Jste*dV = deltaE * (total path length).
 *
 *   gcc -std=c99 -Iinclude examples/host_example.c mocassin_b200/libmocassin_b200.so \
 *       -Wl,-rpath,$PWD/mocassin_b200 -lm -o /tmp/host_example && /tmp/host_example
 *
 * Exit code 0 = checks passed, 2 = no CUDA device (the library has no CPU fallback). */
#include "mcb200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(call)                                                                        \
    do {                                                                                   \
        int rc__ = (call);                                                                 \
        if (rc__ != MCB200_OK) {                                                           \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, mcb200_last_error(ctx));        \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

int main(void)
{
    enum { N = 15, NB = 40 };
    const int nCells = N * N * N, nRows = nCells + 1;
    const float edge = 1.0e17f, kappa = 2.0e-18f;       /* tau across the half box = 0.2 */
    const int64_t nPackets = 200000;
    const float Lstar = 1.0f, deltaE = Lstar / (float)nPackets;
    mcb200_ctx *ctx = NULL;
    int rc = mcb200_create(&ctx, 0, 0, 1, 12345u);
    if (rc == MCB200_ENODEV) { printf("no CUDA device: mcb200_create -> MCB200_ENODEV (no CPU fallback)\n"); return 2; }
    if (rc != MCB200_OK) { fprintf(stderr, "mcb200_create -> %d\n", rc); return 1; }

    /* ---- once: configuration, geometry, spectra, star (mocassin.f90:31-151) ---- */
    mcb200_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.nGrids = 1; cfg.nbins = NB; cfg.nStars = 1; cfg.nAngleBins = 0;
    cfg.totAngleBinsTheta = 10; cfg.totAngleBinsPhi = 20;
    cfg.lgGas = 1;                                       /* gas only: no dust tables needed */
    cfg.nDustComp = 1;
    cfg.dTheta = 3.141592654f / 10.f; cfg.dPhi = 2.f * 3.141592654f / 20.f;
    cfg.R_out = 0.f; cfg.ionEdge1 = 1.0e-9f;            /* every bin ionises: no early escapes */
    CHECK(mcb200_set_config(ctx, &cfg));

    float axis[N];
    for (int i = 0; i < N; ++i) axis[i] = (2.f * (float)i / (float)(N - 1) - 1.f) * edge;   /* fillGrid, :585 */
    int32_t *active = malloc(sizeof(int32_t) * nCells);
    for (int x = 0; x < N; ++x)                          /* cell ids: x outermost, z innermost (:1227-1262) */
        for (int y = 0; y < N; ++y)
            for (int z = 0; z < N; ++z) active[x + N * (y + N * z)] = 1 + z + N * (y + N * x);
    CHECK(mcb200_set_grid(ctx, 1, N, N, N, nCells, 0, axis, axis, axis, active));

    float nu[NB], cdf[2 * NB];
    for (int i = 0; i < NB; ++i) {
        nu[i] = 1.f + 0.1f * (float)i;
        cdf[2 * i + 0] = 0.f;                            /* inSpectrumProbDen(0:nStars, nbins): row 0 = diffuse source */
        cdf[2 * i + 1] = (float)(i + 1) / (float)NB;
    }
    CHECK(mcb200_set_spectra(ctx, nu, NULL, cdf));
    const float starPos[3] = {0.f, 0.f, 0.f};
    const int32_t starIdx[4] = {(N + 1) / 2, (N + 1) / 2, (N + 1) / 2, 1};
    CHECK(mcb200_set_stars(ctx, starPos, starIdx));

    /* ---- every iteration: opacities, re-emission CDFs, transport, estimators ---- */
    float *opacity = calloc((size_t)nRows * NB, sizeof(float));
    float *recPDF = calloc((size_t)nRows * NB, sizeof(float));
    float *totalLines = calloc((size_t)nRows, sizeof(float));
    for (int f = 0; f < NB; ++f)
        for (int c = 1; c < nRows; ++c) {
            opacity[c + (size_t)nRows * f] = kappa;
            recPDF[c + (size_t)nRows * f] = (float)(f + 1) / (float)NB;
        }
    CHECK(mcb200_set_opacity(ctx, 1, opacity, NULL));
    CHECK(mcb200_set_pdfs(ctx, 1, recPDF, NULL, totalLines, NULL));
    CHECK(mcb200_zero_estimators(ctx));
    mcb200_counters cnt;
    CHECK(mcb200_transport(ctx, 1, nPackets, deltaE, &cnt));        /* single rank: folded on return */

    float *Jste = malloc(sizeof(float) * (size_t)nRows * NB);
    float *esc = malloc(sizeof(float) * (size_t)nRows * (NB + 1));
    CHECK(mcb200_fetch_estimators(ctx, 1, Jste, esc, NULL, NULL));

    double escaped = 0.0, jdv = 0.0;
    for (size_t i = 0; i < (size_t)nRows * (NB + 1); ++i) escaped += esc[i];
    for (int x = 0; x < N; ++x)
        for (int y = 0; y < N; ++y)
            for (int z = 0; z < N; ++z) {
                /* cell widths as getVolume (grid_mod.f90:2876-2965), in 1e15 cm */
                double w[3];
                const int p[3] = {x, y, z};
                for (int k = 0; k < 3; ++k)
                    w[k] = (p[k] > 0 && p[k] < N - 1 ? fabs(axis[p[k] + 1] - axis[p[k] - 1]) / 2.0
                            : p[k] == 0 ? fabs(axis[1] - axis[0]) : fabs(axis[N - 1] - axis[N - 2])) / 1.0e15;
                const int c = active[x + N * (y + N * z)];
                for (int f = 0; f < NB; ++f) jdv += (double)Jste[c + (size_t)nRows * f] * w[0] * w[1] * w[2];
            }
    /* Jste = path length * deltaE / dV with dV in 1e45 cm^3 (photon_mod.f90:1563-1574), so
     * sum_cells Jste*dV = deltaE * (total path length) = L * (mean path per packet)               */
    printf("packets %lld  absorptions %lld  crossings %lld  kernel %.3f ms\n", (long long)cnt.nPackets, (long long)cnt.nAbs,
           (long long)cnt.nSegments, cnt.kernel_ms);
    printf("escaped energy %.6f of L = %.6f\n", escaped, (double)Lstar);
    const double mean_path = jdv / (double)Lstar;                   /* cm */
    printf("sum Jste*dV -> mean path per packet %.4e cm (box half edge %.2e cm)\n", mean_path, (double)edge);
    /* tau = 0.2 across the half box: about a quarter of the packets are absorbed and re-emitted once; the CPU
     * oracle gives 1.33 half edges for this set-up */
    int ok = fabs(escaped - Lstar) < 1e-3 * Lstar && cnt.nEscaped == cnt.nPackets && mean_path > 1.2 * edge &&
             mean_path < 1.5 * edge;
    CHECK(mcb200_destroy(ctx));
    free(active); free(opacity); free(recPDF); free(totalLines); free(Jste); free(esc);
    printf(ok ? "OK\n" : "FAILED\n");
    return ok ? 0 : 1;
}
