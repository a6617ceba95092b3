"""Seeded parity cases shared by the CPU (oracle-only) and GPU (CUDA vs oracle) tests."""
from mocassin_b200 import workloads as W

CASES = {
    # name: (builder, kwargs, packets)
    "hii_sym_gas": (W.hii_region, dict(), 40000),
    "hii_sym_gas_debug": (W.hii_region, dict(debug=True, seed=9), 20000),
    "dust_shell_hg": (W.dust_shell, dict(tauV=3.0), 8000),
    "dust_shell_iso": (W.dust_shell, dict(tauV=10.0, isotropic=True, n=12), 3000),
    "multigrid_sym": (W.multigrid, dict(), 12000),
    "multigrid_nonsym": (W.multigrid, dict(symmetric=False, n=15), 12000),
    "cube_uniform_gas": (W.synthetic_cube, dict(n=20, nbins=120, clumpy=False, dust=False, nPhotons=10**6), 20000),
    "cube_clumpy_gasdust": (W.synthetic_cube, dict(n=24, nbins=150, clumpy=True, dust=True, nPhotons=10**6), 8000),
    "viewing_angles": (W.viewing_angles, dict(), 20000),
    "viewing_angles_phifree": (W.viewing_angles, dict(phi_free=True), 10000),
    "plane_slab_gasdust": (W.plane_slab, dict(Hden=30.0), 12000),
    "plane_slab_gas": (W.plane_slab, dict(dust=False, Hden=30.0), 12000),
}


def make(name):
    fn, kw, n = CASES[name]
    return fn(**kw), n
