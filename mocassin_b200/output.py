"""Host-side output of the escaped-packet spectrum: the part of ``writeSED``
(output_mod.f90:2508-2719) that follows the reduction over cells.  The reduction itself,
``SED(freq, imu) = sum_cells escapedPackets(i, freq, imu)``, is ``mcb200_fetch_sed`` (device);
everything here is O(nbins) float32 arithmetic in the reference's order."""
from __future__ import annotations

import numpy as np

from .model import F32, Model

C_LIGHT = F32(2.9979250e10)      # constants_mod.f90: c [cm/s]
FR1RYD = F32(3.28984e15)         # constants_mod.f90:13
PI = F32(3.141592654)


def sed_from_raw(model: Model, widFlx: np.ndarray, raw: np.ndarray):
    """raw: (nbins, nAngleBins+1) sums of escapedPackets as fetched from the library (before the
    host's /8).  Applies iteration_mod.f90:719 (/8 for symmetricXYZ) and writeSED's scaling
    (:2626-2660).  Returns (SED [Jy pc^2], totalE [1e36 erg/s])."""
    at = model.angle_tables()
    sed = np.array(raw, dtype=F32, order="F", copy=True)
    if model.lgSymmetricXYZ:
        sed = (sed / F32(8.0)).astype(F32)                       # iteration_mod.f90:719
        sed[:, 0] = sed[:, 0] * F32(8.0)                         # output_mod.f90:2629
        sed[:, 1:] = sed[:, 1:] * F32(4.0)                       # :2631
    totalE = F32(0.0)
    for f in range(model.nbins):
        totalE = F32(totalE + sed[f, 0])                         # :2635
    wid = np.asarray(widFlx, dtype=F32)
    sed[:, 0] = sed[:, 0] / F32(F32(F32(F32(4.0) * PI) * F32(3.08)) * F32(3.08))       # :2639
    sed[:, 0] = F32(1.0e23) * sed[:, 0] / (F32(3.2898e15) * wid)                       # :2642
    dTheta, dPhi = F32(at["dTheta"]), F32(at["dPhi"])
    for imu in range(1, model.nAngleBins + 1):
        theta1 = F32(int(F32(model.viewPointTheta[imu]) / dTheta)) * dTheta            # :2649
        theta2 = F32(theta1 + dTheta)
        solid = F32(F32(F32(dPhi * F32(3.08)) * F32(3.08)) * F32(abs(F32(np.cos(theta1)) - F32(np.cos(theta2)))))
        sed[:, imu] = sed[:, imu] / solid                                              # :2656
        sed[:, imu] = F32(1.0e23) * sed[:, imu] / (F32(3.2898e15) * wid)               # :2658
    return sed, float(totalE)


def write_sed(path: str, model: Model, widFlx: np.ndarray, raw: np.ndarray) -> float:
    """output/SED.out in the reference's layout (:2551-2554, :2660, :2692-2708); list-directed
    number formatting is the compiler's in the reference, here %.7E."""
    at = model.angle_tables()
    sed, totalE = sed_from_raw(model, widFlx, raw)
    nu = np.asarray(model.nuArray, dtype=F32)
    with open(path, "w") as fh:
        fh.write(" Spectral energy distribution at the surface of the nebula: \n")
        vp = "".join(f" {float(model.viewPointTheta[i]):.7E} {float(model.viewPointPhi[i]):.7E}  , "
                     for i in range(1, model.nAngleBins + 1))
        fh.write("   viewPoints = " + vp + "\n")
        fh.write("    nu [Ryd]        lambda [um]         F(nu)*D^2            \n")
        fh.write("                                        [Jy * pc^2]              \n")
        for f in range(model.nbins):
            lam_um = F32(C_LIGHT / F32(nu[f] * FR1RYD)) * F32(1.0e4)
            cols = " ".join(f"{float(sed[f, a]):.7E}" for a in range(model.nAngleBins + 1))
            fh.write(f" {float(nu[f]):.7E} {float(lam_um):.7E} {cols}\n")
        fh.write(" \n")
        fh.write(f" Total energy radiated out of the nebula [e36 erg/s]: {totalE:.7E}\n")
        fh.write(f" dTheta:  {float(at['dTheta']):.7E}\n")
        fh.write(f" dPhi:  {float(at['dPhi']):.7E}\n")
    return totalE


def cont_cube_from_raw(model: Model, contI_raw: np.ndarray) -> np.ndarray:
    """contI_raw: (nCells+1, nAngleBins+1) frequency sums of escapedPackets of one grid as fetched
    from the library (``mcb200_fetch_contcube``, before the host's /8).  Applies
    iteration_mod.f90:719 (/8 for symmetricXYZ; exact, so it commutes with the sum) and
    writeContCube's scaling (output_mod.f90:2774-2781)."""
    at = model.angle_tables()
    c = np.array(contI_raw, dtype=F32, order="F", copy=True)
    if model.lgSymmetricXYZ:
        c = (c / F32(8.0)).astype(F32)
    for imu in range(1, model.nAngleBins + 1):
        if F32(at["viewPointTheta"][imu]) > 0:
            c[:, imu] = c[:, imu] / F32(at["dTheta"])
        if F32(at["viewPointPhi"][imu]) > 0:
            c[:, imu] = c[:, imu] / F32(at["dPhi"])
    c[:, 0] = c[:, 0] / F32(F32(4.0) * PI)
    return c


def cont_cube_records(model: Model, contI_raw, origin=(1, 1, 1)) -> list:
    """The records of output/contCube.out (output_mod.f90:2753-2794): for every grid point
    ``iG ix iy iz contI(0:nAngleBins)``; zeros for inactive points except the mother-grid origin
    cell (iOrigin, jOrigin, kOrigin), which reads row 0 of the array like the reference does."""
    rows = []
    for iG, (g, raw) in enumerate(zip(model.grids, contI_raw), start=1):
        c = cont_cube_from_raw(model, raw)
        act = np.asarray(g.active)
        for ix in range(1, g.nx + 1):
            for iy in range(1, g.ny + 1):
                for iz in range(1, g.nz + 1):
                    a = int(act[ix - 1, iy - 1, iz - 1])
                    if a > 0 or (iG == 1 and (ix, iy, iz) == tuple(origin)):
                        # an inactive origin cell has active = 0 (row 0); a sub-grid marker (< 0) there
                        # would index out of bounds in the reference
                        rows.append((iG, ix, iy, iz) + tuple(c[max(a, 0), :]))
                    else:
                        rows.append((iG, ix, iy, iz) + (F32(0.0),) * (model.nAngleBins + 1))
    return rows


def write_cont_cube(path: str, model: Model, contI_raw, origin=(1, 1, 1)) -> None:
    """output/contCube.out; list-directed number formatting is the compiler's in the reference."""
    with open(path, "w") as fh:
        for r in cont_cube_records(model, contI_raw, origin):
            fh.write(" " + " ".join(str(v) for v in r[:4]) + " " + " ".join(f"{float(v):.7E}" for v in r[4:]) + "\n")
        fh.write("  \n")
        fh.write(" All continuum intensities given per unit direction - must multiply column 3 by 4. Pi to obtain total "
                 "emission over all directions.\n")


def write_plane_ion_distribution(path: str, planeIonDistribution: np.ndarray) -> None:
    """output/planeIonDistribution.out (iteration_mod.f90:570-577): `i k count` for every (x, z)
    column of the mother grid's illuminated face, x outermost; the array is what
    ``mcb200_fetch_plane_distribution`` returns (already summed over ranks)."""
    p = np.asarray(planeIonDistribution)
    with open(path, "w") as fh:
        for i in range(p.shape[0]):
            for k in range(p.shape[1]):
                fh.write(f" {i + 1:11d} {k + 1:11d} {int(p[i, k]):11d}\n")


# ---------------------------------------------------------------------------------------
# output/tauNu.out: writeTauNu (output_mod.f90:2384-2505) over integratePathTauNu
# (pathIntegration_mod.f90:241-470).  The march from the origin along +x, +z, +y visits the same
# cells for every frequency, so it is done once per direction (`tau_path`) and the optical depth of
# every bin is the float32 running sum of opacity(cell_k, nu)*dlSmall over its steps -- which needs
# the opacity rows of a few dozen cells only (``mcb200_get_opacity_rows``), not the table.
# ---------------------------------------------------------------------------------------
MAX_TAU = 10_000_000             # constants_mod.f90:46


def _axis_start(a: np.ndarray, v) -> int:
    from .model import locate

    p = locate(a, v)
    if 1 <= p < a.shape[0]:      # p = 0 (below the axis) is rejected by the caller, as in the reference
        if F32(v) >= F32(F32(a[p - 1] + a[p]) / F32(2.0)):
            p += 1
    return p


def tau_path(model: Model, uHat, aVec=(0.0, 0.0, 0.0)):
    """The cells integratePathTauNu steps through from aVec along uHat in the mother grid (single
    grid only, like the reference).  Returns (cells int32[nsteps], dlSmall float32)."""
    from .model import locate

    if model.nGrids > 1:
        return np.zeros(0, np.int32), F32(0.0)
    g = model.grids[0]
    ax = [np.asarray(g.xAxis, F32), np.asarray(g.yAxis, F32), np.asarray(g.zAxis, F32)]
    n = [a.shape[0] for a in ax]
    P = [_axis_start(ax[k], aVec[k]) for k in range(3)]
    for k in range(3):
        if P[k] <= 0 or P[k] > n[k]:
            raise ValueError("integratePathTau: starting position is outside the grid")
    dl = F32(0.0)
    for k in range(3):           # half the smallest spacing of the axes the ray advances along (:283-318)
        if uHat[k] > 0:
            dmin = F32(np.abs(np.diff(ax[k])).astype(F32).min())
            dl = dmin if dl <= 0 else F32(min(dl, dmin))
    dl = F32(dl / F32(2.0))
    v = [F32(uHat[0]), F32(uHat[1]), F32(uHat[2])]
    r = [F32(F32(aVec[k]) + F32(dl * v[k])) for k in range(3)]
    sym = bool(model.lgSymmetricXYZ)
    Rout = F32(model.R_out)
    cells = []
    act = np.asarray(g.active)
    for _ in range(MAX_TAU):
        if r[0] > ax[0][-1] or r[1] > ax[1][-1] or r[2] > ax[2][-1]:
            break
        rad = F32(np.sqrt(F32(F32(F32(r[0] / F32(1e10)) ** 2 + F32(r[1] / F32(1e10)) ** 2) + F32(r[2] / F32(1e10)) ** 2))) * F32(1e10)
        if Rout > 0 and rad >= Rout:
            break
        if sym:
            for k in range(3):
                if r[k] < ax[k][0]:
                    v[k], r[k] = F32(-v[k]), F32(-r[k])
                    P[k] = locate(ax[k], r[k])
        stop = False
        for k in range(3):
            if P[k] < n[k]:
                if r[k] > F32(F32(ax[k][P[k] - 1] + ax[k][P[k]]) / F32(2.0)):
                    P[k] += 1
            elif P[k] == n[k]:
                if r[k] > ax[k][P[k] - 1]:
                    stop = True
                    break
            if P[k] > 1:
                if r[k] < F32(F32(ax[k][P[k] - 1] + ax[k][P[k] - 2]) / F32(2.0)):
                    P[k] -= 1
        if stop:
            break
        if not sym and (P[0] < 1 or P[1] < 1 or P[2] < 1):
            break
        if sym:
            for k in range(3):
                if P[k] < 1:
                    v[k], r[k] = F32(-v[k]), F32(-r[k])
                    P[k] = locate(ax[k], r[k])
                    if P[k] < n[k]:
                        if r[k] > F32(F32(ax[k][P[k] - 1] + ax[k][P[k]]) / F32(2.0)):
                            P[k] += 1
                    elif P[k] == n[k] and r[k] > ax[k][P[k] - 1]:
                        stop = True
                        break
            if stop:
                break
        cells.append(int(act[P[0] - 1, P[1] - 1, P[2] - 1]))
        r = [F32(r[k] + F32(dl * v[k])) for k in range(3)]
    return np.asarray(cells, np.int32), dl


TAU_DIRECTIONS = (((1.0, 0.0, 0.0), "1,0,0"), ((0.0, 0.0, 1.0), "0,0,1"), ((0.0, 1.0, 0.0), "0,1,0"))


def tau_nu(model: Model, opacity_rows) -> list:
    """outTau(1:nbins) for the three directions of writeTauNu.  `opacity_rows(cells)` returns the
    (len(cells), nbins) float32 rows opacity(cell, :) -- from the host table or from the device
    (``PacketEngine.get_opacity_rows``)."""
    out = []
    for uHat, _ in TAU_DIRECTIONS:
        cells, dl = tau_path(model, uHat)
        tau = np.zeros(model.nbins, dtype=F32)
        if cells.size:
            uniq, inv = np.unique(cells, return_inverse=True)
            rows = np.asarray(opacity_rows(uniq.astype(np.int32)), dtype=F32)
            for k in inv:
                tau = (tau + (rows[k] * dl).astype(F32)).astype(F32)
        out.append(tau)
    return out


def write_tau_nu(path: str, model: Model, opacity_rows) -> list:
    """output/tauNu.out in the reference's layout."""
    taus = tau_nu(model, opacity_rows)
    nu = np.asarray(model.nuArray, dtype=F32)
    lam = (F32(C_LIGHT) / (nu * FR1RYD)).astype(F32) * F32(1.0e4)
    with open(path, "w") as fh:
        for (uHat, label), tau in zip(TAU_DIRECTIONS, taus):
            fh.write("  Optical depth from the centre to the edge of the nubula \n")
            fh.write(f"  direction: {label}\n")
            fh.write("  lambda [um]    tau(R_max) \n")
            for f in range(model.nbins):
                fh.write(f" {float(lam[f]):.7E} {float(tau[f]):.7E}\n")
            fh.write("  \n")
    return taus
