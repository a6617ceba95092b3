"""Generate tests/golden/ref_<case>.npz from the reference's own code.

    python tests/golden/make_golden.py [case ...]

Runs in the build container only (needs /root/reference): oracle/f90ref translates
photon_mod.f90 & co. into oracle/_ref/mocassin_ref.py, tests/ref_cases.run_reference
executes the reference's energyPacketDriver on each seeded case with the oracle's Philox
stream bound to RANDOM_NUMBER and detmath bound to LOG/SIN/COS/ACOS/ATAN, and the
reference's own output arrays are stored here.  The fixtures travel to the GPU box; the
reference does not.

Per case: fates (n,2) int32 = cell crossings and energyPacketRun calls per packet; draws (n,)
= uniforms consumed per packet; Jste_g<i>, escapedPackets_g<i> (+ Jdif, linePackets in
debug mode) = grid(i)%... float32 exactly as the Fortran accumulates them; Qphot, absInt,
scaInt; plane = planeIonDistribution.

ref_aux_<case>.npz hold the outputs of the reference routines either side of the transport
(oracle/f90ref/harness_aux.py): the opacity block of iterateMC (opacity, scaOpac, absOpac and the
free-free term ff1 = FFOpacity(1) per cell), emissionDriver -> setDustPDF (dustPDF) and
updateCell -> getDustT (Tdust, lgConverged), and the photo-ionisation rate / heating loops of
updateCell / thermBalance (nPhotoSte, nPhotoDif per element and ion, heatSte, heatDif per cell, and
getOuterShell's shell numbers) on seeded inputs built by tests/ref_cases.py.  ref_aux_sed_<case>.npz:
ref_aux_taunu_<case>.npz: the three optical-depth columns writeTauNu writes to output/tauNu.out.  ref_aux_mie.npz: BHmie, getQs, linearMap and the assembly of makeDustXsec on seeded inputs.  ref_aux_contcube_<case>.npz: the records writeContCube writes to output/contCube.out.  ref_aux_sed: 
what writeSED writes to output/SED.out (nu, lambda, SED per viewing angle, total energy) for the
escapedPackets of the transport case of the same name.  ref_aux_writegrid.json (and _2d: with the 2D flag set): the records
writeGrid writes to grid0-3.out, dustGrid.out and photoSource.out (numbers rendered with
checkpoint.py's format; list-directed formatting is the compiler's business).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_cases  # noqa: E402


def main(names):
    for name in names or list(ref_cases.REF_CASES) + list(ref_cases.AUX_CASES) + list(ref_cases.PHOTO_CASES) + \
            ["sed:" + c for c in ref_cases.SED_CASES] + ["contcube:" + c for c in ref_cases.SED_CASES] + ["taunu:" + c for c in ref_cases.TAUNU_CASES] + ["mie", "starpos", "writegrid", "writegrid_2d"]:
        if name in ("writegrid", "writegrid_2d"):
            import json
            res = ref_cases.run_reference_writegrid(lg2D=name.endswith("_2d"))
            path = os.path.join(HERE, f"ref_aux_{name}.json")
            with open(path, "w") as fh:
                json.dump(res, fh, indent=0)
            print(f"{name}: {os.path.getsize(path)} bytes, " + ", ".join(f"{k}: {len(v)} records" for k, v in res.items()))
            continue
        if name.startswith("taunu:"):
            res = ref_cases.run_reference_taunu(name[6:])
            path = os.path.join(HERE, f"ref_aux_taunu_{name[6:]}.npz")
            np.savez_compressed(path, **res)
            print(f"{name}: {os.path.getsize(path)} bytes, max tau {float(res['tau_x'].max()):.4g} {float(res['tau_z'].max()):.4g} "
                  f"{float(res['tau_y'].max()):.4g}")
            continue
        if name == "starpos":
            res = ref_cases.run_reference_starpos()
            path = os.path.join(HERE, "ref_aux_starpos.npz")
            np.savez_compressed(path, **res)
            print(f"{name}: {os.path.getsize(path)} bytes, " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in res.items()))
            continue
        if name == "mie":
            res = ref_cases.run_reference_mie()
            path = os.path.join(HERE, "ref_aux_mie.npz")
            np.savez_compressed(path, **res)
            print(f"{name}: {os.path.getsize(path)} bytes, " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in res.items()))
            continue
        if name.startswith("contcube:"):     # after the transport cases too
            res = ref_cases.run_reference_contcube(name[9:])
            path = os.path.join(HERE, f"ref_aux_contcube_{name[9:]}.npz")
            np.savez_compressed(path, **res)
            print(f"{name}: {os.path.getsize(path)} bytes, {res['index'].shape[0]} records, origin {res['origin'].tolist()}")
            continue
        if name.startswith("sed:"):          # after the transport cases: reads their golden files
            res = ref_cases.run_reference_sed(name[4:])
            path = os.path.join(HERE, f"ref_aux_sed_{name[4:]}.npz")
            np.savez_compressed(path, **res)
            print(f"{name}: {os.path.getsize(path)} bytes, totalE {float(res['totalE'])}")
            continue
        if name in ref_cases.AUX_CASES or name in ref_cases.PHOTO_CASES:
            res = ref_cases.run_reference_aux(name) if name in ref_cases.AUX_CASES else ref_cases.run_reference_photo(name)
            path = os.path.join(HERE, f"ref_aux_{name}.npz")
            np.savez_compressed(path, **res)
            print(f"{name}: {os.path.getsize(path)} bytes, " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in res.items()))
            continue
        res = ref_cases.run_reference(name)
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **res)
        print(f"{name}: {os.path.getsize(path)} bytes, {res['draws'].shape[0]} packets, {int(res['nSegments'])} crossings")


if __name__ == "__main__":
    main(sys.argv[1:])
