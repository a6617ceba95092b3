"""ctypes binding of the C ABI in include/mcb200.h (no torch types cross it)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmocassin_b200.so")

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "nGrids", "nbins", "nStars", "nAngleBins", "totAngleBinsTheta", "totAngleBinsPhi", "nLines",
        "lgDust", "lgGas", "lgSymmetricXYZ", "lgIsotropic", "lgPlaneIonization", "lgDebug",
        "lgMultistars", "lgMultiDustChemistry", "nSpeciesMax", "nSizes", "nDustComp")] + [
        (n, C.c_float) for n in ("dTheta", "dPhi", "R_out", "ionEdge1")]


class Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "nPackets", "nAbs", "nSca", "trapped", "nLinePackets", "nDropped", "nSegments", "nFlights",
        "nEscaped", "nEarlyEscaped")] + [("Qphot", C.c_double), ("kernel_ms", C.c_double), ("total_ms", C.c_double),
                ("nLaunches", C.c_int64), ("nWaves", C.c_int64), ("fly_ms", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_SIGS = {
    "mcb200_create": [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_uint64],
    "mcb200_destroy": [C.c_void_p],
    "mcb200_set_config": [C.c_void_p, C.POINTER(Config)],
    "mcb200_set_grid": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                        c_float_p, c_float_p, c_float_p, c_int32_p],
    "mcb200_set_spectra": [C.c_void_p, c_float_p, c_float_p, c_float_p],
    "mcb200_set_stars": [C.c_void_p, c_float_p, c_int32_p],
    "mcb200_set_viewpoints": [C.c_void_p, c_int32_p, c_int32_p, c_float_p, c_float_p],
    "mcb200_set_dust_species": [C.c_void_p, c_int32_p, c_float_p, c_int32_p, c_float_p, C.c_int32],
    "mcb200_set_opacity": [C.c_void_p, C.c_int32, c_float_p, c_float_p],
    "mcb200_set_xsec": [C.c_void_p, c_float_p, C.c_int64],
    "mcb200_assemble_opacity": [C.c_void_p, C.c_int32, C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
                                C.c_int32, c_float_p, c_float_p, c_float_p, c_float_p, c_int32_p, c_float_p,
                                c_int32_p, c_int32_p, C.c_int32],
    "mcb200_get_opacity": [C.c_void_p, C.c_int32, c_float_p, c_float_p, c_float_p],
    "mcb200_get_opacity_rows": [C.c_void_p, C.c_int32, C.c_int32, c_int32_p, c_float_p],
    "mcb200_set_pdfs": [C.c_void_p, C.c_int32, c_float_p, c_float_p, c_float_p, c_float_p],
    "mcb200_set_dust_state": [C.c_void_p, C.c_int32, c_float_p, c_int32_p],
    "mcb200_set_dust_tables": [C.c_void_p, c_float_p, c_float_p, c_int32_p, C.c_int32, c_float_p, C.c_int32],
    "mcb200_dust_update": [C.c_void_p, C.c_int32, C.c_float, c_float_p, c_int32_p, C.POINTER(C.c_int64)],
    "mcb200_dust_pdf": [C.c_void_p, C.c_int32, c_float_p],
    "mcb200_photo_integrals": [C.c_void_p, C.c_int32, C.c_int32, c_int32_p, c_int32_p, c_int32_p,
                               c_float_p, c_float_p, c_float_p, c_float_p],
    "mcb200_reduce_range": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32],
    "mcb200_escaped_compact": [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)],
    "mcb200_escaped_scatter": [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64],
    "mcb200_fetch_sed": [C.c_void_p, c_float_p, C.POINTER(C.c_int64)],
    "mcb200_fetch_contcube": [C.c_void_p, C.c_int32, c_float_p],
    "mcb200_zero_estimators": [C.c_void_p],
    "mcb200_transport": [C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.POINTER(Counters)],
    "mcb200_transport_diffuse": [C.c_void_p, C.c_int32, c_int32_p, C.c_int64, C.c_float, C.POINTER(Counters)],
    "mcb200_set_res_line_packets": [C.c_void_p, C.c_int32, c_int32_p],
    "mcb200_transport_reslines": [C.c_void_p, C.c_int32, C.c_float, C.POINTER(Counters)],
    "mcb200_tally_buffer": [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), c_int64_p],
    "mcb200_reduce": [C.c_void_p],
    "mcb200_comm_unique_id": [C.c_void_p, C.c_void_p],
    "mcb200_comm_init": [C.c_void_p, C.c_void_p],
    "mcb200_comm_destroy": [C.c_void_p],
    "mcb200_exchange": [C.c_void_p],
    "mcb200_exchange_info": [C.c_void_p, c_int64_p, c_int32_p, c_int32_p],
    "mcb200_exchange_path": [C.c_void_p, c_int32_p, C.c_char_p, C.c_int64, C.POINTER(C.c_double)],
    "mcb200_fetch_estimators": [C.c_void_p, C.c_int32, c_float_p, c_float_p, c_float_p, c_float_p],
    "mcb200_fetch_estimators_cells": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, c_float_p, c_float_p, c_int64_p],
    "mcb200_checksum": [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)],
    "mcb200_fetch_escaped_sparse": [C.c_void_p, C.c_int32, c_float_p, C.c_int32, c_int64_p],
    "mcb200_fetch_estimators_sparse": [C.c_void_p, C.c_int32, c_float_p, c_float_p, C.c_int32, c_int64_p],
    "mcb200_fetch_tallies": [C.c_void_p, C.c_int32, c_int64_p, c_int64_p, c_int64_p, c_int64_p],
    "mcb200_len_unit": [C.c_void_p, C.c_int32, C.POINTER(C.c_double)],
    "mcb200_fetch_plane_distribution": [C.c_void_p, c_int32_p],
    "mcb200_fetch_qphot_counts": [C.c_void_p, c_int64_p],
    "mcb200_fetch_fates": [C.c_void_p, c_int32_p, C.c_int64],
    "mcb200_set_option": [C.c_void_p, C.c_char_p, C.c_int64],
    "mcb200_test_detmath": [C.c_void_p, C.c_int32, c_float_p, c_float_p, C.c_int64],
    "mcb200_test_uniforms": [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int32, c_float_p],
    "mcb200_test_push_kernels": [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_double)],
    "mcb200_test_access_peak": [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_double)],
}

EXPORTS = sorted(list(_SIGS) + ["mcb200_last_error", "mcb200_nccl_info"])

_lib = None


def _prefer_bundled_nccl():
    """The loader keeps ONE object per SONAME: if the library's dlopen("libnccl.so.2") found the
    system copy first, a later `import torch` would be served that (older) copy and fail on a
    missing symbol (seen on the GPU box: ncclDevCommCreate).  So, unless the user chose one,
    point MCB200_NCCL_LIB at the copy torch itself loads (the nvidia-nccl wheel)."""
    if os.environ.get("MCB200_NCCL_LIB"):
        return
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["MCB200_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def load() -> C.CDLL:
    """Load the CUDA library; there is no fallback -- a missing .so is an error."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MCB200_LIB", LIB_PATH)       # A/B builds of the same library (scripts/)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m mocassin_b200.build` "
            "(nvcc, sm_100a). mocassin_b200 has no CPU fallback.")
    _prefer_bundled_nccl()
    lib = C.CDLL(path)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.mcb200_last_error.argtypes = [C.c_void_p]
    lib.mcb200_last_error.restype = C.c_char_p
    lib.mcb200_nccl_info.argtypes = [c_int32_p, C.c_char_p, C.c_int64]
    lib.mcb200_nccl_info.restype = C.c_int
    _lib = lib
    return lib


def nccl_info() -> tuple[int, str]:
    """(ncclGetVersion, path of the libnccl the library bound); raises if NCCL cannot be loaded."""
    lib = load()
    v = C.c_int32()
    buf = C.create_string_buffer(4096)
    rc = lib.mcb200_nccl_info(C.byref(v), buf, 4096)
    if rc != 0:
        raise ImportError("libmocassin_b200.so could not bind NCCL (libnccl.so.2; set MCB200_NCCL_LIB)")
    return v.value, buf.value.decode()
