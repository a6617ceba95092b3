"""The C-ABI library loads without a GPU and exports every symbol include/mcb200.h
declares; without a device it fails loudly instead of falling back to the CPU."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb200_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(cuda_lib):
    from mocassin_b200 import _lib

    names = _declared()
    assert len(names) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mcb200_[a-z_0-9]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert sorted(_lib.EXPORTS) == names
    for n in names:
        assert hasattr(cuda_lib, n)


def test_library_contains_sm100a_code_only(cuda_lib):
    from mocassin_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mocassin_b200 import workloads as W
    from mocassin_b200.api import MocassinError, PacketEngine

    with pytest.raises(MocassinError):
        PacketEngine(W.hii_region())


def test_product_package_never_imports_oracle():
    """The oracle is test infrastructure: the product package and the scripts never touch it (only
    tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do)."""
    for f in os.listdir(os.path.join(ROOT, "scripts")):
        if f.endswith((".py", ".sh")):
            text = open(os.path.join(ROOT, "scripts", f)).read()
            assert "import oracle" not in text and "from oracle" not in text, f
    pkg = os.path.join(ROOT, "mocassin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace(
                    "oracle/mc_oracle.c ERR_STOP", ""), f


def test_fortran_binding_covers_the_header():
    """fortran/mcb200_mod.f90 (the ISO_C_BINDING shim a maintainer of the reference adds) declares
    every entry point of include/mcb200.h except the diagnostic / unit-test hooks."""
    hdr = open(os.path.join(ROOT, "include", "mcb200.h")).read()
    f90 = open(os.path.join(ROOT, "fortran", "mcb200_mod.f90")).read()
    exported = set(re.findall(r"^(?:int|const char \*)\s*(mcb200_\w+)\s*\(", hdr, flags=re.M))
    bound = set(re.findall(r'name="(mcb200_\w+)"', f90))
    diagnostic = {"mcb200_fetch_fates", "mcb200_fetch_tallies", "mcb200_test_access_peak", "mcb200_test_detmath",
                  "mcb200_test_uniforms", "mcb200_nccl_info", "mcb200_test_push_kernels"}
    assert exported - bound == diagnostic, sorted(exported - bound - diagnostic)
    assert bound <= exported, sorted(bound - exported)


def test_nccl_is_bound_at_run_time_not_link_time(cuda_lib):
    """No DT_NEEDED on NCCL (single-rank hosts never load it); mcb200_nccl_info binds it on demand."""
    from mocassin_b200 import _lib

    out = subprocess.run(["readelf", "-d", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in out.lower()
    version, path = _lib.nccl_info()
    assert version >= 22000 and os.path.exists(path), (version, path)


def test_bound_nccl_is_the_one_torch_loads():
    """One object per SONAME and process: the library must bind the copy torch brings, or a later
    `import torch` is served an older system libnccl and dies on a missing symbol (seen on the GPU
    box).  Fresh interpreter: library first, torch second."""
    code = ("from mocassin_b200 import _lib; v, p = _lib.nccl_info(); import torch; "
            "t = torch.cuda.nccl.version(); assert v == t[0] * 10000 + t[1] * 100 + t[2], (v, t); print('ok', p)")
    env = {k: v for k, v in os.environ.items() if k != "MCB200_NCCL_LIB"}
    res = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=env)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-600:]


def test_header_is_plain_c_and_a_c_program_links_against_the_library(cuda_lib, tmp_path):
    """include/mcb200.h is the contract a Fortran/C host compiles against: it must be valid C99 (no
    C++ or torch types), and a C program using it must link against the shared library and fail
    cleanly where there is no device."""
    from mocassin_b200 import _lib

    hdr = os.path.join(ROOT, "include", "mcb200.h")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                ["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    src = tmp_path / "host.c"
    src.write_text('#include "mcb200.h"\n#include <stdio.h>\n'
                   'int main(void) { mcb200_ctx *c = 0; int rc = mcb200_create(&c, 0, 0, 1, 12345u);\n'
                   '  int v = 0; char p[512]; int rn = mcb200_nccl_info(&v, p, 512);\n'
                   '  printf("%d %d %d\\n", rc, rn, v); if (rc == 0) mcb200_destroy(c); return 0; }\n')
    exe = tmp_path / "host"
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        _lib.LIB_PATH, "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rc, rn, v = (int(x) for x in out.stdout.split())
    import torch

    assert (rc == 0) == torch.cuda.is_available()        # MCB200_ENODEV without a device: no CPU fallback
    assert rc in (0, -1)
    assert (rn == 0 and v >= 20000) or rn == -8           # NCCL found through the default search, or MCB200_ECOMM


def test_fortran_binding_interfaces_are_well_formed():
    """No Fortran compiler exists here, so lint the shim structurally: blocks balance, every dummy
    of every bind(C) interface is declared exactly once inside its block, and the number of dummies
    equals the number of parameters of the C prototype it binds."""
    hdr = open(os.path.join(ROOT, "include", "mcb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {m.group(1): [p for p in m.group(2).split(",") if p.strip() and p.strip() != "void"]
              for m in re.finditer(r"\b(mcb200_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)}
    text = open(os.path.join(ROOT, "fortran", "mcb200_mod.f90")).read()
    # join continuation lines, drop comments
    logical, cur = [], ""
    for raw in text.splitlines():
        line = raw.split("!")[0].rstrip() if '"' not in raw.split("!")[0] or raw.count("!") == 0 else raw.rstrip()
        if not line.strip():
            continue
        line = line.strip()
        if line.startswith("&"):
            line = line[1:].lstrip()
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        logical.append(cur + line)
        cur = ""
    low = [s.lower() for s in logical]
    starts = [i for i, s in enumerate(low) if re.search(r"\bfunction\s+mcb200_\w+\s*\(", s) and not s.startswith("end")]
    ends = [i for i, s in enumerate(low) if s.startswith("end function")]
    assert len(starts) == len(ends) >= 40
    assert sum(s.startswith("interface") for s in low) == sum(s.startswith("end interface") for s in low)
    for a, b in zip(starts, ends):
        assert a < b
        sig = logical[a]
        name = re.search(r'name\s*=\s*"(mcb200_\w+)"', sig).group(1)
        assert re.search(r"function\s+" + name + r"\s*\(", sig, flags=re.I), sig      # Fortran name = C name
        dummies = [d.strip().lower() for d in re.search(r"function\s+\w+\s*\((.*?)\)", sig, flags=re.I).group(1).split(",") if d.strip()]
        declared = []
        for s in logical[a + 1:b]:
            for part in s.split(";"):
                if "::" in part:
                    declared += [re.sub(r"\(.*?\)", "", v).strip().lower() for v in part.split("::", 1)[1].split(",") if v.strip()]
        assert sorted(declared) == sorted(dummies), (name, sorted(set(dummies) ^ set(declared)))
        assert len(dummies) == len(protos[name]), (name, len(dummies), len(protos[name]))


def test_c_host_example_compiles_and_reports_no_device(cuda_lib, tmp_path):
    """examples/host_example.c: the whole call sequence from plain C (no Python, no torch in the
    process).  Here it must compile warning-free and exit 2 (MCB200_ENODEV); on a GPU box
    tests/test_zz_gpu_deck.py runs it and expects exit 0."""
    from mocassin_b200 import _lib

    exe = tmp_path / "host_example"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "host_example.c"), _lib.LIB_PATH,
                        "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH), "-lm", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import torch

    if not torch.cuda.is_available():
        out = subprocess.run([str(exe)], capture_output=True, text=True)
        assert out.returncode == 2 and "MCB200_ENODEV" in out.stdout
