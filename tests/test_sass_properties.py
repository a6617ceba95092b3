"""Static properties of the built sm_100a code that the measurements in DESIGN.md rest on, read with
cuobjdump from the in-tree library (no GPU needed): the register budget that gives the FLY kernel its
4 CTAs per SM, the 64-bit reductions of the J tally, the 256-bit loads / 128-bit peer stores of the
packed push, and that the library carries sm_100a code only."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mocassin_b200", "libmocassin_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)), reason="needs the built library and cuobjdump")


def _run(*args):
    return subprocess.run([CUOBJDUMP, *args, LIB], capture_output=True, text=True, check=True).stdout


@pytest.fixture(scope="module")
def usage():
    out = _run("-res-usage")
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
        res[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)), local=int(m.group(5)))
    return res


def _sass(symbol_part):
    out = _run("-sass")
    keep, on = [], False
    for line in out.splitlines():
        if "Function : " in line:
            on = symbol_part in line
        if on:
            keep.append(line)
    return "\n".join(keep)


def test_library_is_sm_100a_only():
    archs = set(re.findall(r"arch = (sm_\w+)", _run("-lelf") + _run("-lptx")))
    elf = set(re.findall(r"\.(sm_\w+)\.", _run("-lelf")))
    assert (archs | elf) == {"sm_100a"}


def test_fly_kernel_register_budget(usage):
    fly = {k: v for k, v in usage.items() if "wf_fly_kernel" in k}
    assert len(fly) == 9                                   # MULTI x DENSE x MODE variants that are launched
    for name, u in fly.items():
        multi = "ILb1E" in name
        # 64 registers = 4 CTAs of 256 threads per SM (single grid); 80 = 3 CTAs (sub-grid variants)
        assert u["reg"] <= (80 if multi else 64), (name, u)
        assert u["stack"] <= 24 and u["local"] == 0, (name, u)


def test_fly_kernel_tallies_with_64_bit_reductions():
    sass = _sass("wf_fly_kernelILb0ELb1ELi2")
    assert "REDG.E.ADD.64" in sass                          # fire-and-forget 64-bit adds into JsteQ
    assert sass.count("ATOMG.E.ADD.64") <= 1                # the one returning 64-bit atomic is the work counter


def test_packed_push_uses_wide_accesses():
    sass = _sass("p2p_push_packed_kernel")
    assert re.search(r"LDG\.E\.[A-Z0-9.]*256", sass)        # ld.global.cs.v4.u64: one sector per lane
    assert re.search(r"STG\.E\.128", sass)                  # 16-byte peer stores
    loads = [i for i, l in enumerate(sass.splitlines()) if "LDG" in l and "256" in l]
    stores = [i for i, l in enumerate(sass.splitlines()) if "STG.E.128" in l]
    assert loads and stores and max(loads) < min(stores)    # both loads issued before the first store
    merge = _sass("p2p_sum_fold_packed_kernelILi8")
    assert re.search(r"STG\.E\.128", merge) and re.search(r"LDG\.E\.[A-Z0-9.]*128", merge)
