"""scripts/run_dust_deck.py end to end on the CPU: the script's own loop and output files
(SED.out, summary.out, dustGrid.out, grid0.out, photoSource.out) with the CUDA engine replaced by
an oracle-backed stand-in.  On a GPU the same script runs the real engine."""
import json
import os
import runpy
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def test_device_opacity_flag_takes_the_same_path_to_the_same_files(tmp_path, monkeypatch, capsys, oracle_lib):
    """--device-opacity (dust opacities rebuilt through assemble_opacity with Tdust = None) and the
    default (host rebuild + upload) must produce identical output files."""
    import deck_runner
    from mocassin_b200 import api

    monkeypatch.setattr(api, "PacketEngine", deck_runner.OracleEngine)
    outs = []
    for flag in ([], ["--device-opacity"]):
        out = tmp_path / ("dev" if flag else "host")
        monkeypatch.setattr(sys, "argv", ["run_dust_deck.py", "--golden", os.path.join(ROOT, "tests", "golden", "deck_p0tau10.npz"),
                                          "--out", str(out), "--max-iter", "2", "--seed", "7"] + flag)
        runpy.run_path(os.path.join(ROOT, "scripts", "run_dust_deck.py"), run_name="__main__")
        outs.append(out)
    capsys.readouterr()
    for fn in ("SED.out", "tauNu.out", "dustGrid.out", "summary.out"):
        assert (outs[0] / fn).read_text() == (outs[1] / fn).read_text(), fn


def test_run_dust_deck_script_writes_the_reference_output_files(tmp_path, monkeypatch, capsys, oracle_lib):
    import deck_runner
    from mocassin_b200 import api, checkpoint, deck

    monkeypatch.setattr(api, "PacketEngine", deck_runner.OracleEngine)
    out = tmp_path / "output"
    monkeypatch.setattr(sys, "argv", ["run_dust_deck.py", "--golden", os.path.join(ROOT, "tests", "golden", "deck_p0tau1.npz"),
                                      "--out", str(out), "--max-iter", "2", "--seed", "40"])
    runpy.run_path(os.path.join(ROOT, "scripts", "run_dust_deck.py"), run_name="__main__")
    lines = [json.loads(x) for x in capsys.readouterr().out.strip().splitlines()]
    assert [r["iteration"] for r in lines[:2]] == [1, 2] and lines[1]["converged_pct"] > lines[0]["converged_pct"]
    summary = lines[-1]
    assert summary["iterations"] == 2 and summary["escaped_packets"] == summary["packets"] == 100000
    assert abs(summary["total_energy_out_e36"] - 38.26) < 1e-3 * 38.26          # writeSED's total = LStar
    T = summary["Tdust_along_x"]
    assert 600 < T[0] < 900 and T[0] > T[3] > T[6]
    for fn in ("SED.out", "tauNu.out", "summary.out", "dustGrid.out", "grid0.out", "photoSource.out"):
        assert (out / fn).stat().st_size > 0, fn
    sed = (out / "SED.out").read_text().splitlines()
    assert sed[0].startswith(" Spectral energy distribution") and len([x for x in sed if x[:2] == " 1" or x[:2] == " 2"]) > 50
    assert "Total energy radiated out of the nebula" in sed[-3]
    assert "% converged cells in grid  1" in (out / "summary.out").read_text()
    tau = np.array([[float(x) for x in ln.split()] for ln in (out / "tauNu.out").read_text().splitlines()[3:3 + 215]])
    k = int(np.argmin(abs(tau[:, 0] - 1.0)))
    assert abs(tau[k, 1] - 1.0) < 0.03                  # tau(1 um) = 1 along +x: the benchmark's definition
    # dustGrid.out reads back to the temperatures of the run
    m, t, d = deck.deck_from_arrays(dict(np.load(os.path.join(ROOT, "tests", "golden", "deck_p0tau1.npz"))))
    checkpoint.read_dust_grid(str(out / "dustGrid.out"), m)
    cells = [int(c) for c in m.grids[0].active[:, 0, 0] if c > 0]
    back = [round(float(m.grids[0].Tdust[0, 0, c]), 2) for c in cells]
    assert np.allclose(back, T, rtol=1e-5)
