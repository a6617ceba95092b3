#!/usr/bin/env python
"""bench.py -- propagated energy packets per second of the transport hot path.

Workload (config.workload): BASELINE.md "S-clumpy": 128^3 full cube, star at the centre,
log-normal clumpy gas + dust, nbins=600, Philox seed 12345.  One *step* = one pass of
the hot path over one batch of packets: `mcb200_transport` (emission + transport kernel)
followed by the fold epilogue, i.e. `call energyPacketDriver` + the estimator merge of one
Lucy iteration (iteration_mod.f90:474-726).  Weak scaling: every GPU transports
--packets packets per step against a replicated grid; for N>1 the integer tallies are
summed with an NCCL all-reduce inside the step (the path's one real exchange).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this implementation
  python bench.py --impl reference ...                            # CPU arm (oracle port)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 12345
ALG_BYTES_PER_SEGMENT = 16       # active 4 + opacity 4 + Jste 4 read + 4 write


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--nbins", type=int, default=600)
    ap.add_argument("--packets", type=int, default=0, help="packets per GPU per step (0 = default)")
    ap.add_argument("--workload", default="clumpy", choices=["clumpy", "uniform"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU-baseline sample time")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def build_model(args, tables: bool):
    from mocassin_b200 import workloads as W

    return W.synthetic_cube(n=args.grid, nbins=args.nbins, clumpy=(args.workload == "clumpy"), dust=True,
                            nPhotons=10 ** 9, seed=2024, build_tables=tables)


def compact_inputs(model):
    """The per-cell inputs of K1 for this workload: 3 gas species columns + dust."""
    from mocassin_b200 import workloads as W
    from mocassin_b200.model import F32, I32

    g = model.grids[0]
    f = model._fields3d
    nu = model.nuArray
    sH, sHe, sHe2 = W.gas_cross_sections(nu)
    cabs, csca, gs = W.dust_optics(nu)
    model.gSca = gs
    nb = model.nbins
    xsec = np.concatenate([sH, sHe, sHe2, csca, cabs]).astype(F32)
    Hd, x = f["Hden"], f["xH0"]
    den = np.zeros((g.nCells + 1, 3), dtype=F32, order="F")
    den[:, 0] = W._per_cell(g.active, Hd * x, g.nCells)
    den[:, 1] = W._per_cell(g.active, 0.1 * Hd * np.minimum(1.0, 3.0 * x), g.nCells)
    den[:, 2] = W._per_cell(g.active, 0.1 * Hd * (1.0 - np.minimum(1.0, 3.0 * x)) * 0.9, g.nCells)
    bands = dict(species=np.array([1, 2, 3], I32), off=np.array([0, nb, 2 * nb], I32),
                 low=np.array([1, 1, 1], I32), high=np.array([nb, nb, nb], I32))
    Nd = W._per_cell(g.active, f["Ndust"], g.nCells)
    Td = np.full((2, 2, g.nCells + 1), 50.0, dtype=F32, order="F")
    dust = dict(Ndust=Nd, Tdust=Td, dustAbunIndex=None, grainWeight=np.ones(1, F32),
                dustScaXsecP=np.array([[3 * nb + 1]], I32), dustAbsXsecP=np.array([[4 * nb + 1]], I32))
    g.Tdust = Td
    g.dustAbunIndex = np.ones(g.nCells + 1, dtype=I32)
    return xsec, bands, den, dust


def rec_tables(model, rng):
    from mocassin_b200 import workloads as W
    from mocassin_b200.model import F32

    g = model.grids[0]
    nu = model.nuArray
    wid = np.gradient(nu).astype(F32)
    Tes = np.linspace(6000.0, 12000.0, 16)
    table = np.stack([W.recombination_cdf(nu, wid, T) for T in Tes]).astype(F32)
    idx = rng.integers(0, 16, size=g.nCells + 1)
    tl = np.zeros(g.nCells + 1, dtype=F32)
    tl[1:] = (0.55 + 0.2 * rng.random(g.nCells)).astype(F32)
    return table, idx, tl


def host_tables(model):
    """The same tables the device path builds (K1 + rec_tables), assembled with numpy in the
    same float32 operation order, for the CPU arm."""
    from mocassin_b200 import workloads as W
    from mocassin_b200.model import F32

    g = model.grids[0]
    nb = model.nbins
    xsec, bands, den, dust = compact_inputs(model)
    sH, sHe, sHe2, csca, cabs = (xsec[i * nb:(i + 1) * nb] for i in range(5))
    op = W._expand(den[:, 0], sH)
    op += W._expand(den[:, 1], sHe)
    op += W._expand(den[:, 2], sHe2)
    sca = W._expand(dust["Ndust"], csca)
    ab = W._expand(dust["Ndust"], cabs)
    ab += sca
    op += ab
    del ab
    g.opacity, g.scaOpac, g.absOpac = op, sca, None
    table, idx, tl = rec_tables(model, np.random.default_rng(2025))
    g.recPDF = W._rows_from_table(idx, table)
    g.recPDF[0, :] = 0.0
    g.totalLines = tl


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None
        self.first = 0

    def mark(self):
        """Start of the timed region: only rows sampled from here on are reported."""
        self.first = len(self.rows)

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "transport_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def default_packets(args):
    return args.packets if args.packets > 0 else 125_000_000   # 1e9 packets per iteration over 8 GPUs (BASELINE.json)


# ---------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the oracle port (faithful restatement of photon_mod.f90, -O2) on all host
    cores, packets split with the reference's load/rest rule, replicated read-only tables,
    shared integer tallies.  The reference binary itself cannot be built here (no Fortran
    compiler)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mocassin_b200 import workloads as W
    from oracle.oracle import Oracle

    cores = os.cpu_count() or 1
    t0 = time.time()
    model = build_model(args, tables=False)
    g = model.grids[0]
    host_tables(model)
    o = Oracle(model, fp32_tallies=False)
    build_s = time.time() - t0
    # size the per-step sample for ~cpu-seconds of work
    t = time.time(); c = o.transport_mt(1, 0, 400000, seed=SEED, threads=cores); dt = time.time() - t
    rate = 400000 / max(dt, 1e-6)
    # bounded sample per step: the whole --steps/--warmup run stays within ~2.5 minutes of CPU time
    per_step_s = min(args.cpu_seconds, 150.0 / max(args.steps + args.warmup, 1))
    sample = int(max(400000, min(rate * per_step_s, 200_000_000)))
    first = 400000
    for _ in range(args.warmup):
        o.transport_mt(1, first, sample, seed=SEED, threads=cores); first += sample
    segs = 0
    t = time.time()
    for _ in range(args.steps):
        c = o.transport_mt(1, first, sample, seed=SEED, threads=cores); first += sample
        segs += c["nSegments"]
    dt = time.time() - t
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "energy packets/sec", "value": value, "unit": "packets/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same config as the b200 arm; each step here is a bounded sample of it (cpu_baseline.sample)
        "config": dict(workload_config(args, default_packets(args)), sample_packets_per_step=sample,
                       note="same workload; every step of this arm transports a bounded sample of it (a rate metric)"),
        "cpu_baseline": {"value": value, "unit": "packets/s", "cores": cores, "kind": "port",
                         "sample_packets_per_step": sample,
                         "sample": f"{sample} packets/step x {args.steps} steps of the same {args.grid}^3 workload, "
                                   f"oracle port (-O2, detmath instead of libm: same algorithm as the CUDA path), {cores} threads "
                                   f"sharing the integer tallies through atomics; table build {build_s:.0f}s untimed"},
        "e2e": {"value": value, "unit": "packets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "segments_per_packet": segs / (sample * args.steps),
    }
    print(json.dumps(line), flush=True)


def workload_config(args, packets):
    table_gb = (args.grid ** 3 + 1) * args.nbins * 4 / 1e9
    return {"workload": f"S-{args.workload} {args.grid}^3 gas+dust nebula, nbins={args.nbins}, star at centre, "
                        f"Philox seed {SEED}", "grid": args.grid, "nbins": args.nbins,
            "packets_per_gpu_per_step": int(packets),
            "cache": f"inputs larger than L2: opacity/scaOpac/recPDF/Jste tables {table_gb:.2f} GB each (float32; "
                     f"JsteQ twice that), L2 126 MB" if table_gb > 0.126 else
                     f"tables {table_gb * 1e3:.1f} MB each: L2-resident (reduced --grid/--nbins, not the headline workload)",
            "parallelism": f"packets sharded over {args.gpus} GPU(s), grid replicated"}


# ---------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    else:
        torch.cuda.set_device(local)
    from mocassin_b200.api import PacketEngine
    from mocassin_b200.model import F32

    P = default_packets(args)
    model = build_model(args, tables=False)
    g = model.grids[0]
    xsec, bands, den, dust = compact_inputs(model)
    rng = np.random.default_rng(2025)
    table, idx, tl = rec_tables(model, rng)
    nRows, nb = g.nCells + 1, model.nbins

    # pinned host buffers for the per-iteration inputs/outputs of the e2e path
    def pinned(shape, dtype=torch.float32):
        n = int(np.prod(shape))
        t = torch.empty(n, dtype=dtype, pin_memory=True)
        return t, t.numpy().reshape(shape, order="F")

    keep = []
    t_rec, recPDF = pinned((nRows, nb)); keep.append(t_rec)
    np.take(table.T, idx, axis=1, out=recPDF.T)
    recPDF[0, :] = 0.0
    g.recPDF, g.totalLines = recPDF, tl

    fake = int(os.environ.get("MCB_PACK_STATS", "0")) if world == 1 else 0      # diagnostic, see pack_stats()
    eng = PacketEngine(model, device=local, rank=rank, nranks=fake or world, seed=SEED)
    eng.set_xsec(xsec)
    for opt in ("order", "agg_steps", "batch", "blocks_per_sm", "wavefront", "step_budget", "tail", "fly_batch", "wave0_order", "wave0_blocks", "wave0_exact"):
        if os.environ.get("MCB_" + opt.upper()):
            eng.set_option(opt, int(os.environ["MCB_" + opt.upper()]))

    def upload_inputs():
        eng.assemble_opacity(1, bands, den, None, dust)       # K1 on device
        eng.set_pdfs()
        eng.set_dust_state()

    upload_inputs()
    nGlobal = P * world
    dE = float(model.deltaE[1])

    if fake:
        # MCB_PACK_STATS=N on one GPU: what rank 0 of N would hand to the packed push -- how many of its
        # partial sums need more than 32 bits, per element and per 256-element flag block
        from mocassin_b200.api import _as_cuda_tensor
        eng.zero_estimators()
        eng.energyPacketDriver(1, P * fake, deltaE=dE)
        ptr, n = eng.tally_buffer(1, 0)
        q = _as_cuda_tensor(ptr, n, "<i8", local)
        nblk = n // 256
        out = {"pack_stats_for_ranks": fake, "elements": n}
        for name, bits in (("over_32_bits", 32), ("over_24_bits", 24), ("over_16_bits", 16), ("non_zero", 0)):
            m = (q >> bits) != 0 if bits else q != 0
            out[name] = {"elements": float(m.float().mean()), "blocks_of_256": float(m[: nblk * 256].view(nblk, 256).any(1).float().mean())}
            del m
        print(json.dumps(out))
        return

    # measured on 2 GPUs: splitting the batch in halves costs more (smaller waves, NCCL sharing
    # the SMs) than hiding half of the exchange gains -> off by default
    overlap = world > 1 and os.environ.get("MCB_OVERLAP", "0") == "1"
    if os.environ.get("MCB_PIPELINED_FOLD"):
        eng.pipelined_fold = os.environ["MCB_PIPELINED_FOLD"] == "1"
    if os.environ.get("MCB_CHUNK_PLANES"):
        eng.exchange_chunk_planes = int(os.environ["MCB_CHUNK_PLANES"])
    if os.environ.get("MCB_SPARSE_ESCAPED"):
        eng.sparse_escaped = os.environ["MCB_SPARSE_ESCAPED"] == "1"
    if world > 1 and os.environ.get("MCB_SED_LOCAL", "0") == "1":
        # exchange the (nu, angle) escape counts instead of the per-cell escapedPackets tallies;
        # escapedPackets then stays rank-local (see mcb200_fetch_sed).  Off by default: the
        # reference all-reduces escapedPackets (iteration_mod.f90:649-659).
        eng.set_sed_local(True)

    native = world > 1 and os.environ.get("MCB_NATIVE_COMM", "1") == "1"
    if native:
        # the library's own NCCL communicator (mcb200_comm_init / mcb200_exchange: what a Fortran/MPI
        # host calls): reduce-scatter of the touched JsteQ planes -> fold of this rank's share ->
        # all-gather of the float32 Jste, sparse all-gather of the escape counts, all on the library
        # stream.  MCB_NATIVE_COMM=0: torch.distributed all-reduces on the tally buffers (round 1).
        eng.comm_init_from_group()
        if os.environ.get("MCB_EXCHANGE_ALLREDUCE", "0") == "1":
            eng.set_option("exchange_allreduce", 1)
        if os.environ.get("MCB_EXCHANGE_PACK"):
            eng.set_option("exchange_pack", int(os.environ["MCB_EXCHANGE_PACK"]))
        if os.environ.get("MCB_EXCHANGE_PUSH"):
            eng.set_option("exchange_push", int(os.environ["MCB_EXCHANGE_PUSH"]))
        if os.environ.get("MCB_EXCHANGE_PUSH_BLOCKS"):
            eng.set_option("exchange_push_blocks", int(os.environ["MCB_EXCHANGE_PUSH_BLOCKS"]))
        if os.environ.get("MCB_EXCHANGE_P2P"):
            eng.set_option("exchange_p2p", int(os.environ["MCB_EXCHANGE_P2P"]))

    def nrank_parity(nCheck=2_000_000):
        """Untimed: the same nCheck packets once sharded over the N ranks (exchange + fold) and once
        by every rank alone (option solo); the estimators every rank holds must be the same bits
        (mcb200_checksum of Jste and escapedPackets), the packet fates must add up."""
        keys = ("nPackets", "nEscaped", "nLinePackets", "nDropped", "trapped", "nAbs", "nSca", "nSegments")
        eng.zero_estimators()
        c = eng.energyPacketDriver(1, nCheck, deltaE=dE)
        eng.reduce()
        sharded = [eng.checksum(1, 0), eng.checksum(1, 1)]
        tot = torch.tensor([c[k] for k in keys], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        eng.set_option("solo", 1)
        eng.zero_estimators()
        c1 = eng.energyPacketDriver(1, nCheck, deltaE=dE)
        solo = [eng.checksum(1, 0), eng.checksum(1, 1)]
        eng.set_option("solo", 0)
        eng.zero_estimators()
        tot = [int(v) for v in tot.tolist()]
        same = sharded == solo and tot == [c1[k] for k in keys]
        conserved = tot[1] + tot[2] + tot[3] + tot[4] == nCheck == tot[0]
        ok = torch.tensor([int(same), int(conserved)], dtype=torch.int64, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return {"nrank_parity": bool(ok[0].item()), "packets_conserved": bool(ok[1].item()), "packets": nCheck,
                "checksums": {"Jste": f"{sharded[0]:016x}", "escapedPackets": f"{sharded[1]:016x}"},
                "fates": dict(zip(keys, tot)),
                "how": f"{nCheck} packets sharded over {world} ranks + exchange + fold vs the same packets on every rank "
                       "alone (option solo): mcb200_checksum of the device-resident Jste / escapedPackets and all counters "
                       "equal on every rank; nEscaped + nLinePackets + nDropped + trapped == packets"}

    parity = nrank_parity() if world > 1 else None

    def step():
        if overlap:      # exchange of the first half hidden behind the transport of the second
            return eng.energyPacketDriverOverlapped(1, nGlobal, deltaE=dE)
        c = eng.energyPacketDriver(1, nGlobal, deltaE=dE)
        if world > 1:
            eng.reduce()
        return c

    eng.zero_estimators()
    # the sampler process is started before the warm-up so that it is already delivering rows
    # when the timed region begins; rows before mark() are discarded
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.mark()
    t0 = time.perf_counter()
    kms, tms, segs, flights, launches, waves, fms = 0.0, 0.0, 0, 0, 0, 0, 0.0
    for _ in range(args.steps):
        ts = time.perf_counter()
        c = step()
        wall_ms = 1e3 * (time.perf_counter() - ts)
        kms += c["kernel_ms"]
        tms += c["total_ms"] if world == 1 else wall_ms      # N>1: include the all-reduce + fold
        segs += c["nSegments"]; flights += c["nFlights"]; launches += c["nLaunches"]; waves += c["nWaves"]
        fms += c["fly_ms"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.finish() if sampler else None
    # max over ranks of the device-timed step time
    tt = torch.tensor([tms, kms, float(segs), wall], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        tms_max, kms_max, wall_max = float(mx[0]), float(mx[1]), float(mx[3])
        segs_all = float(sm[2])
    else:
        tms_max, kms_max, wall_max, segs_all = tms, kms, wall, float(segs)
    value = nGlobal * args.steps / (tms_max / 1e3)

    # ---- e2e through the public API with host buffers (rank-local copies inside) -------
    e2e = None
    if not args.no_e2e:
        # N = 1: the whole Jste (dense) + escapedPackets (sparse) come back, the whole CDF table goes up.
        # N > 1: what each rank of the reference needs after the merge -- the Jste rows of its own
        # round-robin cells (iteration_mod.f90:832), escapedPackets on rank 0 only (:738-740) -- and
        # 1/N of the CDF table per rank over PCIe, all-gathered over NVLink (option pdf_slabs).
        sharded = world > 1 and native and os.environ.get("MCB_E2E_SHARDED", "1") == "1"
        nMine = (g.nCells - (rank + 1)) // world + 1 if sharded else nRows
        t_J, Jh = pinned((nMine, nb)); keep.append(t_J)
        Eh = None
        if not sharded or rank == 0:
            t_E, Eh = pinned((nRows, nb + 1, 1)); keep.append(t_E)
            Eh[...] = 0.0
        nE = 2 if args.steps > 2 else args.steps
        eng.set_option("async_pdfs", 1)                       # PDF upload overlaps the stellar wave
        if sharded:
            eng.set_option("pdf_slabs", 1)

        def e2e_step(sparse):
            eng.assemble_opacity(1, bands, den, None, dust)   # H2D: den, Ndust, Tdust + K1
            eng.set_dust_state()
            eng.set_pdfs()                                    # H2D (async): recPDF (this rank's slab when sharded), totalLines
            eng.zero_estimators()
            step()
            if sharded:
                eng.fetch_cells(1, out=Jh)                    # D2H: Jste(iCell, :) of this rank's cells, compact
                if rank != 0:
                    return 0
                _, nnz = eng.fetch_escaped_sparse(1, out=Eh, clear_previous=True)
                return 8 * nnz + 8 if nnz >= 0 else Eh.nbytes
            if not sparse:
                eng.fetch(1, out={"Jste": Jh, "escapedPackets": Eh})   # D2H, both arrays dense
                return Eh.nbytes
            # D2H: Jste dense; escapedPackets: only its non-zero entries cross PCIe and are written
            # into Eh by host threads while Jste is still copying; the entries of the previous step
            # are zeroed first (clear_previous), so Eh ends up exactly as the dense fetch leaves it
            _, nnz = eng.fetch_sparse(1, out={"Jste": Jh, "escapedPackets": Eh}, clear_previous=True)
            return 8 * nnz + 8 if nnz >= 0 else Eh.nbytes

        e2e_step(True)                                        # untimed: sizes the staging buffers
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        te = time.perf_counter()
        for _ in range(nE):
            esc_bytes = e2e_step(True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dte = time.perf_counter() - te
        dense = None
        if not sharded:
            td = time.perf_counter()
            e2e_step(False)                                   # the dense fetch of both arrays, for comparison
            torch.cuda.synchronize()
            dte_dense = time.perf_counter() - td
            dense = {"value": nGlobal / dte_dense, "d2h_bytes_per_step": int(Jh.nbytes + Eh.nbytes),
                     "note": "same step with mcb200_fetch_estimators for both arrays (one step)"}
        small = tl.nbytes + den.nbytes + dust["Ndust"].nbytes + dust["Tdust"].nbytes + 4 * (g.nCells + 1)
        pdf_bytes = recPDF.nbytes if not sharded else 4 * nRows * min((nb + world - 1) // world, max(nb - rank * ((nb + world - 1) // world), 0))
        h2d = pdf_bytes + small
        d2h = Jh.nbytes + esc_bytes
        if world > 1:
            tmax = torch.tensor([dte, float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dte, h2d, d2h = float(tmax[0]), float(tmax[1]), float(tmax[2])     # the busiest rank's bytes
        e2e = {"value": nGlobal * nE / dte, "unit": "packets/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": nE, "bytes_are": "per rank (max over ranks)",
               "path": ("assemble_opacity + set_dust_state + set_pdfs (async H2D from pinned host: each rank uploads its 1/N slab of "
                        "nu-planes of recPDF, slabs all-gathered over NVLink; overlaps wave 0) -> zero_estimators -> "
                        "energyPacketDriver -> mcb200_exchange + mcb200_reduce -> mcb200_fetch_estimators_cells: Jste rows of this "
                        "rank's round-robin cells (iteration_mod.f90:832), compact, D2H to pinned host; rank 0 alone fetches "
                        "escapedPackets (sparse)") if sharded else
                       ("assemble_opacity + set_dust_state + set_pdfs (async H2D from pinned host, overlaps wave 0) -> "
                        "zero_estimators -> energyPacketDriver -> mcb200_fetch_estimators_sparse: Jste dense D2H to "
                        "pinned host, overlapped with the non-zero escapedPackets entries being written into the host array"),
               "dense_fetch": dense}

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    peak, peak_kind = measured_peak()
    ach = ALG_BYTES_PER_SEGMENT * segs / (kms / 1e3) / 1e9        # rank 0's kernel
    tr = ncu_traffic()
    # `achieved`: algorithmic bytes of one mcb200_transport call / CUDA-event time of ALL its kernels (the
    # wave-front pipeline: FLY + event + sort kernels, on the library stream); `fly_kernel`: the same bytes
    # over the device time of the cell-crossing kernel wf_fly_kernel alone (events around each of its
    # launches).  `traffic` is not measured by this run: it is the dram__bytes_read+write sum of one
    # `ncu --set full` pass over the same command, kept in profiles/transport_traffic.json with its source.
    fly_ach = ALG_BYTES_PER_SEGMENT * segs / (fms / 1e3) / 1e9 if fms > 0 else None
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": (tr or {}).get("dram_bytes_per_launch"), "traffic_source": (tr or {}).get("source"),
                "peak_kind": f"of {peak_kind}",
                "kernel": "all kernels of one mcb200_transport call (wave-front pipeline: mcb::wf_fly_kernel + "
                          "wf_event_kernel<EMIT|SCATTER|CONT> + wf_escape_compact_kernel + sort kernels)",
                "fly_kernel": {"name": "mcb::wf_fly_kernel<false,true,2>", "ms_per_launch": fms / args.steps,
                               "share_of_call": fms / kms if kms > 0 else None, "achieved": fly_ach,
                               "frac": fly_ach / peak if fly_ach else None},
                "algorithmic_bytes_per_segment": ALG_BYTES_PER_SEGMENT,
                "segments_per_launch": segs / args.steps, "kernel_ms_per_launch": kms / args.steps,
                "segments_per_s": segs / (kms / 1e3)}
    access = access_roofline(eng, g, nb, segs / (kms / 1e3)) if world == 1 else None
    cpu = None
    if not args.no_cpu and world == 1:
        cpu = cpu_baseline(args, eng, model, g)
    line = {
        "metric": "energy packets/sec", "value": value, "unit": "packets/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, P), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches),            # counted by the library: wave-front kernels + fold kernels
        "waves_per_step": waves / args.steps,
        "roofline": roofline, "access_roofline": access, "cpu_baseline": cpu,
        "segments_per_packet": segs_all / (nGlobal * args.steps),
        "flights_per_packet": flights / (P * args.steps),
        "kernel_ms_per_step": kms_max / args.steps, "wall_ms_per_step": 1e3 * wall_max / args.steps,
        "exchange_planes": getattr(eng, "last_exchange_planes", None),
        "escaped_exchange": getattr(eng, "last_escaped_exchange", None),
        "exchange": ({"path": "native: mcb200_exchange (flags, sparse all-gather of escapedQ) -> mcb200_reduce (J merge: "
                              + str((getattr(eng, "last_exchange", None) or {}).get("path")) + ")" if native else
                              "torch.distributed all-reduce of the tally buffers -> mcb200_reduce",
                      "bytes_per_rank_per_step": (getattr(eng, "last_exchange", None) or {}).get("bytes"),
                      "detail": getattr(eng, "last_exchange", None),
                      "ms_per_step": (tms_max - kms_max) / args.steps} if world > 1 else None),
    }
    if parity:
        line.update({"nrank_parity": parity["nrank_parity"], "packets_conserved": parity["packets_conserved"],
                     "nrank_parity_detail": parity})
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def access_roofline(eng, g, nb, segs_per_s):
    """The second roofline SURVEY.md 8d asks for: cell crossings per second against the measured
    rate at which this GPU serves the crossing's access pattern and nothing else -- one 4-byte
    opacity read plus one 64-bit Jste reduction at addresses the lanes of a warp do not share
    (mcb200_test_access_peak).  `peak`: both windows the size of one nu-plane (what the
    frequency-ordered FLY kernel keeps in L2); `peak_dram`: windows the size of the whole tables."""
    import ctypes as C

    plane_q, plane_f = 8 * (g.nCells + 1), 4 * (g.nCells + 1)

    def peak(mode, qbytes, fbytes, ops):
        v = C.c_double()
        eng._check(eng.lib.mcb200_test_access_peak(eng.h, mode, int(qbytes), int(fbytes), int(ops), C.byref(v)))
        return float(v.value)

    try:
        mixed_l2 = peak(3, plane_q, plane_f, 1 << 29)
        red_l2 = peak(1, plane_q, plane_f, 1 << 29)
        mixed_dram = peak(3, plane_q * nb, plane_f * nb, 1 << 27)
    except Exception as ex:          # a measurement aid must not take the bench line down
        return {"error": str(ex)}
    return {"bound": "memory-system request rate: 1 LDG.32 + 1 RED.64 per cell crossing at scattered addresses",
            "achieved": segs_per_s, "peak": mixed_l2, "unit": "crossings/s", "frac": segs_per_s / mixed_l2,
            "peak_kind": "measured here: access_peak_kernel, windows = one nu-plane of opacity / JsteQ (L2 resident)",
            "peak_red64_only": red_l2, "peak_dram": mixed_dram,
            "frac_of_dram_regime": segs_per_s / mixed_dram}


def cpu_baseline(args, eng, model, g):
    """Oracle port timed on this box's host cores on a bounded sample of the same workload
    (tables read back from the device so both sides use identical inputs)."""
    from oracle.oracle import Oracle

    cores = os.cpu_count() or 1
    op, sca, _ = eng.get_opacity(1)
    g.opacity, g.scaOpac = op, sca
    o = Oracle(model, fp32_tallies=False)
    t = time.time(); o.transport_mt(1, 0, 400000, seed=SEED, threads=cores); dt = time.time() - t
    rate = 400000 / max(dt, 1e-6)
    sample = int(max(400000, min(rate * args.cpu_seconds, 200_000_000)))
    t = time.time(); c = o.transport_mt(1, 400000, sample, seed=SEED, threads=cores); dt = time.time() - t
    return {"value": sample / dt, "unit": "packets/s", "cores": cores, "kind": "port",
            "sample": f"{sample} packets of the same workload in {dt:.1f}s, oracle port (-O2, detmath instead of libm) on "
                      f"{cores} threads sharing the integer tallies through atomics",
            "segments_per_packet": c["nSegments"] / sample}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
