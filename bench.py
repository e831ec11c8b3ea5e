#!/usr/bin/env python
"""bench.py — one JSON line per run.

A step = one pass of the hot path (element stiffness -> assembly -> Dirichlet elimination -> PCG to
||r||/||b|| <= 1e-9 -> reactions + stress) over one synthetic plate, through mag_solve with the library's
default options.  Default workload: BASELINE.json configs[3], the 16 M-DOF plate (4000 x 2000 cells).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nx NX --ny NY]

`value`   whole-path Melem/s with the mesh already resident in HBM (device-generated plate), timed with
          CUDA events on the launching stream, max over ranks;
`e2e`     the same through mag_solve with pinned HOST buffers (H2D of the mesh and D2H of
          ux,uy,fx,fy,stress inside the timed region);
`roofline` the PCG SpMV kernel (SELL-32 + fused p.q): algorithmic bytes / CUDA-event time of back-to-back
          launches of that kernel, against MEASURED_PEAKS.json hbm_gbs;
`proof`   what the timed steps computed, checked after the timed region at every N: the TRUE residual
          ||b - K_ff x|| / ||b|| of the returned displacements (ordered CSR kernel on a fresh assembly) and the
          relative L2 distance to a tight solve (rel_tol 1e-13) of the same system; at N > 1 also the real
          NCCL / CUDA-IPC / mailbox path against the one-process virtual-rank emulation of the same partition.
          The run EXITS NON-ZERO when a bound is missed;
`cpu_baseline` / `--impl reference`: the oracle port (oracle/magnetite_oracle.c: the reference's arithmetic,
          CSR storage, 1 thread like the reference) on a bounded sample OF THE SAME WORKLOAD: a strip of the
          same plate (same row length, same band) through the whole path with a bounded number of CG
          iterations, scaled to the whole job; a completely solved smaller plate is reported beside it
          together with the GPU on that very plate (`metrics.same_config_sample`).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "whole-path Melem/s (assembly + PCG to 1e-9 + stress); also assembly Melem/s, PCG time-to-solve, SpMV HBM GB/s"
UNIT = "Melem/s"
FALLBACK_HBM_GBS = 6650.0
# Bounds of the correctness proof (exit code 3 when missed).
MAX_TRUE_REL_RESIDUAL = 2e-9        # the solver stops on the recursive residual <= 1e-9
MAX_REL_L2_VS_TIGHT = 1e-5          # distance of the 1e-9 solve to a 1e-13 solve of the same system: a residual of 1e-9 bounds
                                    # the error by kappa*1e-9 only; measured on the 16 M-DOF plate: 1.7e-6 (Jacobi-PCG)
MAX_MULTI_GPU_REL_L2 = 1e-9         # real multi-process path against the virtual-rank emulation

# Jacobi-PCG iterations to ||r||/||b|| <= 1e-9.  Measured on a B200 (profiles/r1_bench_*.json, r2_bench_*.json) except
# 1000x500, which is the count of the CPU restatement (oracle/two_level.py: the same recurrence in the same precision;
# it reproduces the B200's count wherever both ran — 400x200: 2801 Jacobi / 450 two-level, 1000x500: 424 two-level).
JACOBI_ITERS = {("c4", 4000, 2000): 17203, ("c4", 1000, 500): 6856, ("c4", 400, 200): 2801,
                ("c5", 4000, 2000): 33555, ("c5", 11314, 5657): 90354}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=4000)
    ap.add_argument("--ny", type=int, default=2000)
    ap.add_argument("--ref-nx", type=int, default=400, help="completely solved CPU plate (measured pair beside the sample)")
    ap.add_argument("--ref-ny", type=int, default=200)
    ap.add_argument("--sample-elems", type=int, default=512000, help="CPU sample: triangles of the strip")
    ap.add_argument("--sample-iters", type=int, default=30, help="CPU sample: CG iterations run on the strip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the comparison solves / assemblies beside the timed path")
    ap.add_argument("--spmv-reps", type=int, default=100)
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4: the 16 M-DOF plate (strong scaling); c5: perforated plate, 16 M DOF per GPU (weak scaling)")
    ap.add_argument("--spmv-format", type=int, default=0, help="0 SELL-32 with packed 16-bit column offsets when the band allows (default), 3 SELL-32 with 32-bit columns")
    ap.add_argument("--allreduce", type=int, default=0, help="multi-GPU dot products: 0 peer-memory mailbox, 1 NCCL")
    ap.add_argument("--precond", type=int, default=-1, help="-1: the library's default; 1 Jacobi; 2 two-level")
    return ap.parse_args()


def grid_of(args, world):
    """(nx, ny) of the plate this run works on: c5 keeps 16 M DOF per GPU."""
    if args.workload == "c5":
        return int(round(args.nx * world ** 0.5)), int(round(args.ny * world ** 0.5))
    return args.nx, args.ny


def workload_name(nx, ny):
    return f"synthetic structured plate {nx}x{ny} cells: {2 * nx * ny} CST triangles, " \
           f"{2 * (nx + 1) * (ny + 1)} DOF, h=2, E=69e9, nu=0.33, t=0.5, left edge clamped, right edge ux=3"


def config_of(args, world):
    """The `config` object — the SAME for both arms (the reference arm runs a bounded sample of it)."""
    nx, ny = grid_of(args, world)
    weak = args.workload == "c5"
    return {"workload": ("perforated (holes of radius 16h at pitch 64h) " if weak else "") + workload_name(nx, ny),
            "parallelism": "1 GPU" if world == 1 else f"{world} GPUs, one process each: contiguous row blocks of K_ff",
            "l2": "inputs larger than L2 (K_ff + vectors >> 126 MB), no flush needed"}


def jacobi_iters_full(workload, nx, ny):
    """Jacobi-PCG iterations the whole job needs: the table where this plate has been run, else an interpolation of
    the table's iterations PER CELL ROW in log(nx) (7.0 at 400, 6.9 at 1000, 4.3 at 4000: the count grows slower
    than the plate), clamped to the ends."""
    if (workload, nx, ny) in JACOBI_ITERS:
        return JACOBI_ITERS[(workload, nx, ny)], "measured (B200 or CPU-port run of this plate)"
    pts = sorted((n, it / n) for (w, n, _), it in JACOBI_ITERS.items() if w == workload)
    lx, per = math.log(max(nx, 1)), pts[-1][1]
    if lx <= math.log(pts[0][0]):
        per = pts[0][1]
    else:
        for (n0, p0), (n1, p1) in zip(pts, pts[1:]):
            if math.log(n0) <= lx <= math.log(n1):
                per = p0 + (p1 - p0) * (lx - math.log(n0)) / (math.log(n1) - math.log(n0))
                break
    return int(round(per * nx)), "estimated: iterations per cell row interpolated between the measured plates"


def ncu_traffic(nx, ny, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the SpMV kernel from the committed
    ncu --set full capture of this workload (profiles/), or None when there is no capture for it."""
    try:
        cands = sorted((ROOT / "profiles").glob("r*_spmv_traffic.json"))
        for path in reversed(cands):
            for e in json.loads(path.read_text())["entries"]:
                if (e["nx"], e["ny"], e["n_gpus"]) == (nx, ny, world):
                    return int(e["dram_bytes_read"]) + int(e["dram_bytes_write"])
    except Exception:
        pass
    return None


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.path = None
        self.gpu = gpu_index
        if shutil.which("nvidia-smi"):
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(gpu_index)], stdout=self.fh, stderr=subprocess.DEVNULL)

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample of the workload
# ---------------------------------------------------------------------------
def cpu_complete(nx, ny):
    """Whole path through the oracle's sparse mode (reference arithmetic, CSR storage, 1 thread) on a small plate
    of the same family, solved completely with the same Jacobi-PCG to 1e-9.  Returns (Melem/s, seconds, stats)."""
    from magnetite_b200 import meshgen
    from oracle import oracle as O
    mesh = meshgen.plate(nx, ny)
    om = O.Mesh(mesh)
    t0 = time.perf_counter()
    res = O.run(om, meshgen.EXAMPLE_MATERIAL, O.cg_options(jacobi=1, rel_tol=1e-9), dense=False)
    dt = time.perf_counter() - t0
    return mesh.n_elems / dt / 1e6, dt, res["stats"]


class CpuStrip:
    """The bounded CPU sample of the workload `nx x ny`: the strip nx x strip_ny of the SAME plate (same row
    length, same band, same boundary conditions) through the whole path of the oracle port with a bounded
    number of CG iterations.  Every phase is streaming work linear in the rows of cells, and one CG iteration
    costs the same per row wherever the row sits, so the job costs
        (ny / strip_ny) * (t_elem + t_asm + t_part + t_react + t_stress + iters_full * t_per_iteration).
    The strip's vectors fit the CPU's last-level cache better than the job's would: the scaling favours the CPU."""

    def __init__(self, workload, nx, ny, sample_elems, sample_iters):
        from magnetite_b200 import meshgen
        from oracle import oracle as O
        self.O, self.meshgen = O, meshgen
        self.workload, self.nx, self.ny = workload, nx, ny
        self.strip_ny = max(1, min(ny, int(round(sample_elems / (2.0 * nx)))))
        if workload == "c5":          # whole rows of holes (pitch 64 cells)
            self.strip_ny = min(ny, max(64, 64 * int(round(self.strip_ny / 64.0))))
        self.iters = max(1, sample_iters)
        gen = meshgen.perforated_plate if workload == "c5" else meshgen.plate
        self.mesh = gen(nx, self.strip_ny)
        self.om = O.Mesh(self.mesh)
        self.iters_full, self.iters_source = jacobi_iters_full(workload, nx, ny)
        self.elems_full = None        # c5: the perforated job's triangle count is set by the caller when known
        self.last = None

    def step(self):
        """One sample; returns the job's estimated seconds."""
        O = self.O
        opt = O.cg_options(jacobi=1, rel_tol=1e-9, max_iter=self.iters)
        t0 = time.perf_counter()
        res = O.run(self.om, self.meshgen.EXAMPLE_MATERIAL, opt, dense=False)
        wall = time.perf_counter() - t0
        st = res["stats"]
        it = max(int(st["iters"]), 1)
        scale = self.ny / self.strip_ny
        serial = st["t_elem"] + st["t_asm"] + st["t_part"] + st["t_react"] + st["t_stress"]
        t_iter = st["t_solve"] / it
        est = scale * (serial + self.iters_full * t_iter)
        self.last = {"strip": f"{self.nx}x{self.strip_ny} cells ({self.mesh.n_elems} triangles)", "wall_seconds": wall,
                     "cg_iterations_run": it, "seconds_per_cg_iteration": t_iter, "seconds_serial_phases": serial,
                     "scale_rows": scale, "job_cg_iterations": self.iters_full, "job_cg_iterations_source": self.iters_source,
                     "job_seconds_estimated": est}
        return est

    def describe(self):
        s = self.last
        return (f"strip {s['strip']} of the plate, whole path of the oracle port (reference arithmetic, CSR storage, "
                f"Jacobi-PCG, 1 thread) with {s['cg_iterations_run']} CG iterations per step; job = "
                f"{s['scale_rows']:.4g} x (serial phases {s['seconds_serial_phases']:.3f} s + {s['job_cg_iterations']} "
                f"iterations [{s['job_cg_iterations_source']}] x {s['seconds_per_cg_iteration'] * 1e3:.3f} ms)")


def cpu_all_cores(nx, ny, one_thread_seconds, one_thread_stats):
    """Information beside the 1-thread baseline: the completely solved small plate with the CG on every host core
    (oracle/oracle_mt.c, the port's Jacobi-PCG on the oracle's K_ff, pthreads).  Not the reference's algorithm — the
    reference is single-threaded — and never the headline; None if it cannot be produced."""
    try:
        from magnetite_b200 import meshgen
        from oracle import oracle as O, oracle_mt as MT
        mesh = meshgen.plate(nx, ny)
        om = O.Mesh(mesh)
        csr, rhs, _ = O.partition(om, O.assemble_sparse(om, O.element_stiffness(om, meshgen.EXAMPLE_MATERIAL)), dense=False)
        threads = MT.host_cores()
        t0 = time.perf_counter()
        _, iters, _ = MT.pcg(csr, rhs, rel_tol=1e-9, threads=threads)
        cg_s = time.perf_counter() - t0
        total = one_thread_seconds - float(one_thread_stats["t_solve"]) + cg_s      # serial phases + parallel CG
        return {"threads": threads, "cg_seconds": cg_s, "cg_iters": iters, "seconds": total,
                "value": mesh.n_elems / total / 1e6, "unit": UNIT,
                "note": "same plate, CG on all host cores; not the reference's algorithm (the reference is single-threaded)"}
    except Exception as err:                       # an extra figure must not cost the benchmark line
        return {"error": str(err)}


def cpu_baseline_block(args, world, steps, warmup, extras=True):
    """Runs the CPU sample `steps` times; returns (value, seconds per step, cpu_baseline dict)."""
    nx, ny = grid_of(args, world)
    strip = CpuStrip(args.workload, nx, ny, args.sample_elems, args.sample_iters)
    for _ in range(min(warmup, 1)):
        strip.step()
    est, walls = [], []
    for _ in range(max(steps, 1)):
        est.append(strip.step()); walls.append(strip.last["wall_seconds"])
    job_s = statistics.mean(est)
    n_elems_job = 2 * nx * ny if args.workload == "c4" else int(round(strip.mesh.n_elems * ny / strip.strip_ny))
    value = n_elems_job / job_s / 1e6
    cb = {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": strip.describe(),
          "host_cores_available": os.cpu_count(), "job_seconds_estimated": job_s, "sample_detail": strip.last,
          "sample_wall_seconds_per_step": statistics.mean(walls)}
    if extras:
        v, dt, cst = cpu_complete(args.ref_nx, args.ref_ny)
        cb["complete_small"] = {"workload": workload_name(args.ref_nx, args.ref_ny), "value": v, "unit": UNIT, "seconds": dt,
                                "cg_iters": int(cst["iters"]),
                                "note": "solved completely by the oracle port, 1 thread; the GPU on the same plate: "
                                        "metrics.same_config_sample of the GPU arm"}
        cb["all_cores"] = cpu_all_cores(args.ref_nx, args.ref_ny, dt, cst)
    return value, statistics.mean(walls), cb


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    value, wall, cb = cpu_baseline_block(args, world, args.steps, args.warmup)
    # the reference's own data structures (dense (2N)^2 + dense partition + dense->CSR scan, plain CG to an
    # absolute 1e-4) at a size they can hold: the README's linkedin case is ~7 k DOF
    from magnetite_b200 import meshgen
    from oracle import oracle as O
    dmesh = meshgen.plate(80, 40)
    t0 = time.perf_counter()
    dres = O.run(O.Mesh(dmesh), meshgen.EXAMPLE_MATERIAL, O.cg_options(), dense=True)
    ddt = time.perf_counter() - t0
    cb["faithful_dense"] = {"workload": "plate 80x40 cells (6400 triangles, 6642 DOF), the reference's dense algorithm and CG stopping rule",
                            "seconds": ddt, "melem_s": dmesh.n_elems / ddt / 1e6, "cg_iters": dres["stats"]["iters"],
                            "seconds_partition_and_dense_to_csr": dres["stats"]["t_part"], "seconds_cg": dres["stats"]["t_solve"]}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        # ms_per_step: what one step of THIS run took (the bounded sample); the job it stands for: job_ms_estimated
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall, "job_ms_estimated": 1e3 * cb["job_seconds_estimated"],
        "higher_is_better": True, "scaling": "weak" if args.workload == "c5" else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_of(args, world),
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    from magnetite_b200 import _lib, meshgen
    from magnetite_b200.solver import _material

    from magnetite_b200 import dist as mdist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    lib = _lib.load()
    ctx = _lib.Context(local_rank)
    tdist = None
    if world > 1:
        import torch.distributed as tdist
        mdist.init_process_group("nccl")
        mdist.init_comm(ctx)          # the library's own NCCL communicator (bootstrap, one-off gathers)

    def barrier():
        if tdist is not None:
            tdist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if tdist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        if tdist is not None:
            tdist.all_reduce(t, op=tdist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    stream = torch.cuda.Stream()
    meta = meshgen.EXAMPLE_MATERIAL
    mat = _material(meta)

    def options(**kw):
        kw.setdefault("stream", stream.cuda_stream)
        kw.setdefault("allreduce", args.allreduce)
        kw.setdefault("spmv_format", args.spmv_format)
        if args.precond >= 0:
            kw.setdefault("precond", args.precond)
        return _lib.default_options(**kw)

    opt = options()
    precond_names = {0: "plain CG", 1: "Jacobi-PCG", 2: "two-level PCG (Jacobi + aggregation coarse space)",
                     3: "auto: two-level PCG when the system is SPD, else Jacobi-PCG"}
    nx, ny = grid_of(args, world)
    weak = args.workload == "c5"
    failures = []

    # ---- device-resident workload ------------------------------------------------------
    dm = C.c_void_p()
    if weak:
        _lib.check(lib.mag_devmesh_perforated(ctx.handle, nx, ny, 2.0, 64, 16, 3.0, C.byref(dm)), "mag_devmesh_perforated")
    else:
        _lib.check(lib.mag_devmesh_plate(ctx.handle, nx, ny, 2.0, 3.0, C.byref(dm)), "mag_devmesh_plate")
    view = _lib.MagMesh()
    _lib.check(lib.mag_devmesh_view(dm, C.byref(view)), "mag_devmesh_view")
    N, E = int(view.n_nodes), int(view.n_elems)
    dev = torch.device("cuda", local_rank)

    def device_result(n, e):
        t = {k: torch.empty(n, dtype=torch.float64, device=dev) for k in ("ux", "uy", "fx", "fy")}
        t["stress"] = torch.empty(e, dtype=torch.float64, device=dev)
        return t, _lib.MagResult(t["ux"].data_ptr(), t["uy"].data_ptr(), t["fx"].data_ptr(), t["fy"].data_ptr(),
                                 t["stress"].data_ptr(), None, 1)

    out_d, res_d = device_result(N, E)

    def device_step():
        st = _lib.MagStats()
        _lib.check(lib.mag_solve(ctx.handle, C.byref(view), C.byref(mat), C.byref(opt), C.byref(res_d), C.byref(st)),
                   "mag_solve(device)")
        return st

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = []
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            stats.append(device_step())
        ev1.record(stream)
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = E / (ms_step * 1e-3) / 1e6
    last = stats[-1]
    launches = int(sum(s.kernel_launches for s in stats))
    ms_asm = max_over_ranks(statistics.mean(s.ms_elem + s.ms_sort + s.ms_reduce + s.ms_bc for s in stats))
    ms_solve = max_over_ranks(statistics.mean(s.ms_solve for s in stats))
    ms_post = max_over_ranks(statistics.mean(s.ms_post for s in stats))
    ms_upload = max_over_ranks(statistics.mean(s.ms_upload for s in stats))

    # ---- proof + roofline on a fresh assembly of the same mesh ---------------------------------------
    sysh = C.c_void_p()
    st = _lib.MagStats()
    _lib.check(lib.mag_assemble(ctx.handle, C.byref(view), C.byref(mat), C.byref(opt), C.byref(sysh), C.byref(st)),
               "mag_assemble")

    def true_rel_residual(t):
        rr, bb = C.c_double(), C.c_double()
        _lib.check(lib.mag_system_residual(sysh, t["ux"].data_ptr(), t["uy"].data_ptr(), 1, C.byref(rr), C.byref(bb)),
                   "mag_system_residual")
        srr, sbb = sum_over_ranks([rr.value, bb.value])
        return (srr / sbb) ** 0.5 if sbb > 0 else 0.0

    proof = {"true_rel_residual": true_rel_residual(out_d), "max_true_rel_residual": MAX_TRUE_REL_RESIDUAL,
             "solver_rel_residual": last.final_residual / last.b_norm if last.b_norm else 0.0}
    # a tight solve of the same system: how far is the 1e-9 answer from it?
    tight_t, tight_res = device_result(N, E)      # also receives the comparison solves below: out_d stays the timed result
    s_t = _lib.MagStats()
    opt_t = options(rel_tol=1e-13, precond=2)
    rc_t = lib.mag_system_solve(sysh, C.byref(opt_t), C.byref(tight_res), C.byref(s_t))
    if rc_t != 0:                      # not SPD (two-level refuses): Jacobi
        opt_t = options(rel_tol=1e-13, precond=1)
        _lib.check(lib.mag_system_solve(sysh, C.byref(opt_t), C.byref(tight_res), C.byref(s_t)), "mag_system_solve(tight)")
    num = torch.sqrt(((out_d["ux"] - tight_t["ux"]) ** 2).sum() + ((out_d["uy"] - tight_t["uy"]) ** 2).sum())
    den = torch.sqrt((tight_t["ux"] ** 2).sum() + (tight_t["uy"] ** 2).sum())
    proof.update({"rel_l2_vs_tight_solve": float(num / den), "max_rel_l2_vs_tight_solve": MAX_REL_L2_VS_TIGHT,
                  "tight_solve": {"rel_tol": 1e-13, "iters": int(s_t.iters), "seconds": max_over_ranks(s_t.ms_solve) * 1e-3,
                                  "true_rel_residual": true_rel_residual(tight_t)},
                  "stress_rel_inf_vs_tight_solve": float((out_d["stress"] - tight_t["stress"]).abs().max()
                                                         / tight_t["stress"].abs().max())})
    if not (proof["true_rel_residual"] <= MAX_TRUE_REL_RESIDUAL):
        failures.append(f"true relative residual {proof['true_rel_residual']:.3e} > {MAX_TRUE_REL_RESIDUAL}")
    if not (proof["rel_l2_vs_tight_solve"] <= MAX_REL_L2_VS_TIGHT):
        failures.append(f"rel L2 distance to the tight solve {proof['rel_l2_vs_tight_solve']:.3e} > {MAX_REL_L2_VS_TIGHT}")

    ms_spmv, nbytes = C.c_float(), C.c_uint64()
    _lib.check(lib.mag_system_spmv_bench(sysh, 2, args.spmv_reps, C.byref(ms_spmv), C.byref(nbytes)), "spmv_bench")
    n_local_rows = (int(st.spmv_bytes) - 12 * int(st.nnz) - 4) // 20
    peak, peak_src = measured_peak()
    csr_bytes = int(st.spmv_bytes)
    achieved = nbytes.value / (ms_spmv.value * 1e-3) / 1e9         # this rank's block, this rank's GPU
    roofline = {"bound": "hbm", "kernel": "pcg_spmv_kernel (SELL-32 SpMV + fused p.q)", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if weak else ncu_traffic(nx, ny, world),
                "algorithmic_bytes_per_launch": int(nbytes.value), "ms_per_launch": ms_spmv.value,
                # "report both" — the same launch in scalar-CSR bytes (12 B/nnz + 20 B/row)
                "csr_equivalent_bytes": csr_bytes, "csr_equivalent_achieved": csr_bytes / (ms_spmv.value * 1e-3) / 1e9,
                "csr_equivalent_frac": csr_bytes / (ms_spmv.value * 1e-3) / 1e9 / peak}

    # ---- comparison solves on the same system: the other preconditioner -------------------------------
    def solve_metrics(s):
        its = max(int(s.iters), 1)
        ms = max_over_ranks(s.ms_solve)
        iter_bytes = nbytes.value + 88 * n_local_rows
        return {"pcg_iters": int(s.iters), "pcg_time_to_solve_s": ms * 1e-3,
                "pcg_final_rel_residual": s.final_residual / s.b_norm if s.b_norm else 0.0,
                "ms_per_iteration": ms / its,
                "jacobi_iteration_bytes": iter_bytes,
                "pcg_iteration_timeline_us": [round(v / 1e3, 2) for v in list(s.prof)[:7]]}

    default_precond = int(opt.precond)
    solves = {}
    if not args.no_extras:
        for name, pc in (("jacobi", 1), ("two_level", 2)):
            if pc == default_precond:
                continue
            o2 = options(precond=pc)
            runs = []
            for _ in range(2 if pc == 2 else 1):           # the first two-level solve of a system builds the coarse space
                s2 = _lib.MagStats()
                rc = lib.mag_system_solve(sysh, C.byref(o2), C.byref(tight_res), C.byref(s2))
                if rc != 0:
                    runs = None
                    solves[name] = {"error": _lib.last_error()}
                    break
                runs.append(s2)
            if runs:
                m = solve_metrics(runs[-1])
                if pc == 2:
                    m["n_coarse"] = int(runs[0].n_coarse)
                    m["coarse_setup_s"] = max_over_ranks(runs[0].ms_coarse_setup) * 1e-3
                    m["pcg_time_to_solve_s_incl_setup"] = max_over_ranks(runs[0].ms_solve) * 1e-3
                else:
                    it_ms = m["ms_per_iteration"]
                    m["pcg_iteration_achieved_gbs"] = m["jacobi_iteration_bytes"] / (it_ms * 1e-3) / 1e9
                    m["pcg_iteration_frac_of_peak"] = m["pcg_iteration_achieved_gbs"] / peak
                m["true_rel_residual"] = true_rel_residual(tight_t)
                solves[name] = m
    lib.mag_system_free(sysh)
    del tight_t

    # ---- assembly variants beside the default, same mesh, same phase timers (single GPU) -----------------
    assemblies = None
    if world == 1 and not args.no_extras:
        assemblies = {}
        for name, mode in (("gather", 0), ("sorted_coo_keys", 1)):
            try:
                og = options(assembly=mode)
                g_ms, g_st = [], None
                for _ in range(3):
                    gsys, g_st = C.c_void_p(), _lib.MagStats()
                    _lib.check(lib.mag_assemble(ctx.handle, C.byref(view), C.byref(mat), C.byref(og), C.byref(gsys),
                                                C.byref(g_st)), "mag_assemble")
                    lib.mag_system_free(gsys)
                    g_ms.append(g_st.ms_elem + g_st.ms_sort + g_st.ms_reduce + g_st.ms_bc)
                ms_g = statistics.mean(g_ms[1:])           # the first one pays the kernels' first launch
                assemblies[name] = {"assembly_melem_s": E / (ms_g * 1e-3) / 1e6, "assembly_ms": ms_g,
                                    "ms_elem": g_st.ms_elem, "ms_sort": g_st.ms_sort, "ms_reduce": g_st.ms_reduce,
                                    "ms_bc": g_st.ms_bc, "ms_format": g_st.ms_format,
                                    "same_counts_as_default": (int(g_st.nnz), int(g_st.n_free)) == (int(st.nnz), int(st.n_free))}
            except Exception as err:                       # an extra metric must not cost the benchmark line
                assemblies[name] = {"error": str(err)}

    # ---- N > 1: the real multi-process path against the one-process emulation of the same partition ----
    multi = None
    if world > 1:
        mnx, mny = 768, 384
        host = meshgen.plate(mnx, mny)
        c = np.ascontiguousarray
        arrs = [c(host.x), c(host.y), c(host.n0.view(np.uint32)), c(host.n1.view(np.uint32)), c(host.n2.view(np.uint32)),
                c(host.ux), c(host.uy), c(host.fx), c(host.fy), c(host.known)]
        hm = _lib.MagMesh(host.n_nodes, host.n_elems, *[a.ctypes.data for a in arrs], 0)
        multi = {"workload": workload_name(mnx, mny), "cases": []}
        for pc in sorted({1, default_precond if default_precond in (1, 2) else 2}):
            om = options(rel_tol=1e-12, precond=pc)
            r_real = {k: np.empty(host.n_nodes) for k in ("ux", "uy", "fx", "fy")}
            r_real["stress"] = np.empty(host.n_elems)
            res_r = _lib.MagResult(*[r_real[k].ctypes.data for k in ("ux", "uy", "fx", "fy", "stress")], None, 0)
            s_r = _lib.MagStats()
            _lib.check(lib.mag_solve(ctx.handle, C.byref(hm), C.byref(mat), C.byref(om), C.byref(res_r), C.byref(s_r)),
                       "mag_solve(multi-GPU check)")
            if rank == 0:
                solo = _lib.Context(local_rank)              # no communicator: R virtual ranks in this process
                r_emu = {k: np.empty(host.n_nodes) for k in ("ux", "uy", "fx", "fy")}
                r_emu["stress"] = np.empty(host.n_elems)
                res_e = _lib.MagResult(*[r_emu[k].ctypes.data for k in ("ux", "uy", "fx", "fy", "stress")], None, 0)
                s_e = _lib.MagStats()
                oe = _lib.default_options(rel_tol=1e-12, precond=pc, spmv_format=args.spmv_format)
                _lib.check(lib.mag_debug_virtual_solve(solo.handle, C.byref(hm), C.byref(mat), C.byref(oe), world,
                                                       C.byref(res_e), C.byref(s_e)), "mag_debug_virtual_solve")
                solo.close()
                u, ue = np.concatenate([r_real["ux"], r_real["uy"]]), np.concatenate([r_emu["ux"], r_emu["uy"]])
                case = {"precond": pc, "iters_real": int(s_r.iters), "iters_emulated": int(s_e.iters),
                        "rel_l2_u": float(np.linalg.norm(u - ue) / np.linalg.norm(ue)),
                        "bit_identical_u": bool(u.tobytes() == ue.tobytes()),
                        "rel_inf_stress": float(np.abs(r_real["stress"] - r_emu["stress"]).max() / np.abs(r_emu["stress"]).max()),
                        "rel_inf_reactions": float(np.abs(np.concatenate([r_real["fx"] - r_emu["fx"], r_real["fy"] - r_emu["fy"]])).max()
                                                   / np.abs(r_emu["fx"]).max())}
                multi["cases"].append(case)
                if not (case["rel_l2_u"] <= MAX_MULTI_GPU_REL_L2) or abs(case["iters_real"] - case["iters_emulated"]) > max(2, case["iters_emulated"] // 100):
                    failures.append(f"multi-GPU path differs from the virtual-rank emulation: {case}")
        barrier()

    # ---- the CPU arm's completely solved plate on the GPU, host buffers (single GPU) -------------------
    same_cfg = None
    example = None
    if world == 1 and not args.no_extras:
        small = meshgen.plate(args.ref_nx, args.ref_ny)
        c = np.ascontiguousarray
        arrs = [c(small.x), c(small.y), c(small.n0.view(np.uint32)), c(small.n1.view(np.uint32)), c(small.n2.view(np.uint32)),
                c(small.ux), c(small.uy), c(small.fx), c(small.fy), c(small.known)]
        sm = _lib.MagMesh(small.n_nodes, small.n_elems, *[a.ctypes.data for a in arrs], 0)
        r_s = {k: np.empty(small.n_nodes) for k in ("ux", "uy", "fx", "fy")}
        r_s["stress"] = np.empty(small.n_elems)
        res_s = _lib.MagResult(*[r_s[k].ctypes.data for k in ("ux", "uy", "fx", "fy", "stress")], None, 0)
        same_cfg = {"workload": workload_name(args.ref_nx, args.ref_ny),
                    "note": "mag_solve with HOST buffers on the plate the CPU arm solves completely (cpu_baseline.complete_small)"}
        for name, pc in (("jacobi", 1), ("default", default_precond)):
            osm = options(precond=pc)
            ts = []
            for _ in range(4):
                s_s = _lib.MagStats()
                t0 = time.perf_counter()
                _lib.check(lib.mag_solve(ctx.handle, C.byref(sm), C.byref(mat), C.byref(osm), C.byref(res_s), C.byref(s_s)),
                           "mag_solve(same-config sample)")
                ts.append(time.perf_counter() - t0)
            dt = statistics.mean(ts[1:])
            same_cfg[name] = {"melem_s": small.n_elems / dt / 1e6, "ms": dt * 1e3, "pcg_iters": int(s_s.iters),
                              "solver": precond_names.get(pc, str(pc))}
        # BASELINE configs[0], [1]: the reference's example geometries (stand-in mesher fixtures), reference solver
        # semantics; configs[2]: the 1 M-triangle plate with the library's defaults.  Host buffers through mag_solve.
        from magnetite_b200 import solver as msolver
        from magnetite_b200.datatypes import MeshSoA
        example = {}
        for key, fixture, published in (("example_linkedin", "example_linkedin", 0.286), ("example_tensile", "example_tensile", None)):
            try:
                g = np.load(ROOT / "tests" / "golden" / f"{fixture}.npz")
                emesh = MeshSoA(g["x"], g["y"], g["n0"], g["n1"], g["n2"], g["bc_ux"], g["bc_uy"], g["bc_fx"], g["bc_fy"], g["known"])
                emeta = meta.__class__(*g["material"])
                ts = []
                for _ in range(4):
                    t0 = time.perf_counter()
                    sol = msolver.solve_soa(emesh, emeta, ctx, _lib.default_options(compat=1))
                    ts.append(time.perf_counter() - t0)
                u, ur = np.concatenate([sol.ux, sol.uy]), np.concatenate([g["ux"], g["uy"]])
                example[key] = {"workload": f"examples/{fixture[8:]} through the stand-in mesher: {emesh.n_elems} triangles, "
                                            f"{int(sol.stats['n_free'])} free DOF; reference solver semantics (plain CG, cost <= 1e-4)",
                                "iters": int(sol.stats["iters"]), "solve_ms": float(sol.stats["ms_solve"]),
                                "whole_call_ms": statistics.mean(ts[1:]) * 1e3,
                                "us_per_iteration": 1e3 * float(sol.stats["ms_solve"]) / max(int(sol.stats["iters"]), 1),
                                "rel_l2_vs_oracle_fixture": float(np.linalg.norm(u - ur) / np.linalg.norm(ur)),
                                "reference_published_solve_s": published}
                if not (example[key]["rel_l2_vs_oracle_fixture"] <= 1e-9):
                    failures.append(f"{key} differs from the oracle fixture: {example[key]['rel_l2_vs_oracle_fixture']:.3e}")
                # the reference's own published result picture of this example (tests/test_reference_picture.py):
                # distance, in pixels of the picture, of its deformed outline's points to the GPU solution's outline
                try:
                    if str(ROOT / "tests") not in sys.path:
                        sys.path.insert(0, str(ROOT / "tests"))
                    import test_reference_picture as TP
                    tri = np.stack([g["n0"], g["n1"], g["n2"]], 1).astype(np.int64)
                    n_pts, worst, mean, _ = TP.misfit(TP.PICTURES[fixture[8:]]["panels"]["solved"], g["x"] + sol.ux,
                                                      g["y"] + sol.uy, tri)
                    example[key]["reference_picture"] = {"outline_points": n_pts, "worst_px": worst, "mean_px": mean,
                                                         "max_worst_px": TP.TOL_PX}
                    if not (worst <= TP.TOL_PX):
                        failures.append(f"{key}: the GPU solution's outline is {worst:.2f} px off the reference's picture")
                except Exception as err:
                    example[key]["reference_picture"] = {"error": str(err)}
            except Exception as err:
                example[key] = {"error": str(err)}
        try:
            c3 = meshgen.plate(1000, 500)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                sol = msolver.solve_soa(c3, meta, ctx, _lib.default_options())
                ts.append(time.perf_counter() - t0)
            example["config3_plate_1m"] = {"workload": workload_name(1000, 500), "whole_call_ms": statistics.mean(ts[1:]) * 1e3,
                                           "melem_s": c3.n_elems / statistics.mean(ts[1:]) / 1e6, "pcg_iters": int(sol.stats["iters"]),
                                           "precond_used": int(sol.stats["precond_used"]),
                                           "rel_residual": float(sol.stats["final_residual"] / sol.stats["b_norm"])}
        except Exception as err:
            example["config3_plate_1m"] = {"error": str(err)}
        # configs[4]: the perforated plate at its per-GPU size (the weak-scaling unit: 16 M DOF before the holes),
        # device-generated like the timed workload, library defaults.  The whole 8-GPU job: `--workload c5 --gpus 8`.
        if not weak:
            dm5 = C.c_void_p()
            try:
                _lib.check(lib.mag_devmesh_perforated(ctx.handle, args.nx, args.ny, 2.0, 64, 16, 3.0, C.byref(dm5)),
                           "mag_devmesh_perforated")
                view5 = _lib.MagMesh()
                _lib.check(lib.mag_devmesh_view(dm5, C.byref(view5)), "mag_devmesh_view")
                n5, e5 = int(view5.n_nodes), int(view5.n_elems)
                out5, res5 = device_result(n5, e5)
                ms5, st5 = [], None
                for _ in range(3):
                    st5 = _lib.MagStats()
                    ev0.record(stream)
                    _lib.check(lib.mag_solve(ctx.handle, C.byref(view5), C.byref(mat), C.byref(opt), C.byref(res5),
                                             C.byref(st5)), "mag_solve(config 5 unit)")
                    ev1.record(stream)
                    torch.cuda.synchronize()
                    ms5.append(ev0.elapsed_time(ev1))
                sys5, sa5 = C.c_void_p(), _lib.MagStats()
                _lib.check(lib.mag_assemble(ctx.handle, C.byref(view5), C.byref(mat), C.byref(opt), C.byref(sys5), C.byref(sa5)),
                           "mag_assemble(config 5 unit)")
                rr5, bb5 = C.c_double(), C.c_double()
                rc5 = lib.mag_system_residual(sys5, out5["ux"].data_ptr(), out5["uy"].data_ptr(), 1, C.byref(rr5), C.byref(bb5))
                lib.mag_system_free(sys5)
                _lib.check(rc5, "mag_system_residual(config 5 unit)")
                true5 = (rr5.value / bb5.value) ** 0.5 if bb5.value > 0 else 0.0
                step5 = statistics.mean(ms5[1:])
                example["config5_perforated_unit"] = {
                    "workload": "perforated (holes of radius 16h at pitch 64h) " + workload_name(args.nx, args.ny)
                                + f" before the holes; {e5} triangles, {int(st5.n_free)} free DOF: the per-GPU unit of the weak-scaling job",
                    "ms_per_step": step5, "melem_s": e5 / (step5 * 1e-3) / 1e6, "pcg_iters": int(st5.iters),
                    "precond_used": int(st5.precond_used), "pcg_time_to_solve_s": st5.ms_solve * 1e-3,
                    "assembly_ms": st5.ms_elem + st5.ms_sort + st5.ms_reduce + st5.ms_bc,
                    "true_rel_residual": true5}
                if not (true5 <= MAX_TRUE_REL_RESIDUAL):
                    failures.append(f"config 5 unit: true relative residual {true5:.3e} > {MAX_TRUE_REL_RESIDUAL}")
                del out5
            except Exception as err:
                example["config5_perforated_unit"] = {"error": str(err)}
            finally:
                if dm5:
                    lib.mag_devmesh_free(dm5)

    # ---- end to end: pinned host buffers through mag_solve ----------------------------------
    e2e = None
    barrier()
    if not args.no_e2e:
        host = meshgen.perforated_plate(nx, ny) if weak else meshgen.plate(nx, ny)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        hx, hy = pin(host.x), pin(host.y)
        h0, h1, h2 = (pin(a.view(np.int32)) for a in (host.n0, host.n1, host.n2))
        hux, huy, hfx, hfy, hk = pin(host.ux), pin(host.uy), pin(host.fx), pin(host.fy), pin(host.known)
        hm = _lib.MagMesh(N, E, hx.data_ptr(), hy.data_ptr(), h0.data_ptr(), h1.data_ptr(), h2.data_ptr(),
                          hux.data_ptr(), huy.data_ptr(), hfx.data_ptr(), hfy.data_ptr(), hk.data_ptr(), 0)
        out_h = {k: torch.empty(N, dtype=torch.float64).pin_memory() for k in ("ux", "uy", "fx", "fy")}
        out_h["stress"] = torch.empty(E, dtype=torch.float64).pin_memory()
        res_h = _lib.MagResult(out_h["ux"].data_ptr(), out_h["uy"].data_ptr(), out_h["fx"].data_ptr(),
                               out_h["fy"].data_ptr(), out_h["stress"].data_ptr(), None, 0)
        h2d = N * (8 * 6 + 1) + E * 12                   # whole job (at N > 1 every rank moves 1/N of it)
        d2h = N * 32 + E * 8

        # N > 1: every rank passes the whole host mesh (the library moves each array over PCIe once per job: every rank
        # uploads its 1/N slice, the slices travel over NVLink) and reads back its 1/N slice of the result arrays
        # (mag_options.result_scope = 1): the job's inputs and results cross PCIe once, like at N = 1.
        opt_e = options(result_scope=1) if world > 1 else opt

        def host_step():
            s = _lib.MagStats()
            _lib.check(lib.mag_solve(ctx.handle, C.byref(hm), C.byref(mat), C.byref(opt_e), C.byref(res_h), C.byref(s)),
                       "mag_solve(host)")
            return s

        host_step()
        barrier()
        e_steps = max(1, min(args.steps, 5))
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for _ in range(e_steps):
                host_step()
            ev1.record(stream)
        barrier()
        ms_e2e = max_over_ranks(ev0.elapsed_time(ev1)) / e_steps
        # what came back over PCIe is what the device-resident step computed
        if world > 1:                                    # this rank's slice of the result, bit for bit
            nlo, nhi = mdist.partition_nodes(N, world, rank)
            elo, ehi = mdist.partition_nodes(E, world, rank)
            ok = bool(torch.equal(out_h["ux"][nlo:nhi], out_d["ux"][nlo:nhi].cpu()) and
                      torch.equal(out_h["fy"][nlo:nhi], out_d["fy"][nlo:nhi].cpu()) and
                      torch.equal(out_h["stress"][elo:ehi], out_d["stress"][elo:ehi].cpu()))
            same = sum_over_ranks([0.0 if ok else 1.0])[0] == 0.0
        else:
            same = bool(torch.equal(out_h["ux"], out_d["ux"].cpu()) and torch.equal(out_h["stress"], out_d["stress"].cpu()))
        if not same:
            failures.append("the host-buffer step did not return the device-resident step's result bit for bit")
        e2e = {"value": E / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "steps": e_steps,
               "host_result_bit_identical_to_device_step": same,
               "scope": "every rank uploads 1/N of each input array (NVLink allgather after) and reads back its 1/N slice "
                        "of the results" if world > 1 else "one rank: everything"}

    # ---- CPU baseline beside it ----------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        _, _, cpu = cpu_baseline_block(args, world, 3, 1)

    # device memory this process holds at the end of the run: the library's heap keeps its high-water mark reserved,
    # so this is the peak of the step plus the bench's own result / proof buffers and the CUDA context.
    # (max_over_ranks is a collective: every rank calls it, BEFORE the ranks part ways below.)
    free_b, total_b = torch.cuda.mem_get_info(dev)
    mem_gb = max_over_ranks((total_b - free_b) / 1e9)

    if rank != 0:
        barrier()
        lib.mag_devmesh_free(dm)
        ctx.close()
        tdist.destroy_process_group()
        return
    its = max(int(last.iters), 1)
    iter_bytes = nbytes.value + 88 * n_local_rows
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_of(args, world),
        "details": {"solver": f"{precond_names.get(default_precond, default_precond)}, rel_tol 1e-9, SELL-32 SpMV with "
                              f"{int(last.sell_index_bits)}-bit column indices",
                    "multi_gpu": None if world == 1 else "P2P halo stores fused into the CG update kernel, dot products "
                                 + ("allreduced by NCCL" if args.allreduce else "through peer-memory mailboxes inside the CG kernels"),
                    "triangles": E, "n_free": int(last.n_free), "nnz": int(last.nnz), "nnz_structural": int(last.nnz_structural),
                    "n_coarse": int(last.n_coarse), "device_mem_gb_per_gpu_max": mem_gb},
        "proof": dict(proof, multi_gpu_vs_emulation=multi, failures=failures),
        "metrics": {"assembly_melem_s": E / (ms_asm * 1e-3) / 1e6, "assembly_ms": ms_asm,
                    "pcg_time_to_solve_s": ms_solve * 1e-3, "pcg_iters": int(last.iters),
                    "pcg_final_rel_residual": last.final_residual / last.b_norm if last.b_norm else 0.0,
                    "coarse_setup_s_inside_solve": last.ms_coarse_setup * 1e-3,
                    "ms_per_iteration": ms_solve / its,
                    "spmv_hbm_gbs": achieved, "ms_upload": ms_upload, "ms_elem": last.ms_elem, "ms_sort": last.ms_sort,
                    "ms_reduce": last.ms_reduce, "ms_bc": last.ms_bc, "ms_format": last.ms_format,
                    "ms_post": ms_post, "assembly_variants": assemblies, "other_preconditioners": solves,
                    "same_config_sample": same_cfg, "baseline_configs": example,
                    # device-side timeline of one PCG iteration on rank 0 (%globaltimer, us): update_p + launch gap,
                    # spmv, gap, wait for global p.q, update_xr (incl. wait), gap, wait for global r.z
                    "pcg_iteration_timeline_us": [round(v / 1e3, 2) for v in list(last.prof)[:7]]},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    if default_precond == 1:
        line["roofline"]["pcg_iteration"] = {"bytes": iter_bytes, "ms": ms_solve / its,
                                             "achieved": iter_bytes / (ms_solve / its * 1e-3) / 1e9,
                                             "frac": iter_bytes / (ms_solve / its * 1e-3) / 1e9 / peak}
    print(json.dumps(line), flush=True)
    if world > 1:
        barrier()
    lib.mag_devmesh_free(dm)
    ctx.close()
    if tdist is not None:
        tdist.destroy_process_group()
    if failures:
        print("bench.py: CORRECTNESS PROOF FAILED: " + "; ".join(failures), file=sys.stderr, flush=True)
        sys.exit(3)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
