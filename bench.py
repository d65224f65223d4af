#!/usr/bin/env python
"""bench.py -- p121 EBE-PCG throughput on B200 (BASELINE.json metric).

A *step* is one PCG iteration of p121.f90:91-103 (gather -> storkm mat-vec -> scatter ->
dot products / vector updates -> checon_par) over the whole mesh.  Workload at any N:
BASELINE config C, the 125^3 20-node-hexahedra cube (1 953 125 elements, 23 531 000
equations, 56.25 GB of storkm), partitioned over the N ranks with ParaFEM's own partition
(strong scaling).  `value` = MDOF*iterations/s = neq * K / t / 1e6 with everything resident
in HBM; `e2e` = the same through pf_pcg_solve with pinned HOST buffers for r_pp and xnew_pp.

  python bench.py [--gpus N --steps K --warmup W]            this repo's CUDA path
  python bench.py --impl reference [...]                     the CPU restatement of the reference
                                                              (oracle/, all host threads)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "p121 EBE-PCG MDOF-iters/s"
UNIT = "MDOF*iterations/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # (option names must not be prefixes of torchrun's own options: argparse abbreviation matching
    #  rejects e.g. --n / --nod even after the script name)
    ap.add_argument("--cube", dest="n", type=int, default=125, help="cube edge in elements (125 = BASELINE config C)")
    ap.add_argument("--hex", dest="nod", type=int, default=20, choices=[8, 20], help="nodes per brick")
    ap.add_argument("--program", default="p121", choices=["p121", "p123", "p124", "p125"],
                    help="p123 = steady heat conduction, 8-node bricks (BASELINE config B at --cube 100); p124 = transient "
                         "conduction, one PCG solve per time step (a step = one time step); p125 = explicit transient "
                         "conduction (a step = one pass of gather / mat-vec / scatter)")
    ap.add_argument("--cpu-n", type=int, default=40, help="cube edge of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=60)
    ap.add_argument("--matrix-free", type=int, default=0, choices=[0, 1, 2],
                    help="BASELINE config E: 1 = rebuild the element operator from coordinates every iteration, "
                         "2 = same with stored geometric factors (FP64-pipe roofline)")
    ap.add_argument("--layout", type=int, default=0, choices=[0, 1],
                    help="storkm layout of the main measurement: 0 = the reference's storkm_pp (headline), "
                         "1 = packed lower triangles (half the stream)")
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling: n x (n*N) x n elements, i.e. one n^3 slab of y-planes per GPU")
    ap.add_argument("--no-variants", dest="variants", action="store_false",
                    help="skip the two matrix-free variants of the same workload (key `variants`)")
    ap.add_argument("--no-solve", action="store_true", help="skip the solve to convergence (time-to-solution)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return None
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_leg(n, nod, steps, warmup):
    """The oracle (CPU restatement of the reference, kind 'port') on a bounded sample."""
    import oracle
    from parafem_b200 import host
    prob = host.cube_p121(n, n, n, nod)
    t = time.time()
    km = oracle.form_km_elastic(prob.g_coord_pp, prob.nod, prob.nip, prob.e, prob.v)
    t_km = time.time() - t
    cores = oracle.max_threads()
    if warmup > 0:
        oracle.pcg(km, prob.g_g_pp, prob.neq, prob.r_pp, -1.0, warmup, npes=cores, red_mode=0)
    res = oracle.pcg(km, prob.g_g_pp, prob.neq, prob.r_pp, -1.0, steps, npes=cores, red_mode=0)
    secs = res["seconds"]
    val = prob.neq * res["iters"] / secs / 1e6
    return {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{res['iters']} PCG iterations of p121 on a {n}^3 hex{nod} cube ({prob.nels} elements, "
                      f"{prob.neq} equations, storkm {km.nbytes / 1e9:.2f} GB) after {warmup} warm-up iterations; "
                      f"C restatement of the reference (oracle/pf_oracle.c, OpenMP over {cores} emulated ranks); "
                      f"the Fortran+MPI reference itself cannot be built in this image",
            "seconds": secs, "ms_per_step": 1e3 * secs / res["iters"], "storkm_form_seconds": t_km,
            "neq": prob.neq, "steps": res["iters"]}


def run_reference(args, rank):
    if rank != 0:
        return
    c = cpu_leg(args.cpu_n, args.nod, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": c["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": c["steps"], "warmup": args.warmup, "ms_per_step": c["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"p121 {args.n}^3 hex{args.nod} cube (BASELINE config C), timed on the bounded sample "
                                   f"named in cpu_baseline.sample", "step": "one PCG iteration"},
            "cpu_baseline": {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": c["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def transient_bench(args, rank, nranks, local, nccl_id, barrier, maxf):
    """the drivers that reuse the three kernels.  p124: K time steps, each one resident PCG solve
    (value = neq * PCG iterations / s); p125: K explicit steps (value = neq * steps / s).  Same JSON keys as the
    headline line; not a BASELINE config, so `vs_baseline` is null and there is no cpu_baseline leg."""
    from parafem_b200 import host, solver
    n, K, W = args.n, args.steps, max(args.warmup, 3)
    nye = n * nranks if args.weak else n
    make = host.cube_p124 if args.program == "p124" else host.cube_p125
    prob = make(n, nye, n, npes=nranks, numpe=rank + 1)
    s = solver.Solver(rank, nranks, local, nccl_id)
    solver.setup_problem(s, prob)
    pk, pk_kind = peaks()
    sampler = ClockSampler(local)

    def run(k):
        """k steps -> (device ms, PCG iterations or steps)."""
        if args.program == "p125":
            return s.explicit_steps(k), k
        ms_, its_ = 0.0, 0
        for _ in range(k):
            it, _, m = s.transient_step(prob.tol, prob.limit)
            ms_ += m
            its_ += it
        return ms_, its_

    if args.program == "p124":
        s.transient_start(prob.val0)
    else:
        s.explicit_start(prob.val0)
    run(W)
    barrier()
    l0 = s.kernel_launches()
    if rank == 0:
        sampler.start()
    ms, units = run(K)                       # timed region: resident, graph replay of the PCG iteration (p124)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = s.kernel_launches() - l0
    ms = maxf(ms)
    t = time.perf_counter()
    _, e_units = run(K)
    x = s.pcg_get_x()                        # e2e: the same through the C-ABI + the field read back to the host
    e2e_s = maxf(time.perf_counter() - t)
    s.set_profile(True); s.reset_profile()   # a third pass with CUDA events around every launch: kernel shares
    run(K)
    km = {name: s.kernel_ms(i) for i, name in enumerate(("matvec", "scatter", "vector", "halo"))}
    s.set_profile(False)
    mv_ms, mv_n = km["matvec"]
    bytes_pp = prob.nels_pp * 64 * 8
    achieved = bytes_pp / (mv_ms / max(mv_n, 1) / 1e3) / 1e9 if mv_n else None
    if rank == 0:
        line = {"metric": f"{args.program} EBE MDOF-" + ("iters/s" if args.program == "p124" else "steps/s"),
                "value": prob.neq * units / (ms / 1e3) / 1e6, "unit": UNIT if args.program == "p124" else "MDOF*steps/s",
                "n_gpus": nranks, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
                "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{args.program} {n}x{nye}x{n} hex8 box, p12meshgen geometry: {prob.nels} elements, "
                                       f"{prob.neq} equations, dtim {prob.dtim}",
                           "step": "one time step = right-hand side product + one PCG solve to tol 1e-4" if args.program == "p124"
                                   else "one explicit step (gather, store_pm mat-vec, scatter, scale by the lumped mass)",
                           "l2": "element matrices per rank " + f"{bytes_pp / 1e6:.0f} MB" + (" (larger than L2)" if bytes_pp > 126e6 else " (fits L2)"),
                           "nels": prob.nels, "neq": prob.neq},
                "pcg_iterations_timed": units if args.program == "p124" else None,
                "e2e": {"value": prob.neq * e_units / e2e_s / 1e6, "seconds": e2e_s,
                        "unit": UNIT if args.program == "p124" else "MDOF*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": prob.neq_pp * 8 / K,
                        "note": "K steps through the C-ABI + the field read back to the host"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved / pk["hbm_gbs"] if achieved else None, "traffic": None,
                             "kernel": "k_matvec<8> (storka / store_pm stream; gather fused)", "peak_kind": pk_kind,
                             "algorithmic_bytes_per_launch": bytes_pp, "launches_timed": int(mv_n)},
                "kernel_ms_per_step": {k: v[0] / K for k, v in km.items()},
                "clocks": clocks, "x_checksum": float(x.sum())}
        print(json.dumps(line), flush=True)
    s.close()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from parafem_b200 import host, solver

    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"    # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    if world > 1:
        # control plane only (id broadcast, barriers, max over ranks); the data path is the
        # library's own NCCL communicator
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    nranks = world

    def barrier():
        if nranks > 1:
            dist.barrier()

    def maxf(x):
        if nranks == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nccl_id = None
    if nranks > 1:
        idt = torch.from_numpy(solver.nccl_unique_id().copy()) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(idt, src=0)
        nccl_id = idt.numpy()

    t_setup0 = time.time()
    n = args.n
    nye = n * nranks if args.weak else n
    if args.program in ("p124", "p125"):
        transient_bench(args, rank, nranks, local, nccl_id, barrier, maxf)
        if nranks > 1:
            dist.destroy_process_group()
        return
    if args.program == "p123":
        args.nod = 8
        prob = host.cube_p123(n, nye, n, aa=1.0 / n, bb=1.0 / n, cc=1.0 / n, limit=20000, npes=nranks, numpe=rank + 1)
    else:
        prob = host.cube_p121(n, nye, n, args.nod, aa=10.0 / n, bb=10.0 / n, cc=10.0 / n, limit=20000, npes=nranks,
                              numpe=rank + 1)
    t_mesh = time.time() - t_setup0
    s = solver.Solver(rank, nranks, local, nccl_id)
    t0 = time.time()
    solver.setup_problem(s, prob, matrix_free=args.matrix_free, layout=args.layout)
    barrier()
    t_dev_setup = time.time() - t0
    fp64_tflops = s.measure_fp64() if args.matrix_free else None
    hbm_read_gbs = s.measure_hbm_read() if not args.matrix_free else None   # read-only stream ceiling (informational)
    ntot = prob.ntot
    storkm_bytes_pp = prob.nels_pp * (ntot * (ntot + 1) // 2 if args.layout == 1 else ntot * ntot) * 8

    K, W = args.steps, max(args.warmup, 3)
    import ctypes as C
    from parafem_b200._lib import lib, ptr
    r_pin = torch.empty(prob.neq_pp, dtype=torch.float64).pin_memory()
    x_pin = torch.empty(prob.neq_pp, dtype=torch.float64).pin_memory()
    r_np, x_np = r_pin.numpy(), x_pin.numpy()
    r_np[:] = prob.r_pp
    it_c, cv_c = C.c_int(), C.c_int()

    def measure():
        """W warm-up + K timed iterations device-resident (`value`), then the same through
        pf_pcg_solve with pinned host buffers (`e2e`)."""
        s.pcg_load_rhs(prob.r_pp)
        s.pcg_run(-1.0, W)                       # W untimed warm-up steps (tol < 0: never converges)
        s.set_profile(True)
        s.reset_profile()
        sampler = ClockSampler(local)
        barrier()
        launches0 = s.kernel_launches()
        if rank == 0:
            sampler.start()
        iters, _, ms = s.pcg_run(-1.0, K)        # exactly K steps, CUDA events on the solver stream
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        launches = s.kernel_launches() - launches0
        assert iters == K
        ms = maxf(ms)
        km = {name: s.kernel_ms(i) for i, name in enumerate(("matvec", "scatter", "vector", "halo"))}
        s.set_profile(False)

        def e2e_once(k):
            barrier()
            t = time.perf_counter()
            rc = lib().pf_pcg_solve(s._h, ptr(r_np), -1.0, k, ptr(x_np), C.byref(it_c), C.byref(cv_c))
            assert rc == 0 and it_c.value == k
            chk = float(x_np[0])                 # the step's result is read on the host
            return maxf(time.perf_counter() - t), chk

        e2e_once(W)
        e2e_s, _ = e2e_once(K)
        return {"ms": ms, "value": prob.neq * K / (ms / 1e3) / 1e6, "launches": launches, "clocks": clocks,
                "kernel_ms": km, "e2e_s": e2e_s, "e2e_value": prob.neq * K / e2e_s / 1e6}

    def solve_to_convergence():
        s.pcg_load_rhs(prob.r_pp)
        barrier()
        it_full, conv, ms_full = s.pcg_run(prob.tol, prob.limit)
        ms_full = maxf(ms_full)
        x = s.pcg_get_x()
        return {"iters": it_full, "converged": bool(conv), "solve_s": ms_full / 1e3,
                "mdof_iters_per_s": prob.neq * it_full / (ms_full / 1e3) / 1e6,
                "x1": float(x[0]) if rank == 0 else None, "tol": prob.tol}

    def mf_flops(mode, nodn):
        # flops per element of k_apply_mf as executed (fma = 2 flop, mul/add = 1), 8 Gauss points:
        # phase 1  H = der x p: 9 fma per node and point; per point G (9 mul + 18 fma), eps (3 add),
        # sigma (12 mul + 30 fma), T (9 mul + 18 fma) = 165 flop; phase 2: per dof 1 mul + 23 fma.
        # Mode 1 adds the Jacobian pass: 9 fma per node and point + det/inverse/det*w (51 flop) per point.
        fl = 8 * 18 * nodn + 8 * 165 + 3 * nodn * 47
        if mode == 1:
            fl += 8 * 18 * nodn + 8 * 51
        return fl

    def mf_roofline(mode, m, peak):
        mv_ms_, mv_n_ = m["kernel_ms"]["matvec"]
        avg = mv_ms_ / max(mv_n_, 1)
        fl = prob.nels_pp * mf_flops(mode, args.nod)
        return {"bound": "fp64", "achieved": fl / (avg / 1e3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                "frac": fl / (avg / 1e3) / 1e12 / peak, "traffic": None,
                "kernel": "k_apply_mf (matrix-free operator form, gather fused)",
                "peak_kind": "DFMA micro-benchmark on this device (pf_measure_fp64), fma = 2 flop",
                "algorithmic_flops_per_launch": fl, "flops_per_element": mf_flops(mode, args.nod),
                "avg_launch_ms": avg, "launches_timed": int(mv_n_)}

    m = measure()
    ms, value, launches, clocks = m["ms"], m["value"], m["launches"], m["clocks"]
    (mv_ms, mv_n), (sc_ms, sc_n), (vec_ms, vec_n), (halo_ms, halo_n) = (m["kernel_ms"][k] for k in ("matvec", "scatter", "vector", "halo"))
    e2e_s, e2e_value = m["e2e_s"], m["e2e_value"]

    # ---- solve to convergence: time-to-solution, iteration count ---------------------------
    tts = None if args.no_solve else solve_to_convergence()

    pk, pk_kind = peaks()
    mv_avg_ms = mv_ms / max(mv_n, 1)
    achieved = storkm_bytes_pp / (mv_avg_ms / 1e3) / 1e9 if mv_n else None
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = f"hex{args.nod}_n{n}_gpus{nranks}" if args.program == "p121" else f"{args.program}_n{n}_gpus{nranks}"
        traffic = tj.get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- the matrix-free variants of the same workload (BASELINE config E), same process ----------
    variants = None
    if args.variants and args.program == "p121" and not args.matrix_free and args.layout == 0:
        variants = {}
        peak = None
        # packed lower triangles (pf_set_storkm_layout 1): same algorithm and summation order, half the stream
        solver.setup_problem(s, prob, layout=1)
        mm = measure()
        sym_bytes = prob.nels_pp * (ntot * (ntot + 1) // 2) * 8
        sv_ms, sv_n = mm["kernel_ms"]["matvec"]
        sv_avg = sv_ms / max(sv_n, 1)
        variants["stored_symmetric_packed"] = {
            "value": mm["value"], "unit": UNIT, "ms_per_step": mm["ms"] / K, "e2e": {"value": mm["e2e_value"], "unit": UNIT},
            "kernel_ms_per_step": {k: mm["kernel_ms"][k][0] / K for k in mm["kernel_ms"]},
            "roofline": {"bound": "hbm", "achieved": sym_bytes / (sv_avg / 1e3) / 1e9, "peak": peaks()[0]["hbm_gbs"],
                         "unit": "GB/s", "frac": sym_bytes / (sv_avg / 1e3) / 1e9 / peaks()[0]["hbm_gbs"], "traffic": None,
                         "kernel": "k_matvec_sym (packed lower triangles, gather fused)",
                         "algorithmic_bytes_per_launch": sym_bytes, "avg_launch_ms": sv_avg, "launches_timed": int(sv_n)},
            "gpu_launches": int(mm["launches"]),
            "note": "K(i,j) = L(max,min): bit-identical to MATMUL on the symmetrised matrices; not the headline "
                    "because the reference's storkm_pp is only symmetric to rounding"}
        if not args.no_solve:
            variants["stored_symmetric_packed"]["time_to_solution"] = solve_to_convergence()
        for mode, name in ((2, "matrix_free_geometric_factors"), (1, "matrix_free_rebuilt_from_coordinates")):
            solver.setup_problem(s, prob, matrix_free=mode)
            peak = peak or s.measure_fp64()
            mm = measure()
            v = {"value": mm["value"], "unit": UNIT, "ms_per_step": mm["ms"] / K,
                 "e2e": {"value": mm["e2e_value"], "unit": UNIT},
                 "kernel_ms_per_step": {k: mm["kernel_ms"][k][0] / K for k in mm["kernel_ms"]},
                 "roofline": mf_roofline(mode, mm, peak), "gpu_launches": int(mm["launches"])}
            if mode == 2 and not args.no_solve:
                v["time_to_solution"] = solve_to_convergence()
            variants[name] = v

    cpu = None
    if rank == 0 and nranks == 1 and not args.no_cpu and args.program == "p121":
        cpu = cpu_leg(args.cpu_n, args.nod, args.cpu_steps, 3)

    if rank == 0:
        line = {
            "metric": METRIC if args.program == "p121" else "p123 EBE-PCG MDOF-iters/s", "value": value, "unit": UNIT, "n_gpus": nranks, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak" if args.weak else "strong",
            "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.program} {n}x{nye}x{n} hex{args.nod} cube, p12meshgen geometry (BASELINE config "
                                   f"{'C' if (n, args.nod) == (125, 20) else 'custom'}): {prob.nels} elements, "
                                   f"{prob.neq} equations, storkm {prob.nels * ntot * ntot * 8 / 1e9:.2f} GB",
                       "step": "one PCG iteration (gather, storkm mat-vec, scatter, dots/updates, checon_par)",
                       "partition": f"ParaFEM calc_nels_pp/calc_neq_pp over {nranks} rank(s)",
                       "l2": "inputs larger than L2 (storkm stream per rank >> 126 MB), no flush needed",
                       "nels": prob.nels, "neq": prob.neq, "nod": args.nod, "nip": prob.nip},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": prob.neq_pp * 8 / K,
                    "d2h_bytes_per_step": prob.neq_pp * 8 / K, "seconds": e2e_s,
                    "note": "pf_pcg_solve with pinned host r_pp in / xnew_pp out, K iterations per call"},
            "gpu_launches": int(launches),
            "roofline": ({"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                          "frac": (achieved / pk["hbm_gbs"]) if achieved else None, "traffic": traffic,
                          "kernel": ("k_matvec_sym" if args.layout == 1 else "k_matvec") + " (storkm stream; gather fused)", "peak_kind": pk_kind,
                          "read_only_stream_gbs": hbm_read_gbs,
                          "frac_of_read_only_stream": (achieved / hbm_read_gbs) if (achieved and hbm_read_gbs) else None,
                          "algorithmic_bytes_per_launch": storkm_bytes_pp, "avg_launch_ms": mv_avg_ms,
                          "launches_timed": int(mv_n)} if not args.matrix_free else
                         mf_roofline(args.matrix_free, m, fp64_tflops)),
            "storkm_layout": "packed lower triangles" if args.layout == 1 else "storkm_pp(ntot,ntot,nels_pp) as the reference",
            "variant": {0: "stored storkm", 1: "matrix-free, rebuilt from coordinates (config E)",
                        2: "matrix-free, stored geometric factors"}[args.matrix_free],
            "kernel_ms_per_step": {"matvec": mv_ms / K, "scatter": sc_ms / K, "vector_and_reductions": vec_ms / K,
                                   "halo": halo_ms / K},
            "clocks": clocks,
            "time_to_solution": tts,
            "variants": variants,
            "setup_s": {"mesh_host": t_mesh, "device_setup_incl_storkm": t_dev_setup},
            "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
        }
        print(json.dumps(line), flush=True)
    s.close()
    if nranks > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
