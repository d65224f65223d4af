#!/usr/bin/env python
"""bench.py -- p121 EBE-PCG throughput on B200 (BASELINE.json metric).

A *step* is one PCG iteration of p121.f90:91-103 (gather -> storkm mat-vec -> scatter ->
dot products / vector updates -> checon_par) over the whole mesh.  Workload of the headline
(`value`, `e2e`, `roofline`) at any N: BASELINE config C, the 125^3 20-node-hexahedra cube
(1 953 125 elements, 23 531 000 equations, 56.25 GB of storkm), partitioned over the N ranks with
ParaFEM's own partition (strong scaling).  `value` = MDOF*iterations/s = neq * K / t / 1e6 with
everything resident in HBM; `e2e` = the same through pf_pcg_solve with pinned HOST buffers for
r_pp and xnew_pp.

The default run also carries, in the same JSON line:
  parity_check          an N-rank solve of two small meshes compared bit for bit with the oracle
                        emulating the same N ranks (before anything is timed)
  configs               BASELINE config B (p123 100^3, N = 1) and D (p121 hex8 200^3, every N)
  weak                  100^3 hex20 elements PER GPU (weak scaling)
  variants              packed lower triangles and the two matrix-free modes (config E) of config C
  reference_gpu_kernel  the reference's own CUDA mat-vec (xx3/cuda_helpers.cu, oracle/_ref) timed on the
                        same device buffers beside this library's (N = 1)
  cpu_baseline          the oracle on a bounded sample on the host cores (N = 1)

  python bench.py [--gpus N --steps K --warmup W]            this repo's CUDA path
  python bench.py --impl reference [...]                     the CPU restatement of the reference
                                                              (oracle/ only, all host cores, same workload
                                                              where the host RAM admits it)
"""
import argparse
import ctypes as C
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
KERNEL_KINDS = ("matvec", "scatter", "vector", "halo")


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
    ap.add_argument("--program", default="p121", choices=["p121", "p123", "p124", "p125", "p1210"],
                    help="p123 = steady heat conduction, 8-node bricks (BASELINE config B at --cube 100); p124 = transient "
                         "conduction, one PCG solve per time step (a step = one time step); p125 = explicit transient "
                         "conduction (a step = one pass of gather / mat-vec / scatter)")
    ap.add_argument("--cpu-n", type=int, default=0,
                    help="cube edge of the CPU run: --impl reference runs the GPU arm's own cube unless this is given "
                         "(or the host RAM does not admit it); the cpu_baseline leg of the GPU arm uses a bounded sample")
    ap.add_argument("--cpu-steps", type=int, default=30)
    ap.add_argument("--matrix-free", type=int, default=0, choices=[0, 1, 2],
                    help="BASELINE config E: 1 = rebuild the element operator from coordinates every iteration, "
                         "2 = same with stored geometric factors (FP64-pipe roofline)")
    ap.add_argument("--layout", type=int, default=0, choices=[0, 1],
                    help="storkm layout of the main measurement: 0 = the reference's storkm_pp (headline), "
                         "1 = packed lower triangles (half the stream)")
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling: n x (n*N) x n elements, i.e. one n^3 slab of y-planes per GPU")
    ap.add_argument("--no-variants", dest="variants", action="store_false",
                    help="skip the packed / matrix-free variants of the same workload (key `variants`)")
    ap.add_argument("--no-solve", action="store_true", help="skip the solves to convergence (time-to-solution)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", dest="extra", action="store_false",
                    help="skip parity_check, configs B / D, the weak block and reference_gpu_kernel")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 0


def workload_config(program, nx, ny, nz, nod, nels, neq, nip):
    """`config` of the JSON line: the linear system, not who solves it -- both arms print the same dict for the
    same cube."""
    ntot = nod * (3 if program == "p121" else 1)
    tag = {("p121", 125, 20): "C", ("p123", 100, 8): "B", ("p121", 200, 8): "D"}.get((program, nx, nod) if nx == ny == nz else None, "custom")
    return {"workload": f"{program} {nx}x{ny}x{nz} hex{nod} cube, p12meshgen geometry (BASELINE config {tag}): {nels} elements, "
                        f"{neq} equations, storkm {nels * ntot * ntot * 8 / 1e9:.2f} GB",
            "step": "one PCG iteration (gather, storkm mat-vec, scatter, dots/updates, checon_par)",
            "l2": "inputs larger than L2 (storkm stream per rank >> 126 MB), no flush needed",
            "nels": nels, "neq": neq, "nod": nod, "nip": nip}


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


# --------------------------------------------------------------------------------------------
# the CPU side: oracle/ only (never the product's library)
# --------------------------------------------------------------------------------------------
def cpu_bytes_needed(program, n, nod):
    nels = n ** 3
    ntot = nod * (3 if program == "p121" else 1)
    neq = 3.1 * (2 * n + 1) ** 3 / 2 if nod == 20 else (n + 1) ** 3 * (3 if program == "p121" else 1)
    return nels * ntot * ntot * 8 + 3 * nels * ntot * 8 + nels * nod * (24 + 8) + 12 * neq * 8 + (3 << 30)


def cpu_leg(program, n, nod, steps, warmup):
    """The oracle (C restatement of the reference, kind 'port') on an n^3 cube, every host core as one emulated MPI
    rank.  The mesh comes from the oracle's own p12meshgen restatement: this leg never loads libparafem_b200.so."""
    import oracle
    cores = oracle.use_all_cores()          # NOT OMP_NUM_THREADS: torchrun exports 1
    t = time.time()
    if program == "p123":
        m = oracle.cube_p123(n, n, n, aa=1.0 / n, bb=1.0 / n, cc=1.0 / n, limit=20000)
        km = oracle.form_kc_laplace(m.g_coord_pp, m.nip, m.kx, m.ky, m.kz)
    else:
        m = oracle.cube_p121(n, n, n, nod, aa=10.0 / n, bb=10.0 / n, cc=10.0 / n, limit=20000)
        km = oracle.form_km_elastic(m.g_coord_pp, m.nod, m.nip, m.e, m.v)
    t_setup = time.time() - t
    if warmup > 0:
        oracle.pcg(km, m.g_g_pp, m.neq, m.r_pp, -1.0, warmup, npes=cores, red_mode=0)
    res = oracle.pcg(km, m.g_g_pp, m.neq, m.r_pp, -1.0, steps, npes=cores, red_mode=0)
    secs = res["seconds"]
    val = m.neq * res["iters"] / secs / 1e6
    return {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{res['iters']} PCG iterations of {program} on a {n}^3 hex{nod} cube ({m.nels} elements, "
                      f"{m.neq} equations, storkm {km.nbytes / 1e9:.2f} GB) after {warmup} warm-up iterations; "
                      f"C restatement of the reference (oracle/pf_oracle.c, OpenMP over {cores} emulated MPI ranks = the host "
                      f"cores this process may use); the Fortran+MPI reference itself cannot be built in this image "
                      f"(no gfortran, no MPI)",
            "seconds": secs, "ms_per_step": 1e3 * secs / res["iters"], "mesh_and_storkm_seconds": t_setup,
            "neq": m.neq, "nels": m.nels, "steps": res["iters"], "n": n, "nip": m.nip,
            "hbm_equivalent_gbs": km.nbytes * res["iters"] / secs / 1e9}


def run_reference(args, rank):
    """`--impl reference`: the reference's algorithm on the host cores (rank 0 alone under torchrun), on the GPU arm's
    own workload when the host RAM admits its storkm_pp, else on the largest smaller cube -- config.workload says
    which was really run."""
    if rank != 0:
        return
    program = args.program if args.program in ("p121", "p123") else "p121"
    nod = 8 if program == "p123" else args.nod
    avail = mem_available_bytes()
    if args.cpu_n:
        n = args.cpu_n
    else:
        n = next((c for c in [args.n] + [c for c in (100, 80, 64, 50, 40, 32, 20) if c < args.n]
                  if not avail or cpu_bytes_needed(program, c, nod) < 0.92 * avail), 20)
    c = cpu_leg(program, n, nod, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC if program == "p121" else "p123 EBE-PCG MDOF-iters/s",
            "value": c["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": c["steps"], "warmup": args.warmup, "ms_per_step": c["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(program, n, n, n, nod, c["nels"], c["neq"], c["nip"]),
            "same_workload_as_gpu_arm": n == args.n,
            "host": {"cores": c["cores"], "mem_available_gb": avail / 1e9, "storkm_stream_gbs": c["hbm_equivalent_gbs"],
                     "mesh_and_storkm_seconds": c["mesh_and_storkm_seconds"]},
            "cpu_baseline": {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": c["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# the GPU side
# --------------------------------------------------------------------------------------------
class Ctx:
    """One rank of the GPU arm: handle, control plane (gloo), pinned staging."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from parafem_b200 import solver
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.nranks = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"    # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        if self.nranks > 1:
            # control plane only (id broadcast, barriers, max over ranks); the data path is the library's own
            # NCCL communicator / peer mappings
            dist.init_process_group(backend="gloo", rank=self.rank, world_size=self.nranks)
        assert self.nranks == args.gpus or self.nranks == 1, "launch with torchrun --nproc-per-node N for --gpus N"
        nccl_id = None
        if self.nranks > 1:
            idt = torch.from_numpy(solver.nccl_unique_id().copy()) if self.rank == 0 else torch.zeros(128, dtype=torch.uint8)
            dist.broadcast(idt, src=0)
            nccl_id = idt.numpy()
        self.s = solver.Solver(self.rank, self.nranks, self.local, nccl_id)
        self.pk, self.pk_kind = peaks()
        self.K, self.W = args.steps, max(args.warmup, 3)
        self._pin = {}
        try:
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            self.traffic = {}

    def barrier(self):
        if self.nranks > 1:
            self.dist.barrier()

    def maxf(self, x):
        if self.nranks == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok):
        if self.nranks == 1:
            return bool(ok)
        t = self.torch.tensor([0 if ok else 1], dtype=self.torch.int64)
        self.dist.all_reduce(t)
        return int(t.item()) == 0

    def pinned(self, n):
        if n not in self._pin:
            self._pin[n] = (self.torch.empty(max(n, 1), dtype=self.torch.float64).pin_memory().numpy()[:n],
                            self.torch.empty(max(n, 1), dtype=self.torch.float64).pin_memory().numpy()[:n])
        return self._pin[n]

    def close(self):
        self.s.close()
        if self.nranks > 1:
            self.dist.destroy_process_group()


def make_problem(ctx, program, n, nod, weak=False):
    from parafem_b200 import host
    nye = n * ctx.nranks if weak else n
    if program == "p123":
        return host.cube_p123(n, nye, n, aa=1.0 / n, bb=1.0 / n, cc=1.0 / n, limit=20000, npes=ctx.nranks, numpe=ctx.rank + 1), nye
    return host.cube_p121(n, nye, n, nod, aa=10.0 / n, bb=10.0 / n, cc=10.0 / n, limit=20000, npes=ctx.nranks,
                          numpe=ctx.rank + 1), nye


def measure(ctx, prob):
    """W warm-up + K timed iterations device-resident (`value`), then the same through pf_pcg_solve with pinned
    host buffers (`e2e`)."""
    from parafem_b200._lib import lib, ptr
    s, K, W = ctx.s, ctx.K, ctx.W
    r_np, x_np = ctx.pinned(prob.neq_pp)
    r_np[:] = prob.r_pp
    it_c, cv_c = C.c_int(), C.c_int()
    s.pcg_load_rhs(prob.r_pp)
    s.pcg_run(-1.0, W)                       # W untimed warm-up steps (tol < 0: never converges)
    sampler = ClockSampler(ctx.local)
    ctx.barrier()
    launches0 = s.kernel_launches()
    if ctx.rank == 0:
        sampler.start()
    iters, _, ms = s.pcg_run(-1.0, K)        # THE timed region: exactly K steps, CUDA events on the solver stream
    ctx.barrier()
    clocks = sampler.stop() if ctx.rank == 0 else None
    launches = s.kernel_launches() - launches0
    assert iters == K
    ms = ctx.maxf(ms)
    # the same K steps again with CUDA events around every launch: kernel shares and the roofline's launch time
    s.pcg_load_rhs(prob.r_pp)
    s.pcg_run(-1.0, W)
    s.set_profile(True)
    s.reset_profile()
    ctx.barrier()
    _, _, ms_prof = s.pcg_run(-1.0, K)
    ms_prof = ctx.maxf(ms_prof)
    km = {name: s.kernel_ms(i) for i, name in enumerate(KERNEL_KINDS)}
    s.set_profile(False)
    # the mat-vec kernel alone, max(K, 20) launches back to back between ONE pair of events: what a launch costs without
    # the two event records around it (they add ~10 us -- nothing at config C, 10 % at config B's 0.1 ms launches)
    ctx.barrier()
    mv_b2b = ctx.maxf(s.measure_matvec(max(K, 20)))

    def e2e_once(k):
        ctx.barrier()
        t = time.perf_counter()
        rc = lib().pf_pcg_solve(s._h, ptr(r_np), -1.0, k, ptr(x_np), C.byref(it_c), C.byref(cv_c))
        assert rc == 0 and it_c.value == k
        chk = float(x_np[0]) if x_np.size else 0.0   # the step's result is read on the host
        return ctx.maxf(time.perf_counter() - t), chk

    e2e_once(W)
    e2e_s, _ = e2e_once(K)
    return {"ms": ms, "ms_profiled": ms_prof, "value": prob.neq * K / (ms / 1e3) / 1e6,
            "launches": launches, "clocks": clocks, "kernel_ms": km, "matvec_back_to_back_ms": mv_b2b, "e2e_s": e2e_s,
            "e2e_value": prob.neq * K / e2e_s / 1e6}


def solve_to_convergence(ctx, prob):
    """The whole solve through pf_pcg_solve with pinned host buffers: device time of the iteration loop (`solve_s`)
    and wall clock of the call with both copies inside (`e2e_s`)."""
    from parafem_b200._lib import lib, ptr
    s = ctx.s
    r_np, x_np = ctx.pinned(prob.neq_pp)
    r_np[:] = prob.r_pp
    it_c, cv_c = C.c_int(), C.c_int()
    ctx.barrier()
    t = time.perf_counter()
    rc = lib().pf_pcg_solve(s._h, ptr(r_np), prob.tol, prob.limit, ptr(x_np), C.byref(it_c), C.byref(cv_c))
    assert rc == 0
    x1 = float(x_np[0]) if x_np.size else 0.0
    wall = ctx.maxf(time.perf_counter() - t)
    ms_full = ctx.maxf(s.last_solve_ms())
    return {"iters": it_c.value, "converged": bool(cv_c.value), "solve_s": ms_full / 1e3, "e2e_s": wall,
            "mdof_iters_per_s": prob.neq * it_c.value / (ms_full / 1e3) / 1e6,
            "e2e_mdof_iters_per_s": prob.neq * it_c.value / wall / 1e6,
            "x1": x1 if ctx.rank == 0 else None, "tol": prob.tol}


def _b2b(work, m, peak, scale):
    """The same kernel, same work, timed as max(K, 20) launches back to back between one pair of events."""
    ms = m.get("matvec_back_to_back_ms")
    if not ms:
        return None
    ach = work / (ms / 1e3) / scale
    return {"avg_launch_ms": ms, "achieved": ach, "frac": ach / peak,
            "note": "pf_measure_matvec: the kernel alone, launches back to back between ONE pair of events (right-hand sides "
                    "resident, no per-launch event records)"}


def hbm_roofline(ctx, prob, m, layout=0):
    ntot = prob.ntot
    per = ntot * (ntot + 1) // 2 if layout == 1 else ntot * ntot
    bytes_pp = prob.nels_pp * per * 8
    mv_ms, mv_n = m["kernel_ms"]["matvec"]
    avg = mv_ms / max(mv_n, 1)
    achieved = bytes_pp / (avg / 1e3) / 1e9 if mv_n else None
    # DRAM bytes per launch from ncu --set full (profiles/traffic.json): captured on ONE GPU at this element count, or
    # scaled from the capture's bytes per element when this rank holds a different number of elements of the same type
    traffic, src = None, None
    elem = f"p{prob.program}_hex{prob.nod}"
    cands = [c for c in ctx.traffic.get("captures", []) if c.get("elem") == elem and c.get("layout", 0) == layout]
    if cands:
        ent = min(cands, key=lambda c: abs(c["nels_pp"] - prob.nels_pp))
        if ent["nels_pp"] == prob.nels_pp:
            traffic, src = ent["dram_bytes_per_launch"], ent.get("source", "ncu --set full, this workload")
        else:
            traffic = ent["dram_bytes_per_launch"] / ent["nels_pp"] * prob.nels_pp
            src = (f"ncu --set full of the same kernel on one GPU at {ent['nels_pp']} elements "
                   f"({ent.get('source', '')}), scaled per element to this rank's {prob.nels_pp}")
    return {"bound": "hbm", "achieved": achieved, "peak": ctx.pk["hbm_gbs"], "unit": "GB/s",
            "frac": (achieved / ctx.pk["hbm_gbs"]) if achieved else None, "traffic": traffic, "traffic_source": src,
            "kernel": ("k_matvec_sym (packed lower triangles" if layout == 1 else "k_matvec (storkm stream") + "; gather fused)",
            "peak_kind": ctx.pk_kind, "algorithmic_bytes_per_launch": bytes_pp, "avg_launch_ms": avg,
            "launches_timed": int(mv_n),
            "timed_in": "a second pass of the same K steps with CUDA events on the solver stream around every launch "
                        "(the headline pass has no per-launch events: graph replay)",
            "back_to_back": _b2b(bytes_pp, m, ctx.pk["hbm_gbs"], 1e9)}


def mf_flops(mode, nodn, kernel=None):
    # flops per element as executed (fma = 2 flop, mul/add = 1), 8 Gauss points:
    # phase 1  H = der x p: 9 fma per node and point; per point G (9 mul + 18 fma), eps (3 add), sigma, T (9 mul + 18 fma);
    # phase 3: per freedom a 24-term chain.  sigma = D eps: k_apply_mf / k_apply_mf2 multiply the full 6x6 deemat
    # (12 mul + 30 fma = 72 flop -> 165 per point, first term of the phase-3 chain a product: 47 flop per freedom);
    # the tensor-core kernels leave deemat's structural zeros out (24 flop -> 117 per point) and run the chain as
    # 24 fma from 0.0 (48 flop per freedom).  The padding of the node tiles (20 -> 24) is NOT counted.
    # Mode 1 adds the Jacobian pass: 9 fma per node and point + det/inverse/det*w (51 flop) per point.
    tensor = (kernel or mf_kernel_name()) in ("k_apply_mf3", "k_apply_mf4")
    fl = 8 * 18 * nodn + 8 * (117 if tensor else 165) + 3 * nodn * (48 if tensor else 47)
    if mode == 1:
        fl += 8 * 18 * nodn + 8 * 51
    return fl


def mf_kernel_name():
    sel = os.environ.get("PF_MF", "")
    return {"1lane": "k_apply_mf", "2lane": "k_apply_mf2", "3": "k_apply_mf3"}.get(sel, "k_apply_mf4")


def mf_hbm_side(ctx, prob, mode, avg_ms):
    """The matrix-free kernel's OTHER roofline.  It streams, per element, the geometric factors (mode 2: 8 x (jac^-1, det*w)
    = 640 B) or the coordinates (mode 1: 24 x nod B), the gather indices (4 x ntot B), writes utemp (8 x ntot B) and reads
    every right-hand side at least once (8 B per equation): with the tensor-core kernel that is no longer negligible."""
    per_el = (640 if mode == 2 else 24 * prob.nod) + 12 * prob.ntot
    alg = prob.nels_pp * per_el + prob.neq_pp * 8
    traffic, src = None, None
    cands = [c for c in ctx.traffic.get("captures", []) if c.get("elem") == f"p{prob.program}_hex{prob.nod}" and c.get("layout") == f"mf{mode}"]
    if cands:
        ent = min(cands, key=lambda c: abs(c["nels_pp"] - prob.nels_pp))
        traffic = ent["dram_bytes_per_launch"] / ent["nels_pp"] * prob.nels_pp
        src = ent.get("source", "") + ("" if ent["nels_pp"] == prob.nels_pp else f", scaled per element to this rank's {prob.nels_pp}")
    ach = alg / (avg_ms / 1e3) / 1e9
    return {"bound": "hbm", "algorithmic_bytes_per_launch": alg, "bytes_per_element": per_el, "achieved": ach, "unit": "GB/s",
            "peak": ctx.pk["hbm_gbs"], "frac": ach / ctx.pk["hbm_gbs"], "traffic": traffic, "traffic_source": src}


def mf_roofline(ctx, prob, mode, m, peaks):
    """peaks = (DFMA loop, DMMA loop) TFLOP/s measured on this device.  k_apply_mf4 / k_apply_mf3 run their two node sums on
    the FP64 tensor pipe, so their denominator is the tensor figure (the higher one); the DFMA figure stays in the line."""
    dfma, dmma = peaks
    mv_ms_, mv_n_ = m["kernel_ms"]["matvec"]
    avg = mv_ms_ / max(mv_n_, 1)
    fl = prob.nels_pp * mf_flops(mode, prob.nod)
    tensor = mf_kernel_name() in ("k_apply_mf3", "k_apply_mf4")
    peak = dmma if tensor else dfma
    ach = fl / (avg / 1e3) / 1e12
    return {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
            "frac": ach / peak, "traffic": None,
            "kernel": mf_kernel_name() + " (matrix-free operator form, gather fused)",
            "peak_kind": ("mma.sync.m8n8k4.f64 micro-benchmark on this device (pf_measure_fp64_tensor), 512 flop per warp "
                          "instruction" if tensor else "DFMA micro-benchmark on this device (pf_measure_fp64), fma = 2 flop"),
            "peak_dfma": dfma, "peak_fp64_tensor": dmma, "frac_of_dfma_peak": ach / dfma,
            "algorithmic_flops_per_launch": fl, "flops_per_element": mf_flops(mode, prob.nod),
            "flops_note": "flops of the operator form as this kernel executes it (see mf_flops); padding of the node tiles to 24 not counted",
            "avg_launch_ms": avg, "launches_timed": int(mv_n_),
            "back_to_back": _b2b(fl, m, peak, 1e12),
            "hbm_side": mf_hbm_side(ctx, prob, mode, avg)}


def block(ctx, prob, m, roofline, tts=None):
    """value / e2e / roofline / kernel shares of one measured configuration."""
    K = ctx.K
    d = {"value": m["value"], "unit": UNIT, "ms_per_step": m["ms"] / K,
         "ms_per_step_with_per_launch_events": m["ms_profiled"] / K,
         "e2e": {"value": m["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": prob.neq_pp * 8 / K,
                 "d2h_bytes_per_step": prob.neq_pp * 8 / K, "seconds": m["e2e_s"]},
         "kernel_ms_per_step": {k: m["kernel_ms"][k][0] / K for k in KERNEL_KINDS},
         "roofline": roofline, "gpu_launches": int(m["launches"]), "clocks": m["clocks"]}
    if tts is not None:
        d["time_to_solution"] = tts
    return d


def parity_check(ctx):
    """Before anything is timed: two small meshes solved on these N GPUs through the C-ABI and compared with the
    oracle EMULATING THE SAME N RANKS (ParaFEM's partition, rank-ordered partial sums, blocked reductions) --
    np.array_equal on the field, equal iteration counts and convergence histories."""
    import oracle
    from parafem_b200 import host, solver
    specs, ok_all = [], True
    for name, dims, nod in (("hex20_6x7x5", (6, 7, 5), 20), ("hex8_10x9x6", (10, 9, 6), 8)):
        p = host.cube_p121(*dims, nod, aa=1., bb=1., cc=1., limit=500, npes=ctx.nranks, numpe=ctx.rank + 1)
        solver.setup_problem(ctx.s, p)
        x, iters, conv = ctx.s.pcg_solve(p.r_pp, p.tol, p.limit)
        hist = ctx.s.ratio_history()
        full = oracle.cube_p121(*dims, nod=nod, aa=1., bb=1., cc=1., limit=500)      # the oracle's own mesh
        km = oracle.form_km_elastic(full.g_coord_pp, nod, full.nip, full.e, full.v)
        ref = oracle.pcg(km, full.g_g_pp, full.neq, full.r_pp, full.tol, full.limit, npes=ctx.nranks, red_mode=1)
        lo = p.ieq_start - 1
        ok = (conv and iters == ref["iters"] and np.array_equal(x, ref["x"][lo:lo + p.neq_pp])
              and np.array_equal(hist, ref["ratio"]))
        ok = ctx.all_true(ok)
        ok_all = ok_all and ok
        specs.append({"mesh": name, "neq": p.neq, "iters": iters, "oracle_iters": ref["iters"], "bit_equal": ok})
    return {"nranks": ctx.nranks, "specs": specs, "bit_equal": ok_all, "transport": ctx.s.halo_transport(),
            "oracle": f"oracle.pcg(npes={ctx.nranks}, red_mode=1) on the oracle's own p12meshgen restatement",
            "compared": "xnew_pp (np.array_equal), iteration count, convergence flag, checon ratio of every iteration"}


def reference_gpu_kernel(ctx):
    """The reference's own GPU kernel beside this library's, same device buffers: xx3/cuda_helpers.cu
    matrix_vector_multiplies (MultiMatVecMultiply1, cuda_helpers.cu:144-177, 277-368) compiled into oracle/_ref by
    oracle/Makefile, and the symbol of the same name exported by libparafem_b200.so (include/parafem_xx3_compat.h),
    on the largest hex20 batch the reference's 32-bit index arithmetic admits (matrix_id*n_col*n_row < 2^31)."""
    import oracle
    from parafem_b200._lib import XX3_SIGNATURES, lib
    if not oracle.ref_available():
        return {"unavailable": "oracle/_ref/libxx3_cuda_helpers.so was not built (needs /root/reference at build time)"}
    torch = ctx.torch
    n_mat, ntot = 596000, 60
    dev = torch.device("cuda", ctx.local)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev); g.manual_seed(1)
    km = torch.empty(n_mat * ntot * ntot, dtype=torch.float64, device=dev)
    km.normal_(generator=g)
    pm = torch.randn(n_mat * ntot, dtype=torch.float64, device=dev, generator=g)
    out = {}
    bytes_alg = n_mat * ntot * ntot * 8
    results = {}
    for name, L in (("reference_cuda_helpers_default_build", oracle.ref_lib(nofma=False)),
                    ("reference_cuda_helpers_fmad_false", oracle.ref_lib(nofma=True)),
                    ("parafem_b200", lib())):
        fn = L.matrix_vector_multiplies
        fn.restype, fn.argtypes = XX3_SIGNATURES["matrix_vector_multiplies"]
        ut = torch.zeros(n_mat * ntot, dtype=torch.float64, device=dev)
        a, b, c = C.c_int(n_mat), C.c_int(ntot), C.c_int(ntot)
        d_l, d_m, d_r = C.c_void_p(pm.data_ptr()), C.c_void_p(km.data_ptr()), C.c_void_p(ut.data_ptr())
        call = lambda: fn(C.byref(a), C.byref(b), C.byref(c), C.byref(d_l), C.byref(d_m), C.byref(d_r))  # noqa: E731
        for _ in range(3):
            assert call() == 0
        times = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()                      # legacy default stream: where both libraries launch (and then synchronise)
            assert call() == 0
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = statistics.median(times)
        results[name] = ut
        out[name] = {"ms_per_call": ms, "gbs": bytes_alg / (ms / 1e3) / 1e9,
                     "frac_of_hbm_peak": bytes_alg / (ms / 1e3) / 1e9 / ctx.pk["hbm_gbs"]}
    out["same_bits_as_reference_fmad_false"] = bool(torch.equal(results["parafem_b200"], results["reference_cuda_helpers_fmad_false"]))
    out["max_rel_diff_vs_reference_default_build"] = float(
        (results["parafem_b200"] - results["reference_cuda_helpers_default_build"]).abs().max() /
        results["reference_cuda_helpers_default_build"].abs().max())
    out["speedup_vs_reference_default_build"] = (out["reference_cuda_helpers_default_build"]["ms_per_call"] /
                                                 out["parafem_b200"]["ms_per_call"])
    out["workload"] = (f"{n_mat} element matrices of 60x60 FP64 ({bytes_alg / 1e9:.2f} GB), pmul_pp / utemp_pp materialised as in "
                       f"xx3.f90:496-529; CUDA events around each synchronous call, median of 10 after 3 warm-ups")
    del km, pm, results
    torch.cuda.empty_cache()
    return out


def transient_bench(ctx):
    """the drivers that reuse the three kernels.  p124: K time steps, each one resident PCG solve
    (value = neq * PCG iterations / s); p125: K explicit steps (value = neq * steps / s).  Same JSON keys as the
    headline line; not a BASELINE config, so `vs_baseline` is null and there is no cpu_baseline leg."""
    from parafem_b200 import host, solver
    args, rank, nranks, s = ctx.args, ctx.rank, ctx.nranks, ctx.s
    n, K, W = args.n, ctx.K, ctx.W
    nye = n * nranks if args.weak else n
    if args.program == "p1210":
        # explicit elasto-plastic dynamics on 20-node bricks; --matrix-free 1 selects the tensor-core operator form
        prob = host.cube_p1210(n, nye, n, dtim=2.0e-3 * 4.0 / n, npes=nranks, numpe=rank + 1, form=1 if args.matrix_free else 0)
    else:
        make = host.cube_p124 if args.program == "p124" else host.cube_p125
        prob = make(n, nye, n, npes=nranks, numpe=rank + 1)
    solver.setup_problem(s, prob)
    sampler = ClockSampler(ctx.local)

    def run(k):
        """k steps -> (device ms, PCG iterations or steps)."""
        if args.program == "p125":
            return s.explicit_steps(k), k
        if args.program == "p1210":
            return s.vm_explicit_steps(k), k
        ms_, its_ = 0.0, 0
        for _ in range(k):
            it, _, m = s.transient_step(prob.tol, prob.limit)
            ms_ += m
            its_ += it
        return ms_, its_

    if args.program == "p124":
        s.transient_start(prob.val0)
    elif args.program == "p125":
        s.explicit_start(prob.val0)
    run(W)
    ctx.barrier()
    l0 = s.kernel_launches()
    if rank == 0:
        sampler.start()
    ms, units = run(K)                       # timed region: resident, graph replay of the PCG iteration (p124)
    ctx.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = s.kernel_launches() - l0
    ms = ctx.maxf(ms)
    t = time.perf_counter()
    _, e_units = run(K)
    # e2e: the same through the C-ABI + the field read back to the host
    x = s.vm_explicit_get()[0] if args.program == "p1210" else s.pcg_get_x()
    e2e_s = ctx.maxf(time.perf_counter() - t)
    s.set_profile(True); s.reset_profile()   # a third pass with CUDA events around every launch: kernel shares
    run(K)
    km = {name: s.kernel_ms(i) for i, name in enumerate(KERNEL_KINDS)}
    s.set_profile(False)
    mv_ms, mv_n = km["matvec"]
    bytes_pp = prob.nels_pp * 64 * 8
    achieved = bytes_pp / (mv_ms / max(mv_n, 1) / 1e3) / 1e9 if mv_n else None
    if rank == 0:
        pk = ctx.pk
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
                             "kernel": "k_matvec<8> (storka / store_pm stream; gather fused)", "peak_kind": ctx.pk_kind,
                             "algorithmic_bytes_per_launch": bytes_pp, "launches_timed": int(mv_n)},
                "kernel_ms_per_step": {k: v[0] / K for k, v in km.items()},
                "clocks": clocks, "x_checksum": float(x.sum())}
        if args.program == "p1210":
            # no element matrices: the element kernel is judged by its flops (form 1: FP64 tensor cores, the matrix-free
            # operator with the Jacobian rebuilt + ~80 flop per Gauss point of plasticity) and by its bytes (coordinates
            # 480, indices 240, products 480, Gauss-point state 8 x 96 read + written per element)
            avg = mv_ms / max(mv_n, 1)
            form = 1 if args.matrix_free else 0
            per_el = 480 + 240 + 480 + 2 * 8 * 96
            alg = prob.nels_pp * per_el + prob.neq_pp * 8
            fl = prob.nels_pp * (mf_flops(1, 20, "k_apply_mf4") + 8 * 80)
            line["config"]["workload"] = (f"p1210 {n}x{nye}x{n} 20-node bricks (p12meshgen's p121 cube run as explicit elasto-plastic "
                                          f"dynamics): {prob.nels} elements, {prob.neq} equations, dtim {prob.dtim}, form {form}")
            line["config"]["step"] = "one explicit step (predictor, gather, Gauss-point stress update, scatter, velocity / acceleration update)"
            line["config"]["l2"] = f"Gauss-point state per rank {prob.nels_pp * 8 * 96 / 1e6:.0f} MB"
            dfma, dmma = s.measure_fp64(), s.measure_fp64_tensor()
            line["roofline"] = {"bound": "fp64" if form else "fp64 (one CTA per element, not tuned)", "unit": "TFLOP/s",
                                "achieved": fl / (avg / 1e3) / 1e12, "peak": dmma if form else dfma,
                                "frac": fl / (avg / 1e3) / 1e12 / (dmma if form else dfma), "traffic": None,
                                "kernel": "k_apply_mf4<MID 1> (operator form, FP64 tensor cores)" if form else "k_p1210_elements (elements_2 as written)",
                                "flops_per_element_operator_form": fl // max(prob.nels_pp, 1), "avg_launch_ms": avg, "launches_timed": int(mv_n),
                                "hbm_side": {"bytes_per_element": per_el, "algorithmic_bytes_per_launch": alg,
                                             "achieved": alg / (avg / 1e3) / 1e9, "peak": pk["hbm_gbs"],
                                             "frac": alg / (avg / 1e3) / 1e9 / pk["hbm_gbs"]}}
        print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return

    from parafem_b200 import solver
    ctx = Ctx(args)
    rank, nranks, s, K, W = ctx.rank, ctx.nranks, ctx.s, ctx.K, ctx.W
    if args.program in ("p124", "p125", "p1210"):
        transient_bench(ctx)
        ctx.close()
        return
    if args.program == "p123":
        args.nod = 8
    headline = (args.program, args.n, args.nod, args.matrix_free, args.layout, args.weak) == ("p121", 125, 20, 0, 0, False)
    extra = args.extra and headline

    parity = parity_check(ctx) if extra else None

    # ---- the main measurement (BASELINE config C by default) -------------------------------------
    t0 = time.time()
    prob, nye = make_problem(ctx, args.program, args.n, args.nod, args.weak)
    t_mesh = time.time() - t0
    t0 = time.time()
    solver.setup_problem(s, prob, matrix_free=args.matrix_free, layout=args.layout)
    ctx.barrier()
    t_dev_setup = time.time() - t0
    fp64_tflops = (s.measure_fp64(), s.measure_fp64_tensor()) if args.matrix_free else None
    hbm_read_gbs = s.measure_hbm_read() if not args.matrix_free else None   # read-only stream ceiling (informational)
    m = measure(ctx, prob)
    tts = None if args.no_solve else solve_to_convergence(ctx, prob)
    if args.matrix_free:
        roof = mf_roofline(ctx, prob, args.matrix_free, m, fp64_tflops)
    else:
        roof = hbm_roofline(ctx, prob, m, args.layout)
        roof["read_only_stream_gbs"] = hbm_read_gbs
        roof["frac_of_read_only_stream"] = (roof["achieved"] / hbm_read_gbs) if (roof["achieved"] and hbm_read_gbs) else None
    main_cfg = workload_config(args.program, args.n, nye, args.n, args.nod, prob.nels, prob.neq, prob.nip)

    # ---- variants of the same workload: packed lower triangles, matrix-free (BASELINE config E) ---------
    variants = None
    if args.variants and args.program == "p121" and not args.matrix_free and args.layout == 0:
        variants = {}
        solver.setup_problem(s, prob, layout=1)
        mm = measure(ctx, prob)
        v = block(ctx, prob, mm, hbm_roofline(ctx, prob, mm, 1),
                  None if args.no_solve else solve_to_convergence(ctx, prob))
        v["note"] = ("K(i,j) = L(max,min): bit-identical to MATMUL on the symmetrised matrices; not the headline because the "
                     "reference's storkm_pp is only symmetric to rounding")
        variants["stored_symmetric_packed"] = v
        peak = None
        for mode, name in ((2, "matrix_free_geometric_factors"), (1, "matrix_free_rebuilt_from_coordinates")):
            solver.setup_problem(s, prob, matrix_free=mode)
            peak = peak or (s.measure_fp64(), s.measure_fp64_tensor())
            mm = measure(ctx, prob)
            variants[name] = block(ctx, prob, mm, mf_roofline(ctx, prob, mode, mm, peak),
                                   solve_to_convergence(ctx, prob) if (mode == 2 and not args.no_solve) else None)

    # ---- the other named configurations, and weak scaling ---------------------------------------------
    configs, weak = None, None
    if extra:
        configs = {}
        todo = [("D_p121_hex8_200", "p121", 200, 8)]
        if nranks == 1:
            todo.insert(0, ("B_p123_100", "p123", 100, 8))
        for name, program, n, nod in todo:
            p2, nye2 = make_problem(ctx, program, n, nod)
            solver.setup_problem(s, p2)
            mm = measure(ctx, p2)
            b = block(ctx, p2, mm, hbm_roofline(ctx, p2, mm),
                      None if args.no_solve else solve_to_convergence(ctx, p2))
            b["config"] = workload_config(program, n, nye2, n, nod, p2.nels, p2.neq, p2.nip)
            b["n_gpus"] = nranks
            configs[name] = b
        pw, nyew = make_problem(ctx, "p121", 100, 20, weak=True)
        solver.setup_problem(s, pw)
        mm = measure(ctx, pw)
        weak = block(ctx, pw, mm, hbm_roofline(ctx, pw, mm))
        weak["config"] = workload_config("p121", 100, nyew, 100, 20, pw.nels, pw.neq, pw.nip)
        weak["n_gpus"], weak["scaling"] = nranks, "weak"
        weak["per_gpu"] = "100 x 100 x 100 hex20 elements (28.8 GB of storkm) per GPU: the cube grows along y with N"

    ref_kernel = reference_gpu_kernel(ctx) if (extra and rank == 0 and nranks == 1) else None

    cpu = None
    if rank == 0 and nranks == 1 and not args.no_cpu:
        avail = mem_available_bytes()
        cn = args.cpu_n or next((c for c in (64, 40, 20) if not avail or cpu_bytes_needed(args.program, c, args.nod) < 0.5 * avail), 20)
        cpu = cpu_leg(args.program, cn, args.nod, args.cpu_steps, 3)

    if rank == 0:
        line = {
            "metric": METRIC if args.program == "p121" else "p123 EBE-PCG MDOF-iters/s", "value": m["value"], "unit": UNIT,
            "n_gpus": nranks, "steps": K, "warmup": W,
            "ms_per_step": m["ms"] / K, "higher_is_better": True, "scaling": "weak" if args.weak else "strong",
            "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": main_cfg,
            "partition": f"ParaFEM calc_nels_pp / calc_neq_pp over {nranks} rank(s), halo transport: {s.halo_transport()}",
            "e2e": {"value": m["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": prob.neq_pp * 8 / K,
                    "d2h_bytes_per_step": prob.neq_pp * 8 / K, "seconds": m["e2e_s"],
                    "note": f"ONE pf_pcg_solve call of K = {K} iterations with pinned host r_pp in / xnew_pp out: the copies "
                            f"(neq_pp*8 bytes each way) are paid once per call, so the per-step bytes are those divided by K. "
                            f"A real solve (time_to_solution: 1561 iterations at config C) amortises them further "
                            f"(time_to_solution.e2e_mdof_iters_per_s); a 1-iteration call would pay them in full."},
            "gpu_launches": int(m["launches"]),
            "roofline": roof,
            "storkm_layout": "packed lower triangles" if args.layout == 1 else "storkm_pp(ntot,ntot,nels_pp) as the reference",
            "variant": {0: "stored storkm", 1: "matrix-free, rebuilt from coordinates (config E)",
                        2: "matrix-free, stored geometric factors"}[args.matrix_free],
            "ms_per_step_with_per_launch_events": m["ms_profiled"] / K,
            "kernel_ms_per_step": {("vector_and_reductions" if k == "vector" else k): m["kernel_ms"][k][0] / K for k in KERNEL_KINDS},
            "clocks": m["clocks"],
            "time_to_solution": tts,
            "parity_check": parity,
            "configs": configs,
            "weak": weak,
            "variants": variants,
            "reference_gpu_kernel": ref_kernel,
            "setup_s": {"mesh_host": t_mesh, "device_setup_incl_storkm": t_dev_setup},
            "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
        }
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
