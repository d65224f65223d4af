"""p121 / p123 program flow (p121.f90, p123.f90) on top of host.py + solver.py.

`run(prob, solver)` performs, for one rank, what the Fortran driver does between make_ggl and
the output section, and returns the quantities the reference prints to <job>.res; `write_res`
emits those lines in the reference's formats (p121.f90:72-76,83,106-109,115,121,139).
"""
import time

import numpy as np

from . import solver as _solver


def _fe(x, w=12, d=4):
    """Fortran Ew.d: 0.dddd E+xx"""
    if x == 0.0:
        s = "0." + "0" * d + "E+00"
    else:
        e = int(np.floor(np.log10(abs(x)))) + 1
        m = x / 10.0 ** e
        if abs(round(m, d)) >= 1.0:
            m /= 10.0
            e += 1
        s = f"{m:.{d}f}E{e:+03d}"
        if s.startswith("-0."):
            s = "-0." + s[3:]
    return s.rjust(w)


def run(prob, s, profile=False):
    """-> dict(iters, converged, x, solve_s, setup_s, sigma, total_load); profile=True brackets every kernel of the
    solve with CUDA events and adds kernels = {name: (total_ms, launches)}."""
    t0 = time.time()
    _solver.setup_problem(s, prob)
    t_setup = time.time() - t0
    if profile:
        s.set_profile(True)
        s.reset_profile()
    t1 = time.time()
    x, iters, conv = s.pcg_solve(prob.r_pp, prob.tol, prob.limit)
    t_solve = time.time() - t1
    out = dict(iters=iters, converged=conv, x=x, solve_s=t_solve, setup_s=t_setup, total_load=prob.total_load)
    if profile:
        names = ("mat-vec kernel (storkm stream)", "scatter kernel", "vector kernels + reductions", "halo exchange")
        out["kernels"] = {n: s.kernel_ms(i) for i, n in enumerate(names)}
        s.set_profile(False)
    if prob.program == 121:
        out["sigma"] = s.centroid_stress(0, prob.e, prob.v) if prob.numpe == 1 else None
    return out


def run_p124(prob, s, out_base=None, source=None, decimals=5):
    """Program p124 (p124.f90) for one rank: setup, then the time-stepping loop with one device-resident
    PCG solve per step.  `source` = value of the loaded freedom nres (loads_pp = source*dtim every step,
    p124.f90:143-148; None = no loads, as every shipped deck).  With `out_base` the nodal temperature
    files <out_base>.ensi.NDTTR-NNNNNN are written at step 0 and every npri steps (p124.f90:180-191,
    219-231; single rank).  -> dict(times, temps (x at freedom nres), iters, rows (the printed steps),
    x (last field), solve_s, setup_s)."""
    from . import host
    t0 = time.time()
    _solver.setup_problem(s, prob)
    s.transient_start(prob.val0, prob.val_f if prob.no_f.size else None)
    t_setup = time.time() - t0
    lo = prob.ieq_start
    owns = lo <= prob.nres < lo + prob.neq_pp
    loads = None
    if source is not None:
        loads = np.zeros(prob.neq_pp)
        if owns:
            loads[prob.nres - lo] = source * prob.dtim
    rows = [(0.0, prob.val0, None)] if owns else []
    if out_base and prob.npes == 1:
        x0 = np.full(prob.neq_pp, prob.val0)
        if prob.no_f.size:
            x0[prob.no_f - lo] = prob.val_f
        host.write_ensi(f"{out_base}.ensi.NDTTR-{0:06d}", host.nodal_values(prob, x0), decimals=decimals)
    iters, solve_ms, x = [], 0.0, None
    for j in range(1, prob.nstep + 1):
        it, _, ms = s.transient_step(prob.tol, prob.limit, loads)
        iters.append(it)
        solve_ms += ms
        if j // prob.npri * prob.npri == j:
            x = s.pcg_get_x()
            if out_base and prob.npes == 1:
                host.write_ensi(f"{out_base}.ensi.NDTTR-{j:06d}", host.nodal_values(prob, x), decimals=decimals)
            if owns:
                rows.append((j * prob.dtim, float(x[prob.nres - lo]), it))
    if x is None or prob.nstep % prob.npri:
        x = s.pcg_get_x()
    return dict(iters=iters, rows=rows, x=x, solve_s=solve_ms / 1e3, setup_s=t_setup,
                times=[r[0] for r in rows], temps=[r[1] for r in rows])


def write_res_p124(path, prob, res, t_total=0.0, t_output=0.0):
    """<job>.res of p124 as the rank owning freedom nres writes it (p124.f90:56-62,136,169,228,233-238)."""
    with open(path, "w") as f:
        f.write(f"This job ran on {prob.npes:5d}  processes\n")
        f.write(f"There are {prob.nn:12d} nodes{prob.nr:12d} restrained and   {prob.neq:12d} equations\n")
        f.write(f"Time after setup is {res['setup_s']:10.4f}\n")
        f.write("  Time       Temperature  Iterations \n")
        for t, v, it in res["rows"]:
            f.write(_fe(t) + _fe(v) + ("" if it is None else f"{it:10d}") + "\n")
        f.write(f"The solution phase took {res['solve_s']:10.4f}\n")
        f.write(f"Writing the output took {t_output:10.4f}\n")
        f.write(f"This analysis took      {t_total:10.4f}\n")


def run_p125(prob, s, out_base=None, decimals=5):
    """Program p125 (p125.f90) for one rank: element integration, then nstep passes of the explicit recursion
    on the device, npri at a time; with `out_base` the nodal files <out_base>.ensi.NDPRE-NNNNNN every npri steps
    (single rank).  -> dict(rows [(time, value at freedom nres)], x, step_s, setup_s)."""
    from . import host
    t0 = time.time()
    _solver.setup_problem(s, prob)
    s.explicit_start(prob.val0)
    t_setup = time.time() - t0
    lo = prob.ieq_start
    owns = lo <= prob.nres < lo + prob.neq_pp
    rows = [(0.0, prob.val0)] if owns else []
    ms, done, x = 0.0, 0, None
    while done < prob.nstep:
        n = min(prob.npri, prob.nstep - done)
        ms += s.explicit_steps(n)
        done += n
        if done // prob.npri * prob.npri == done:
            x = s.pcg_get_x()
            if owns:
                rows.append((done * prob.dtim, float(x[prob.nres - lo])))
            if out_base and prob.npes == 1:
                host.write_ensi(f"{out_base}.ensi.NDPRE-{done:06d}", host.nodal_values(prob, x), decimals=decimals)
    if x is None or prob.nstep % prob.npri:
        x = s.pcg_get_x()
    return dict(rows=rows, x=x, step_s=ms / 1e3, setup_s=t_setup)


def write_res_p125(path, prob, res, t_read=0.0, t_total=0.0):
    """<job>.res of p125 (p125.f90:49-56,80-81,86-87,101,116-118)."""
    with open(path, "w") as f:
        f.write(f"This job ran on {prob.npes:5d}  processes\n")
        f.write(f"There are {prob.nn:12d} nodes{prob.nr:12d} restrained and{prob.neq:12d} equations\n")
        f.write(f"Time to read input is:{t_read:10.4f}\n")
        f.write(f"Time after setup is:{t_read + res['setup_s']:10.4f}\n")
        f.write(f"Time for element integration is :{res['setup_s']:10.4f}\n")
        f.write("  Time        Pressure\n")
        for t, v in res["rows"]:
            f.write(_fe(t) + _fe(v) + "\n")
        f.write(f"Time stepping recursion took  :{res['step_s']:10.4f}\n")
        f.write(f"This analysis took  :{t_total:10.4f}\n")


XX3_SECTIONS = ("Setup", "Read element steering array", "Convert Abaqus to S&G node ordering", "Read nodal coordinates",
                "Read restrained nodes", "Compute steering array and neq", "Compute interprocessor communication tables",
                "Allocate neq_pp arrays", "Compute element stiffness matrices", "Build the preconditioner",
                "Get starting r", "Solve equations", "Output results")


def run_p129(prob, s, out_base=None, decimals=4, nstep=None):
    """Program p129 (p129.f90) for one rank: setup, then the time loop -- per step the harmonic load factor on the host
    (p129.f90:124-125) and one device call.  With `out_base` the displacement files <out_base>.ensi.DISPL-NNNNNN are
    written every npri steps (single rank).  -> dict(rows = [(time, cos(omega t), x1(nres), iters)], x, d1x, d2x, dtim,
    solve_s, setup_s)."""
    import math
    from . import host
    t0 = time.time()
    _solver.setup_problem(s, prob)
    t_setup = time.time() - t0
    dtim, theta, omega = prob.dtim, prob.theta, prob.omega
    c1 = (1.0 - theta) * dtim
    lo = prob.ieq_start
    owns = lo <= prob.nres < lo + prob.neq_pp
    rows, ms_total, real_time = [], 0.0, 0.0
    for j in range(1, (nstep or prob.nstep) + 1):
        real_time = real_time + dtim
        factor = theta * dtim * math.cos(omega * real_time) + c1 * math.cos(omega * (real_time - dtim))
        it, _, ms = s.dynamic_step(factor, prob.tol, prob.limit)
        ms_total += ms
        if j // prob.npri * prob.npri == j:
            x1, _, _ = s.dynamic_get()
            if owns:
                rows.append((real_time, math.cos(omega * real_time), float(x1[prob.nres - lo]), it))
            if out_base and prob.npes == 1:
                host.write_ensi(f"{out_base}.ensi.DISPL-{j:06d}", host.nodal_values(prob, x1), decimals=decimals)
    x, d1x, d2x = s.dynamic_get()
    return dict(rows=rows, x=x, d1x=d1x, d2x=d2x, dtim=dtim, solve_s=ms_total / 1e3, setup_s=t_setup)


def write_res_p129(path, prob, res, t_total=0.0):
    """<job>.res as p129.f90:50-56,113,150-152,176-178 writes it."""
    with open(path, "w") as f:
        f.write(f"This job ran on {prob.npes:6d} processes\n")
        f.write(f"There are {prob.nn:12d} nodes {prob.nr:12d} restrained and {prob.neq:12d} equations\n")
        f.write(f"Time after setup was:{res['setup_s']:10.4f}\n")
        f.write("   Time t  cos(omega*t) Displacement Iterations\n")
        for t, c, x, it in res["rows"]:
            f.write(f"{_fe(t)}{_fe(c)}{_fe(x)}{it:10d}\n")
        f.write(f"This analysis took:{t_total:10.4f}\n")


def run_p1210(prob, s, out_base=None, decimals=4, nstep=None):
    """Program p1210 (p1210.f90) for one rank: setup (lumped mass, loads), then the explicit loop -- npri steps per
    device call.  With `out_base` the displacement files <out_base>.ensi.DISPL-NNNNNN are written every npri steps
    (single rank).  -> dict(rows = [(time, x1(nres), d1x1(nres), d2x1(nres))] starting with the t = 0 row, x, d1x, d2x,
    fields = {step: x1_pp}, solve_s, setup_s)."""
    from . import host
    t0 = time.time()
    _solver.setup_problem(s, prob)
    t_setup = time.time() - t0
    lo = prob.ieq_start
    owns = lo <= prob.nres < lo + prob.neq_pp
    k = prob.nres - lo
    rows, fields, ms_total = [], {}, 0.0
    if owns:
        rows.append((0.0, 0.0, 0.0, 0.0))
    nstep = prob.nstep if nstep is None else nstep
    real_time, done = 0.0, 0
    while done + prob.npri <= nstep:
        ms_total += s.vm_explicit_steps(prob.npri)
        done += prob.npri
        for _ in range(prob.npri):
            real_time = real_time + prob.dtim                 # real_time=real_time+dtim, step by step (p1210.f90:115)
        x1, d1, d2 = s.vm_explicit_get()
        fields[done] = x1
        if owns:
            rows.append((real_time, float(x1[k]), float(d1[k]), float(d2[k])))
        if out_base and prob.npes == 1:
            host.write_ensi(f"{out_base}.ensi.DISPL-{done:06d}", host.nodal_values(prob, x1), decimals=decimals)
    if done < nstep:
        ms_total += s.vm_explicit_steps(nstep - done)
    x, d1x, d2x = s.vm_explicit_get()
    return dict(rows=rows, x=x, d1x=d1x, d2x=d2x, fields=fields, solve_s=ms_total / 1e3, setup_s=t_setup)


def write_res_p1210(path, prob, res, t_total=0.0):
    """<job>.res as p1210.f90:55-60,113-114,152-153 writes it."""
    with open(path, "w") as f:
        f.write(f"This job ran on {prob.npes:6d} processes\n")
        f.write(f"There are {prob.nn:12d} nodes {prob.nr:12d} restrained and {prob.neq:12d} equations\n")
        f.write(f"Time after setup was:{res['setup_s']:10.4f}\n")
        f.write("  Time      Displacement  Velocity   Acceleration \n")
        for t, x, v, a in res["rows"]:
            f.write(f"{_fe(t)}{_fe(x)}{_fe(v)}{_fe(a)}\n")
        f.write(f"This analysis took:{t_total:10.4f}\n")


def run_p122(prob, s, out_base=None, decimals=4):
    """Program p122 (p122.f90) for one rank: setup, then the load-increment loop -- every increment one device call
    (plastic iterations, each a PCG solve restarted from the current x plus the Gauss-point stress update).  With
    `out_base` the displacement files <out_base>.ensi.DISPL-NNNNNN are written after every increment (p122.f90:215-228;
    single rank).  -> dict(rows = [(totd(1), sigma z, sigma x, sigma y, cjtot, plasiters)], totd, dt, solve_s, setup_s)."""
    from . import host
    t0 = time.time()
    _solver.setup_problem(s, prob)
    t_setup = time.time() - t0
    rows, ms_total = [], 0.0
    ld0 = prob.r_pp if getattr(prob, "loaded_nodes", 0) else None
    for iy, q in enumerate(prob.qinc, 1):
        plasiters, cjtot, ms = s.plastic_increment(q, prob.plasits, prob.plastol, prob.cjits, prob.cjtol, ld0_pp=ld0,
                                                   valf_pp=prob.val_f if prob.no_f.size else None)
        ms_total += ms
        totd = s.plastic_totd()
        t = s.plastic_tensor(0, 0) if prob.numpe == 1 else np.zeros(6)
        rows.append((float(totd[0]) if prob.numpe == 1 else None, t[2], t[0], t[1], cjtot, plasiters))
        if out_base and prob.npes == 1:
            host.write_ensi(f"{out_base}.ensi.DISPL-{iy:06d}", host.nodal_values(prob, totd), decimals=decimals)
        if plasiters == prob.plasits:
            break
    return dict(rows=rows, totd=totd, dt=prob.dt, solve_s=ms_total / 1e3, setup_s=t_setup)


def write_res_p122(path, prob, res, t_total=0.0):
    """<job>.res as p122.f90:59-64,93,117,206-214,230 writes it."""
    with open(path, "w") as f:
        f.write(f"This job ran on {prob.npes:5d}  processes\n")
        f.write(f"There are {prob.nn:7d} nodes{prob.nr:7d} restrained and   {prob.neq:7d} equations\n")
        f.write(f"Time after setup is:{res['setup_s']:10.4f}\n")
        f.write(f"The critical timestep is   {_fe(res['dt'])}\n")
        for iy, (d1, sz, sx, sy, cjtot, plasiters) in enumerate(res["rows"], 1):
            f.write(f"\nLoad Increment   {iy:5d}\n")
            f.write(f"The displacement is  {_fe(d1)}\n")
            f.write("  sigma z    sigma x     sigma y\n")
            f.write(f"{_fe(sz)}{_fe(sx)}{_fe(sy)}\n")
            f.write(f"The total number of cj iterations was  {cjtot:12d}\n")
            f.write(f"The number of plastic iterations was  {plasiters:12d}\n")
            f.write(f"cj iterations per plastic iteration were {cjtot / plasiters:11.2f}\n")
        f.write(f"This analysis took: {t_total:10.4f}\n")


def write_res_xx3(path, prob, res, loaded_nodes, seconds, kernels=None):
    """<job>.res in the layout of the reference's GPU driver xx3 (programs/dev/xx3/xx3.f90, golden
    examples/dev/xx3/demo/xx3-tiny.res): BASIC JOB DATA, then the 'section / seconds / %total' table -- SURVEY 5.1's
    hook for the device timings.  `seconds`: {section name: seconds} for the names in XX3_SECTIONS (missing = 0);
    `kernels`: optional {name: (total_ms, launches)} from Solver.kernel_ms, appended as indented rows under
    'Solve equations' (CUDA events on the solver stream)."""
    total = sum(seconds.get(k, 0.0) for k in XX3_SECTIONS)
    with open(path, "w") as f:
        f.write("\n" + "BASIC JOB DATA".ljust(48) + "\n")
        for label, val in (("Number of processors used", prob.npes), ("Number of nodes in the mesh", prob.nn),
                           ("Number of nodes that were restrained", prob.nr), ("Number of equations solved", prob.neq),
                           ("Number of PCG iterations", res["iters"]), ("Number of loaded nodes", loaded_nodes)):
            f.write(f"{label:<44}{val:12d}\n")
        f.write(f"{'Total load applied':<44}{_fe(res['total_load'])}\n\n")
        f.write("PROGRAM SECTION EXECUTION TIMES                  SECONDS  %TOTAL    \n")
        for name in XX3_SECTIONS:
            t = seconds.get(name, 0.0)
            f.write(f"{name:<44}{t:12.6f}{100.0 * t / total if total > 0 else 0.0:8.2f}\n")
            if name == "Solve equations" and kernels:
                for kname, (ms, n) in kernels.items():
                    f.write(f"{'  ' + kname + f' ({n} launches)':<44}{ms / 1e3:12.6f}{100.0 * ms / 1e3 / total if total > 0 else 0.0:8.2f}\n")
        f.write(f"{'Total execution time':<44}{total:12.6f}{100.0 if total > 0 else 0.0:8.2f}\n")


def write_res(path, prob, res, t_read=0.0, t_total=0.0):
    """<job>.res as rank 1 writes it."""
    with open(path, "w") as f:
        if prob.program == 121:
            f.write(f"This job ran on {prob.npes:7d} processes\n")
            f.write(f"There are {prob.nn:12d} nodes{prob.nr:12d} restrained and {prob.neq:12d} equations\n")
            f.write(f"Time to read input is:{t_read:10.4f}\n")
            f.write(f"Time after setup is:{t_read + res['setup_s']:10.4f}\n")
            f.write(f"The total load is:{_fe(res['total_load'])}\n")
            f.write(f"The number of iterations to convergence was {res['iters']:6d}\n")
            f.write(f"Time to solve equations was  :{res['solve_s']:10.4f}\n")
            f.write(f"The central nodal displacement is :{_fe(res['x'][0])}\n")
            if res.get("sigma") is not None:
                f.write("The Centroid point stresses for element 1 are\n")
                f.write(f"Point {1:5d}\n")
                f.write("".join(_fe(v) for v in res["sigma"]) + "\n")
            f.write(f"This analysis took  :{t_total:10.4f}\n")
        else:
            f.write(f"This job ran on {prob.npes:5d}  processes\n")
            f.write(f"There are {prob.nn:12d} nodes{prob.nr:12d} restrained and   {prob.neq:12d} equations\n")
            f.write(f"Time after setup is {res['setup_s']:10.4f}\n")
            f.write(f"The number of iterations to convergence was {res['iters']:5d}\n")
            f.write(f"The total load is {_fe(res['total_load'])}\n")
            f.write("The potentials are:\n Freedom       Potential\n")
            lo = prob.ieq_start
            for i in range(4):
                q = prob.nres + i
                if lo <= q < lo + prob.neq_pp:
                    f.write(f"{q:8d}     {_fe(res['x'][q - lo])}\n")
            f.write(f"Time spent in the solver was {res['solve_s']:10.4f}\n")


def main(argv=None):
    """python -m parafem_b200.driver (--deck JOB | --cube N [--hex 8|20] | --p123 N) [--matrix-free M] [--out DIR]

    Runs program p121 (or p123) on cuda:0 and writes <job>.res and the EnSight nodal file as the
    reference does (p121.f90:70-77,105-110,124-140 / p123.f90:153-181)."""
    import argparse
    import os

    from . import host
    ap = argparse.ArgumentParser(prog="parafem_b200.driver")
    g = ap.add_mutually_exclusive_group(required=True)
    g.add_argument("--deck", help="ParaFEM p121 deck base name (<job>.dat/.d/.bnd/.lds)")
    g.add_argument("--cube", type=int, help="p12meshgen p121 cube with N^3 elements, generated in memory")
    g.add_argument("--p123", type=int, help="p12meshgen p123 box with N^3 8-node bricks")
    g.add_argument("--p124", type=int, help="p12meshgen p124 box with N^3 8-node bricks (transient conduction, "
                                            "150 steps of 0.01 as the shipped p124_*.mg)")
    g.add_argument("--p125", type=int, help="p12meshgen p125 box with N^3 8-node bricks (explicit transient conduction, "
                                            "5000 steps of 0.0002 as the shipped p125_*.mg)")
    ap.add_argument("--hex", type=int, default=20, choices=[8, 20])
    ap.add_argument("--matrix-free", type=int, default=0, choices=[0, 1, 2])
    ap.add_argument("--out", default=".")
    ap.add_argument("--xx3-res", action="store_true",
                    help="p121: also write <job>.xx3.res, the section / seconds / %%total table of the reference's GPU "
                         "driver xx3 with the device kernels (CUDA events) listed under 'Solve equations'")
    a = ap.parse_args(argv)
    t0 = time.time()
    if a.deck:
        prob, job = host.read_deck_p121(a.deck), os.path.basename(a.deck)
    elif a.p124:
        prob, job = host.cube_p124(a.p124, a.p124, a.p124), f"p124_box{a.p124}"
        base = os.path.join(a.out, job)
        with _solver.Solver(0, 1, 0) as s:
            res = run_p124(prob, s, out_base=base)
        write_res_p124(base + ".res", prob, res, t_total=time.time() - t0)
        print(open(base + ".res").read(), end="")
        return 0
    elif a.p125:
        prob, job = host.cube_p125(a.p125, a.p125, a.p125), f"p125_box{a.p125}"
        base = os.path.join(a.out, job)
        with _solver.Solver(0, 1, 0) as s:
            res = run_p125(prob, s, out_base=base)
        write_res_p125(base + ".res", prob, res, t_total=time.time() - t0)
        print(open(base + ".res").read(), end="")
        return 0
    elif a.cube:
        prob, job = host.cube_p121(a.cube, a.cube, a.cube, a.hex, limit=20000), f"p121_cube{a.cube}_hex{a.hex}"
    else:
        prob, job = host.cube_p123(a.p123, a.p123, a.p123, limit=20000), f"p123_box{a.p123}"
    t_read = time.time() - t0
    with _solver.Solver(0, 1, 0) as s:
        if a.matrix_free and prob.program == 121:
            _solver.setup_problem(s, prob, matrix_free=a.matrix_free)
            x, iters, conv = s.pcg_solve(prob.r_pp, prob.tol, prob.limit)
            res = dict(iters=iters, converged=conv, x=x, solve_s=0.0, setup_s=0.0, total_load=prob.total_load,
                       sigma=s.centroid_stress(0, prob.e, prob.v))
        else:
            res = run(prob, s, profile=a.xx3_res)
    base = os.path.join(a.out, job)
    if a.xx3_res and prob.program == 121 and "kernels" in res:
        secs = {"Setup": t_read, "Compute element stiffness matrices": res["setup_s"], "Solve equations": res["solve_s"]}
        write_res_xx3(base + ".xx3.res", prob, res, 0, secs, kernels=res["kernels"])
    write_res(base + ".res", prob, res, t_read=t_read, t_total=time.time() - t0)
    kind = "DISPL" if prob.program == 121 else "NDPTL"
    host.write_ensi(f"{base}.ensi.{kind}-000001", host.nodal_values(prob, res["x"]), decimals=5)
    print(open(base + ".res").read(), end="")
    return 0 if res["converged"] else 3


if __name__ == "__main__":
    raise SystemExit(main())
