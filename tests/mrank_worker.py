"""Multi-rank parity worker, launched by torchrun (one process per GPU) from
tests/test_gpu_multirank.py:  N-GPU solve through the C-ABI == the oracle emulating the same
N ranks (same partition, same rank-ordered partial sums, same blocked reductions)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402
from parafem_b200 import host, solver  # noqa: E402
from shuffle_util import shuffled  # noqa: E402


def uneven(nels, npes):
    """A deliberately lopsided external partition (what a .psize file would hold)."""
    w = np.array([3 + (5 * r) % 7 for r in range(npes)], float)
    c = np.maximum(1, np.floor(nels * w / w.sum()).astype(int))
    c[-1] += nels - c.sum()
    return [int(v) for v in c]


def problem(name, npes, numpe):
    if name.endswith("_shuffled"):
        return shuffled(problem(name[:-len("_shuffled")], 1, 1), npes, numpe)
    if name == "hex20_psize":   # partitioner 2 (read_nels_pp, input.f90:3108-3196)
        ps = uneven(6 * 7 * 5, npes) if npes > 1 else None
        return host.cube_p121(6, 7, 5, 20, aa=1., bb=1., cc=1., limit=500, npes=npes, numpe=numpe, psize=ps)
    if name == "hex20":
        return host.cube_p121(6, 7, 5, 20, aa=1., bb=1., cc=1., limit=500, npes=npes, numpe=numpe)
    if name == "hex20_thin":   # element and equation cuts badly misaligned -> +-2 neighbours
        return host.cube_p121(5, 9, 2, 20, aa=1., bb=.5, cc=1., limit=500, npes=npes, numpe=numpe)
    if name == "hex8":
        return host.cube_p121(10, 9, 6, 8, aa=1., bb=1., cc=1., limit=500, npes=npes, numpe=numpe)
    if name == "p123":
        return host.cube_p123(9, 11, 8, limit=500, npes=npes, numpe=numpe)
    if name == "p123_fixed":
        return host.cube_p123(9, 11, 8, limit=500, npes=npes, numpe=numpe, fixed=True)
    if name in ("p124", "p124_fixed"):   # transient conduction: one PCG solve per time step
        return host.cube_p124(9, 11, 8, aa=.1, bb=.1, cc=.1, nstep=6, npes=npes, numpe=numpe, fixed=name.endswith("fixed"))
    if name == "p125":                   # explicit transient conduction
        return host.cube_p125(9, 11, 8, aa=.1, bb=.1, cc=.1, dtim=1e-4, nstep=30, npes=npes, numpe=numpe)
    if name == "p129":                   # forced vibration: three matrix sets, one PCG solve per step
        return host.cube_p129(3, 6, 2, .25, .25, .25, e=1.0e4, nip=8, nstep=4, npes=npes, numpe=numpe)
    if name == "p1210":                  # explicit elasto-plastic dynamics: no PCG, halo exchanges around the element kernel
        from p1210_util import synthetic
        return synthetic(host, 4, 9, 3, npes=npes, numpe=numpe)
    if name == "p122":                   # elasto-plasticity, loaded-nodes branch
        p = host.cube_p121(4, 5, 3, 8, aa=1., bb=1., cc=1., e=100.0, v=0.3, npes=npes, numpe=numpe)
        p.program, p.phi, p.c, p.psi = 122, 20.0, 4.0, 0.0
        p.qinc, p.plasits, p.cjits, p.plastol, p.cjtol, p.loaded_nodes = [0.5, 0.3], 12, 80, 1e-4, 1e-6, 1
        return p
    if name == "tet4":                   # 4-node tetrahedra, elastic (12x12 element matrices)
        from tet_util import tet_problem
        return tet_problem(host.cube_p121(5, 6, 3, 8, aa=1., bb=1., cc=1., limit=3000), npes, numpe)
    if name == "tet4_scalar":            # 4-node tetrahedra, p123 (4x4)
        from tet_util import tet_problem
        return tet_problem(host.cube_p123(6, 7, 5, limit=800), npes, numpe)
    if name == "hex20_mat":              # per-element materials (xx2)
        p = host.cube_p121(6, 7, 5, 20, aa=1., bb=1., cc=1., limit=800, npes=npes, numpe=numpe)
        rng = np.random.RandomState(3)
        p.prop = np.column_stack([rng.uniform(50., 5000., 5), rng.uniform(0.05, 0.45, 5)])
        p.etype_pp = rng.randint(1, 6, p.nels).astype(np.int32)[p.iel_start - 1:p.iel_start - 1 + p.nels_pp]
        return p
    raise KeyError(name)


def transient_specs(s, name, p, full, world):
    """p124 / p125 on N ranks == the oracle emulating the same N ranks, step by step."""
    lo = p.ieq_start - 1
    solver.setup_problem(s, p)
    ok = True
    if name == "p125":
        store, mass = oracle.form_k_explicit(full.g_coord_pp, full.nip, full.kx, full.ky, full.kz, full.dtim)
        ref = oracle.p125(store, mass, full.g_g_pp, full.neq, full.val0, full.nstep, npes=world, keep=(1, 7, full.nstep))
        s.explicit_start(p.val0)
        done = 0
        for j in (1, 7, full.nstep):
            s.explicit_steps(j - done)
            done = j
            ok = ok and np.array_equal(s.pcg_get_x(), ref["fields"][j][lo:lo + p.neq_pp])
        return ok, f"steps={full.nstep}"
    a, b = oracle.form_k_transient(full.g_coord_pp, full.nip, full.kx, full.ky, full.kz, full.rho, full.cp, full.theta, full.dtim)
    fixed = full.no_f.size > 0
    ref = oracle.p124(a, b, full.g_g_pp, full.neq, full.val0, full.nstep, full.tol, full.limit, npes=world, red_mode=1,
                      keep=tuple(range(1, full.nstep + 1)), no_f=full.no_f if fixed else None,
                      val_f=full.val_f if fixed else None)
    s.transient_start(p.val0, p.val_f if p.no_f.size else None)
    its = []
    for j in range(1, full.nstep + 1):
        it, conv, _ = s.transient_step(p.tol, p.limit)
        its.append(it)
        ok = ok and it == ref["iters"][j - 1] and np.array_equal(s.pcg_get_x(), ref["fields"][j][lo:lo + p.neq_pp])
    return ok, f"iters={its}/{ref['iters']}"


def driver_specs(s, name, p, full, world):
    """p129 / p122 / p1210 on N ranks == the oracle emulating the same N ranks."""
    from parafem_b200 import driver
    lo = p.ieq_start - 1
    if name == "p1210":
        form = getattr(p, "form", 0)             # "p1210:mf" = the operator form on the tensor cores
        ref = oracle.p1210(full.g_coord_pp, full.g_g_pp, full.neq, full.r_pp, full.e, full.v, full.sbary, full.rho, full.dtim,
                           full.pload, full.nstep, full.npri, npes=world, form=form)
        out = driver.run_p1210(p, s)
        ok = all(np.array_equal(out["fields"][step], x1[lo:lo + p.neq_pp]) for step, x1, _, _ in ref["snaps"])
        ok = ok and np.array_equal(out["d2x"], ref["snaps"][-1][3][lo:lo + p.neq_pp])
        return ok, f"steps={full.nstep}"
    if name == "p129":
        km = oracle.form_km_elastic(full.g_coord_pp, 20, full.nip, full.e, full.v)
        mm = oracle.form_mass(full.g_coord_pp, 20, full.nip, full.rho)
        ref = oracle.p129(km, mm, full.g_g_pp, full.neq, full.r_pp, full.theta, full.omega, full.alpha1, full.beta1, full.nstep,
                          full.tol, full.limit, npes=world, red_mode=1)
        out = driver.run_p129(p, s)
        ok = (np.array_equal(out["x"], ref["x"][lo:lo + p.neq_pp]) and np.array_equal(out["d1x"], ref["d1x"][lo:lo + p.neq_pp])
              and np.array_equal(out["d2x"], ref["d2x"][lo:lo + p.neq_pp]))
        if out["rows"]:                                    # the rank that owns nres
            ok = ok and [r[3] for r in out["rows"]] == [r[2] for r in ref["rows"]]
        return ok, f"iters={[r[2] for r in ref['rows']]}"
    from oracle import p122_oracle
    ref_rows, ref_totd = p122_oracle.p122(full.g_coord_pp, full.g_g_pp, full.neq, full.phi, full.c, full.psi, full.e, full.v,
                                          full.qinc, full.plasits, full.cjits, full.plastol, full.cjtol, ld0=full.r_pp, npes=world,
                                          red_mode=1)
    out = driver.run_p122(p, s)
    # the Lode-angle functions are CUDA's on the device and glibc's in the oracle: counts equal or within a few cj
    # iterations, fields to the cj tolerance (tests/test_gpu_plastic.py)
    ok = len(out["rows"]) == len(ref_rows)
    for r, o in zip(out["rows"], ref_rows):
        ok = ok and r[5] == o["plasiters"] and abs(r[4] - o["cjtot"]) <= max(2, 0.02 * o["cjtot"])
    mine = ref_totd[lo:lo + p.neq_pp]
    err = np.linalg.norm(out["totd"] - mine) / max(np.linalg.norm(ref_totd), 1e-300)
    return ok and err <= 1e-6, f"plas={[o['plasiters'] for o in ref_rows]} cj={[r[4] for r in out['rows']]}/{[o['cjtot'] for o in ref_rows]} err={err:.1e}"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    idt = torch.from_numpy(solver.nccl_unique_id().copy()) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(idt, src=0)
    s = solver.Solver(rank, world, local, idt.numpy())
    failures = []
    for spec in sys.argv[1:]:
        # "name" or "name:sym" (packed lower triangles) or "name:mf1" / "name:mf2" (matrix-free variants)
        name, _, variant = spec.partition(":")
        p = problem(name, world, rank + 1)
        full = problem(name, 1, 1)
        if name in ("p124", "p124_fixed", "p125", "p129", "p122", "p1210"):
            oracle.set_element_partition(None)
            if name == "p1210" and variant == "mf":
                p.form = 1
            ok, info = (driver_specs if name in ("p129", "p122", "p1210") else transient_specs)(s, name, p, full, world)
            line = f"[rank {rank}] {spec}: equal={ok} {info}"
            print(line, flush=True)
            if not ok:
                failures.append(line)
            continue
        r0 = p.r_pp.copy()
        oracle.set_element_partition(uneven(full.nels, world) if name == "hex20_psize" else None)
        mf_mode = int(variant[2:]) if variant.startswith("mf") else 0
        solver.setup_problem(s, p, matrix_free=mf_mode, layout=1 if variant == "sym" else 0)
        lo = p.ieq_start - 1
        # operator pieces
        rng = np.random.RandomState(11)
        pv = rng.randn(p.neq)
        qv = rng.randn(p.neq)
        if full.prop is not None:
            km = oracle.form_km_elastic_mat(full.g_coord_pp, full.nod, full.nip, full.prop, full.etype_pp)
        else:
            km = (oracle.form_km_elastic(full.g_coord_pp, full.nod, full.nip, full.e, full.v) if p.program == 121
                  else oracle.form_kc_laplace(full.g_coord_pp, full.nip, full.kx, full.ky, full.kz))
        if variant == "sym":      # K(i,j) = L(max,min): the oracle on the symmetrised matrices
            km = np.triu(km) + np.triu(km, 1).transpose(0, 2, 1)
        mf = dict(g_coord_pp=full.g_coord_pp, nod=full.nod, nip=full.nip, e=full.e, v=full.v, mode=mf_mode) if mf_mode else None
        pm_ref = oracle.gather(full.g_g_pp, pv)
        e0 = p.iel_start - 1
        ok_g = np.array_equal(s.gather(pv[lo:lo + p.neq_pp]), pm_ref[e0:e0 + p.nels_pp])
        ut_ref = oracle.apply_mf(full.g_coord_pp, full.nod, full.nip, full.e, full.v, pm_ref, mode=mf_mode) if mf_mode else oracle.matvec(km, pm_ref)
        u_ref = oracle.scatter(full.g_g_pp, ut_ref, p.neq, npes=world)
        s.nfixed_saved = None
        u = s.apply(pv[lo:lo + p.neq_pp])
        if p.no_f.size:   # fixed freedom rows are overridden by u = p*store in the solver path
            mask = np.ones(p.neq_pp, bool); mask[p.no_f - p.ieq_start] = False
            ok_u = np.array_equal(u[mask], u_ref[lo:lo + p.neq_pp][mask])
        else:
            ok_u = np.array_equal(u, u_ref[lo:lo + p.neq_pp])
        ok_d = s.dot(pv[lo:lo + p.neq_pp], qv[lo:lo + p.neq_pp]) == oracle.dot_ranks(pv, qv, npes=world, red_mode=1)
        # full solve
        x, iters, conv = s.pcg_solve(p.r_pp, p.tol, p.limit)
        no_f = full.no_f if full.no_f.size else None
        ref = oracle.pcg(km, full.g_g_pp, p.neq, full.r_pp if no_f is None else np.zeros(p.neq), p.tol, p.limit,
                         npes=world, red_mode=1, no_f=no_f, val_f=full.val_f if no_f is not None else None, mf=mf)
        ok_x = np.array_equal(x, ref["x"][lo:lo + p.neq_pp])
        ok_i = iters == ref["iters"] and conv == ref["converged"]
        rel = np.linalg.norm(x - ref["x"][lo:lo + p.neq_pp]) / max(np.linalg.norm(ref["x"][lo:lo + p.neq_pp]), 1e-300)
        line = f"[rank {rank}] {spec}: gather={ok_g} apply={ok_u} dot={ok_d} iters={iters}/{ref['iters']} x_equal={ok_x} rel={rel:.2e}"
        print(line, flush=True)
        if not (ok_g and ok_u and ok_d and ok_x and ok_i):
            failures.append(line)
    s.close()
    flag = torch.tensor([len(failures)], dtype=torch.int64)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if flag.item():
        sys.exit(1)
    if rank == 0:
        print("MRANK_OK", flush=True)


if __name__ == "__main__":
    main()
