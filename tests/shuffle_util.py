"""Test helper: an "unstructured" renumbering of a generated problem (used by the 1-GPU and N-GPU parity tests)."""
import dataclasses

import numpy as np

from parafem_b200 import host


def shuffled(full, npes, numpe, seed=17):
    """The same problem with random equation numbers and a random element order (seeded, identical on every rank):
    an "unstructured" numbering -- every element touches equations owned all over the ranks, so the halo exchange is
    all-to-all instead of slab neighbours.  Returns rank numpe's share."""
    rng = np.random.RandomState(seed)
    pe = np.concatenate([[0], rng.permutation(full.neq) + 1]).astype(np.int32)      # old equation -> new equation
    order = rng.permutation(full.nels)
    g_g = pe[full.g_g_pp][order]
    r = np.empty(full.neq)
    r[pe[1:] - 1] = full.r_pp
    nf = pe[full.nf]
    nels_pp, iel_start = host.calc_nels_pp(full.nels, npes, numpe)
    neq_pp, ieq_start = host.calc_neq_pp(full.neq, npes, numpe)
    e0 = iel_start - 1
    no_f = pe[full.no_f] if full.no_f.size else full.no_f
    mine = (no_f >= ieq_start) & (no_f < ieq_start + neq_pp) if no_f.size else np.zeros(0, bool)
    return dataclasses.replace(
        full, npes=npes, numpe=numpe, nels_pp=nels_pp, iel_start=iel_start, neq_pp=neq_pp, ieq_start=ieq_start,
        g_num_pp=np.ascontiguousarray(full.g_num_pp[order][e0:e0 + nels_pp]),
        g_coord_pp=np.ascontiguousarray(full.g_coord_pp[order][e0:e0 + nels_pp]),
        g_g_pp=np.ascontiguousarray(g_g[e0:e0 + nels_pp]), nf=nf,
        r_pp=np.ascontiguousarray(r[ieq_start - 1:ieq_start - 1 + neq_pp]),
        no_f=np.ascontiguousarray(no_f[mine]) if no_f.size else no_f,
        val_f=np.ascontiguousarray(full.val_f[mine]) if no_f.size else full.val_f)
