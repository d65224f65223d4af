"""N>1 host logic on CPU: two gloo processes run the halo protocol of device.cu (tables from
pf_make_ggl, wanted-list exchange, forward exchange, reverse exchange with owner-first then
ascending-rank accumulation, all-gather + rank-ordered scalar sum) in numpy and must reproduce
the serial oracle's emulated-rank results exactly."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, psize=None):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from parafem_b200 import host
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = host.cube_p121(5, 6, 4, 20, aa=1., bb=1., cc=1.)
        p = host.cube_p121(5, 6, 4, 20, aa=1., bb=1., cc=1., npes=world, numpe=rank + 1, psize=psize)
        oracle.set_element_partition(psize)        # partitioner 2: the oracle emulates the same uneven ranks
        ggl, halo, get_cnt = host.make_ggl(p)
        # counts matrix, then the wanted equation lists (pf_setup_mesh does this over NCCL)
        cnts = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(cnts, torch.from_numpy(get_cnt.copy()))
        put_cnt = np.array([int(cnts[r][rank]) for r in range(world)])
        get_off = np.concatenate([[0], np.cumsum(get_cnt)])
        put_off = np.concatenate([[0], np.cumsum(put_cnt)])
        reqs, asked = [], [None] * world
        for r in range(world):
            if r == rank:
                continue
            if get_cnt[r]:
                reqs.append(dist.isend(torch.from_numpy(halo[get_off[r]:get_off[r + 1]].copy()), r))
            if put_cnt[r]:
                asked[r] = torch.zeros(int(put_cnt[r]), dtype=torch.int32)
                reqs.append(dist.irecv(asked[r], r))
        for w in reqs:
            w.wait()
        put_slot = [None if a is None else a.numpy().astype(np.int64) - p.ieq_start + 1 for a in asked]

        km = oracle.form_km_elastic(full.g_coord_pp, 20, 8, full.e, full.v)
        rng = np.random.RandomState(5)
        pv = rng.randn(full.neq)
        lo = p.ieq_start - 1
        # forward exchange -> p_ext, then gather
        p_ext = np.zeros(1 + p.neq_pp + halo.size)
        p_ext[1:1 + p.neq_pp] = pv[lo:lo + p.neq_pp]
        reqs, bufs = [], {}
        for r in range(world):
            if r == rank:
                continue
            if put_cnt[r]:
                reqs.append(dist.isend(torch.from_numpy(p_ext[put_slot[r]].copy()), r))
            if get_cnt[r]:
                bufs[r] = torch.zeros(int(get_cnt[r]), dtype=torch.float64)
                reqs.append(dist.irecv(bufs[r], r))
        for w in reqs:
            w.wait()
        for r, b in bufs.items():
            p_ext[1 + p.neq_pp + get_off[r]:1 + p.neq_pp + get_off[r + 1]] = b.numpy()
        pmul = p_ext[ggl]
        e0 = p.iel_start - 1
        ok_gather = np.array_equal(pmul, oracle.gather(full.g_g_pp, pv)[e0:e0 + p.nels_pp])
        # local mat-vec + slot-centric scatter in ascending element order
        ut = oracle.matvec(km[e0:e0 + p.nels_pp], pmul)
        u_ext = np.zeros_like(p_ext)
        for e in range(p.nels_pp):                       # ascending element order
            for k in range(p.ntot):
                if ggl[e, k]:
                    u_ext[ggl[e, k]] += ut[e, k]
        # reverse exchange; owner adds after its own partial sum, sources ascending
        reqs, bufs = [], {}
        for r in range(world):
            if r == rank:
                continue
            if get_cnt[r]:
                reqs.append(dist.isend(torch.from_numpy(u_ext[1 + p.neq_pp + get_off[r]:1 + p.neq_pp + get_off[r + 1]].copy()), r))
            if put_cnt[r]:
                bufs[r] = torch.zeros(int(put_cnt[r]), dtype=torch.float64)
                reqs.append(dist.irecv(bufs[r], r))
        for w in reqs:
            w.wait()
        u_own = u_ext.copy()                              # own partial sums, before the peers' are added
        for r in sorted(bufs):
            np.add.at(u_ext, put_slot[r], bufs[r].numpy())
        # ---- the same two exchanges as the FUSED kernels of the peer transport do them (device.cu: k_pupdate stores
        # into the peers' halo segments through pf_make_put_tables; k_dot accumulates chunk by chunk through
        # pf_make_acc_chunks): identical p_ext halo segments and identical owned sums
        from parafem_b200._lib import lib, ptr
        import ctypes as C
        allc = np.array([[int(cnts[r][o]) for o in range(world)] for r in range(world)])   # allc[r][o]: r gathers from o
        neq_pp_of = [host.calc_neq_pp(full.neq, world, r + 1)[0] for r in range(world)]
        fwd_dst = np.array([1 + neq_pp_of[r] + allc[r][:rank].sum() for r in range(world)], np.int64)
        put_flat = np.concatenate([np.zeros(0, np.int64)] + [ps for ps in put_slot if ps is not None]).astype(np.int32)
        nput = int(put_off[-1])
        bits = np.zeros((p.neq_pp + 31) // 32 + 1, np.uint32)
        slot0, rnk = np.zeros(max(nput, 1), np.int32), np.zeros(max(nput, 1), np.int32)
        pptr, dst = np.zeros(max(nput, 1) + 1, np.uint32), np.zeros(max(nput, 1), np.int64)
        nu = C.c_int64()
        assert lib().pf_make_put_tables(world, p.neq_pp, ptr(put_off.astype(np.int64)), ptr(put_flat), ptr(fwd_dst), ptr(bits),
                                        ptr(slot0), ptr(pptr), ptr(rnk), ptr(dst), C.byref(nu)) == 0
        msgs = []                                          # what k_pupdate's threads store into peer memory
        for i in range(p.neq_pp):
            if (bits[i >> 5] >> (i & 31)) & 1:
                j = int(np.searchsorted(slot0[:nu.value], i))
                assert slot0[j] == i
                for jj in range(pptr[j], pptr[j + 1]):
                    msgs.append((int(rnk[jj]), int(dst[jj]), float(p_ext[1 + i])))
        assert len(msgs) == nput
        everyone = [None] * world
        dist.all_gather_object(everyone, msgs)
        p_ext2 = np.zeros_like(p_ext)
        p_ext2[1:1 + p.neq_pp] = p_ext[1:1 + p.neq_pp]
        for src in everyone:
            for r, d, v in src:
                if r == rank:
                    p_ext2[d] = v
        ok_gather = ok_gather and np.array_equal(p_ext2, p_ext)
        recv = np.zeros(max(nput, 1))                      # receive buffer: grouped by source rank ascending
        for r, b in bufs.items():
            recv[put_off[r]:put_off[r + 1]] = b.numpy()
        pairs = sorted((int(sl), k) for k, sl in enumerate(put_flat))
        aslot, aptr, apos = [], [], []
        for k, (sl, pos) in enumerate(pairs):
            if k == 0 or sl != pairs[k - 1][0]:
                aslot.append(sl); aptr.append(k)
            apos.append(pos)
        aptr.append(len(pairs))
        chunk = 64
        nchunks = (p.neq_pp + chunk - 1) // chunk
        cptr = np.zeros(nchunks + 1, np.uint32)
        assert lib().pf_make_acc_chunks(p.neq_pp, chunk, len(aslot), ptr(np.array(aslot + [0], np.int32)), ptr(cptr)) == 0
        u2 = u_own.copy()
        for c in range(nchunks):                           # k_dot<ACC>: each block adds the entries of its chunk
            for k in range(cptr[c], cptr[c + 1]):
                assert (aslot[k] - 1) // chunk == c
                v = u2[aslot[k]]
                for qq in range(aptr[k], aptr[k + 1]):
                    v = v + recv[apos[qq]]
                u2[aslot[k]] = v
        assert cptr[nchunks] == len(aslot)
        ok_fused_rev = np.array_equal(u2[1:1 + p.neq_pp], u_ext[1:1 + p.neq_pp])
        u_ref = oracle.scatter(full.g_g_pp, oracle.matvec(km, oracle.gather(full.g_g_pp, pv)), full.neq, npes=world)
        ok_scatter = np.array_equal(u_ext[1:1 + p.neq_pp], u_ref[lo:lo + p.neq_pp]) and ok_fused_rev
        # dot: blocked local partial, all-gather, ranks ascending
        part = torch.tensor([oracle.dot_blocked(pv[lo:lo + p.neq_pp], u_ref[lo:lo + p.neq_pp])], dtype=torch.float64)
        parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, part)
        s = float(parts[0])
        for r in range(1, world):
            s = s + float(parts[r])
        ok_dot = s == oracle.dot_ranks(pv, u_ref, npes=world, red_mode=1)
        q.put((rank, bool(ok_gather), bool(ok_scatter), bool(ok_dot), int(halo.size), int(put_cnt.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,psize", [(2, None), (3, None), (2, [85, 35]), (3, [70, 14, 36])])
def test_halo_protocol_over_gloo(world, psize):
    """psize: external element partition (.psize, partitioner 2) -- element and equation cuts far apart,
    so ranks exchange with non-adjacent ranks too."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, psize)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, okg, oks, okd, nh, npu in sorted(res):
        assert okg and oks and okd, (rank, okg, oks, okd)
        assert psize is not None or (nh > 0 and npu > 0)
    # (with a lopsided external partition a rank may own every equation its elements touch)
    assert sum(r[4] for r in res) > 0 and sum(r[4] for r in res) == sum(r[5] for r in res)
