"""CPU suite: element partition + interface exchange (host logic of the multi-GPU path) with world_size 2 and 3
over gloo.  The 'assembly' here is a stand-in with integer-valued element matrices (so sums are exact); the point
is the ownership rule, the ghost-cell pattern, the slot maps and the exchange."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cpfem_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_cell_values(cell_gid):
    """Deterministic integer K_e (24x24) and r_e (24) per GLOBAL cell id."""
    c = cell_gid[:, None, None].astype(np.float64)
    p = np.arange(24)[None, :, None]
    q = np.arange(24)[None, None, :]
    Ke = np.mod(c * 7 + p * 3 + q * 5, 11.0) - 5.0
    re = np.mod(cell_gid[:, None] * 3 + np.arange(24)[None, :], 7.0) - 3.0
    return Ke, re


def _worker(rank, world, port, N, structured, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import sys
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path[:0] = [os.path.join(here, '..', 'jax-cpfem_b200'), os.path.join(here, '..', 'oracle')]
        from cpfem_b200 import partition
        pts, cells = O.box_mesh(N, N, N)
        if structured:
            rm = partition.slab_partition_structured(N, world, rank)
            layers = [(N * r) // world for r in range(world + 1)]
            rm2 = partition.partition_cells(cells, pts, world, rank, bounds=[l * N * N for l in layers])
            assert np.array_equal(rm.cells, rm2.cells) and np.array_equal(rm.node_gid, rm2.node_gid)
            assert rm.n_owned_cells == rm2.n_owned_cells and np.array_equal(rm.cell_gid, rm2.cell_gid)
            assert np.array_equal(rm.node_owner == rank, rm2.node_owner == rank)
            assert np.allclose(rm.points, rm2.points)
            assert sorted(rm.send_nodes) == sorted(rm2.send_nodes) and sorted(rm.recv_nodes) == sorted(rm2.recv_nodes)
            for p in rm.send_nodes:
                assert np.array_equal(rm.send_nodes[p], rm2.send_nodes[p])
            for p in rm.recv_nodes:
                assert np.array_equal(rm.recv_nodes[p], rm2.recv_nodes[p])
        else:
            rm = partition.partition_cells(cells, pts, world, rank)          # ragged ranges cutting through layers
        nl = len(rm.node_gid)
        I, J = O.coo_indices(rm.cells)
        Ke, re = _fake_cell_values(rm.cell_gid)
        Ke[rm.n_owned_cells:] = 0.0                  # ghost cells contribute pattern only
        re[rm.n_owned_cells:] = 0.0
        A = scipy.sparse.csr_array((Ke.reshape(-1), (I, J)), shape=(3 * nl, 3 * nl))
        res = np.zeros((nl, 3))
        np.add.at(res, rm.cells.reshape(-1), re.reshape(-1, 3))
        indptr = torch.as_tensor(A.indptr.astype(np.int64))
        indices = torch.as_tensor(A.indices.astype(np.int32))
        data = torch.as_tensor(A.data.copy())
        res_t = torch.as_tensor(res)
        ex = partition.ExchangePlan(rm, indptr, indices)
        # the overlap trigger: every cell that feeds a row sent to a peer lies in [0, send_cell_prefix)
        pre = ex.send_cell_prefix()
        sent = np.zeros(nl, dtype=bool)
        for nodes in rm.send_nodes.values():
            sent[nodes] = True
        feeds = sent[rm.cells[:rm.n_owned_cells]].any(axis=1)
        assert not feeds[pre:].any() and (pre == 0 or feeds[pre - 1])
        if structured and rank > 0:
            assert pre == N * N                      # one layer of cells of a z-slab
        ex.exchange(res_t, data)
        nrm = ex.global_res_norm(res_t).item()
        # global truth
        Ig, Jg = O.coo_indices(cells)
        Kg, rg = _fake_cell_values(np.arange(len(cells)))
        Ag = scipy.sparse.csr_array((Kg.reshape(-1), (Ig, Jg)), shape=(3 * len(pts), 3 * len(pts)))
        resg = np.zeros((len(pts), 3))
        np.add.at(resg, cells.reshape(-1), rg.reshape(-1, 3))
        own = np.nonzero(rm.owned_node_mask)[0]
        assert np.array_equal(res_t.numpy()[own], resg[rm.node_gid[own]])
        d = data.numpy()
        for n in own:
            for i in range(3):
                lr, gr = 3 * n + i, 3 * rm.node_gid[n] + i
                lc = A.indices[A.indptr[lr]:A.indptr[lr + 1]]
                gcols = 3 * rm.node_gid[lc // 3] + lc % 3
                assert np.array_equal(gcols, Ag.indices[Ag.indptr[gr]:Ag.indptr[gr + 1]])
                assert np.array_equal(d[A.indptr[lr]:A.indptr[lr + 1]], Ag.data[Ag.indptr[gr]:Ag.indptr[gr + 1]])
        assert abs(nrm - np.linalg.norm(resg)) < 1e-12 * np.linalg.norm(resg)
        # every node is owned by exactly one rank
        cnt = torch.zeros(len(pts), dtype=torch.int64)
        cnt[torch.as_tensor(rm.node_gid[own])] = 1
        dist.all_reduce(cnt)
        assert bool((cnt == 1).all())
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,N,structured', [(2, 4, True), (3, 5, True), (2, 3, False), (3, 4, False)])
def test_partition_exchange_gloo(world, N, structured):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, N, structured, ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == list(range(world))


def _solver_worker(rank, world, port, N, structured, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import sys
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path[:0] = [os.path.join(here, '..', 'jax-cpfem_b200'), os.path.join(here, '..', 'oracle')]
        from cpfem_b200 import partition
        pts, cells = O.box_mesh(N, N, N)
        rm = partition.slab_partition_structured(N, world, rank) if structured else partition.partition_cells(cells, pts, world, rank)
        # a global FE-patterned, non-symmetric, diagonally dominant matrix with a few identity (Dirichlet) rows
        rng = np.random.default_rng(11)
        Ig, Jg = O.coo_indices(cells)
        ndof = 3 * len(pts)
        Ag = scipy.sparse.csr_array((rng.normal(size=len(Ig)), (Ig, Jg)), shape=(ndof, ndof)).tolil()
        Ag.setdiag(np.abs(Ag).sum(axis=1).ravel() + 1.0)
        for r_ in range(0, ndof, 17):
            Ag.rows[r_] = [r_]
            Ag.data[r_] = [1.0]
        Ag = Ag.tocsr()
        bg = rng.normal(size=ndof)
        x0g = rng.normal(size=ndof) * 0.1
        gd = (3 * rm.node_gid[:, None] + np.arange(3)[None, :]).reshape(-1)
        Al = Ag[gd][:, gd].tocsr()                                   # local rows x local columns
        matvec = lambda v: torch.as_tensor(Al @ v.numpy())
        halo = partition.HaloPlan(rm, 'cpu')
        sol = partition.DistributedBicgstab(rm, halo)
        minv = torch.as_tensor(1.0 / Ag.diagonal()[gd])
        x, k, err = sol.solve(matvec, torch.as_tensor(bg[gd]), x0=torch.as_tensor(x0g[gd]), minv=minv, tol=1e-10, atol=1e-10, maxiter=500)
        jac = Ag.diagonal()
        xo, ko = O.bicgstab_ref(Ag, bg, x0=x0g, M=lambda v: v * (1. / jac), tol=1e-10, atol=1e-10, maxiter=500)
        own = np.nonzero(rm.owned_node_mask)[0]
        od = (3 * own[:, None] + np.arange(3)[None, :]).reshape(-1)
        assert k > 0 and abs(k - ko) <= 2, (k, ko)
        assert np.abs(x.numpy()[od] - xo[gd][od]).max() < 1e-9 * np.abs(xo).max()
        assert err < 1e-8 * np.linalg.norm(bg)
        # the halo of the returned x is fresh: non-owned local entries equal the owners' values
        assert np.abs(x.numpy() - xo[gd]).max() < 1e-9 * np.abs(xo).max()
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,N,structured', [(2, 4, True), (3, 4, False)])
def test_distributed_bicgstab_gloo(world, N, structured):
    """Row-partitioned Jacobi-BiCGStab (halo exchange + all-reduced dot products) against the serial restatement of
    jax.scipy.sparse.linalg.bicgstab on the global system."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_solver_worker, args=(world, port, N, structured, ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == list(range(world))
