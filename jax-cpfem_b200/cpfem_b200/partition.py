"""Element partition of a hex8 mesh over ranks and the interface exchange that follows an assembly.

Design (DESIGN.md "Multi-GPU"):
  * cells are split into contiguous ranges (z-slabs for the structured meshes of the reference); the constitutive
    update needs no communication at all;
  * a node belongs to the LOWEST rank whose cells touch it; the global CSR is row-partitioned by that owner;
  * each rank's local mesh = its owned cells followed by GHOST cells (cells of other ranks that touch one of its
    owned nodes).  Ghost cells only contribute their sparsity pattern, so that the owner's rows already have a slot
    for every column a neighbour will send; kernels run over the owned cells only (n_active);
  * after the local assembly every rank sends, per neighbour, the residual entries and CSR rows of the nodes that
    neighbour owns (packed in the sender's row order) and the owner adds them through a precomputed slot map.
    One scalar allreduce gives the global residual norm over owned rows.

The exchange is index plumbing around torch.distributed (NCCL on GPUs, gloo in the CPU tests); on CUDA tensors the
pack / unpack-add / sum-of-squares steps are the library's own kernels (cpfem_gather / cpfem_scatter_add /
cpfem_sumsq), on CPU tensors (tests only) they are torch index ops.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List, Optional

import numpy as onp
import torch


@dataclasses.dataclass
class RankMesh:
    rank: int
    world: int
    cells: onp.ndarray            # (n_local_cells, 8) int32, LOCAL node ids; owned cells first, ghost cells last
    n_owned_cells: int
    points: onp.ndarray           # (n_local_nodes, 3)
    node_gid: onp.ndarray         # (n_local_nodes,) int64 ascending
    node_owner: onp.ndarray       # (n_local_nodes,) int32 owner rank of each local node
    cell_gid: onp.ndarray         # (n_local_cells,) int64
    send_nodes: Dict[int, onp.ndarray]   # peer -> local node ids (ascending gid) whose owner is `peer`
    recv_nodes: Dict[int, onp.ndarray]   # peer -> local node ids (ascending gid) owned here, touched by peer's cells
    n_global_nodes: int

    @property
    def owned_node_mask(self):
        return self.node_owner == self.rank


def cell_ranges(nc, world):
    """Contiguous, balanced cell ranges."""
    b = [(nc * r) // world for r in range(world + 1)]
    return b


def partition_cells(cells, points, world, rank, bounds=None) -> RankMesh:
    """Generic partition of any hex8 mesh by contiguous cell ranges (host numpy)."""
    cells = onp.asarray(cells, dtype=onp.int64)
    nc, nn = len(cells), len(points)
    bounds = cell_ranges(nc, world) if bounds is None else list(bounds)
    rank_of_cell = onp.searchsorted(onp.asarray(bounds[1:]), onp.arange(nc), side='right').astype(onp.int32)
    owner = onp.full(nn, world, dtype=onp.int32)
    onp.minimum.at(owner, cells.reshape(-1), onp.repeat(rank_of_cell, 8))
    c0, c1 = bounds[rank], bounds[rank + 1]
    owned = onp.arange(c0, c1)
    touches_mine = (owner[cells] == rank).any(axis=1)
    ghost = onp.nonzero(touches_mine & (rank_of_cell != rank))[0]
    cell_gid = onp.concatenate([owned, ghost])
    lc = cells[cell_gid]
    node_gid = onp.unique(lc)
    local = onp.searchsorted(node_gid, lc).astype(onp.int32)
    # nodes touched by my OWNED cells but owned elsewhere -> send to owner
    mine_touched = onp.unique(cells[owned])
    send, recv = {}, {}
    own_t = owner[mine_touched]
    for p in onp.unique(own_t):
        if p != rank:
            send[int(p)] = onp.searchsorted(node_gid, mine_touched[own_t == p]).astype(onp.int64)
    # nodes I own that other ranks' owned cells touch -> receive from them
    for p in range(world):
        if p == rank:
            continue
        pt = onp.unique(cells[bounds[p]:bounds[p + 1]])
        sel = pt[owner[pt] == rank]
        if len(sel):
            recv[p] = onp.searchsorted(node_gid, sel).astype(onp.int64)
    return RankMesh(rank, world, local, len(owned), onp.asarray(points)[node_gid], node_gid.astype(onp.int64),
                    owner[node_gid], cell_gid.astype(onp.int64), send, recv, nn)


def slab_partition_structured(N, world, rank, lengths=(1., 1., 1.)) -> RankMesh:
    """Analytic z-slab partition of the N^3 box mesh (generate_mesh.box_mesh numbering) - same result as
    partition_cells(box_mesh(N,N,N), bounds = whole layers) without ever building the global mesh."""
    layers = [(N * r) // world for r in range(world + 1)]
    k0, k1 = layers[rank], layers[rank + 1]
    has_ghost = rank < world - 1
    kk1 = k1 + (1 if has_ghost else 0)               # cell layers in the local mesh: [k0, kk1)
    sx, sy = N + 1, (N + 1) ** 2
    xs = onp.linspace(0, lengths[0], N + 1)
    ys = onp.linspace(0, lengths[1], N + 1)
    zs = onp.linspace(0, lengths[2], N + 1)[k0:kk1 + 1]
    Z, Y, X = onp.meshgrid(zs, ys, xs, indexing='ij')
    points = onp.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    n0 = k0 * sy
    node_gid = onp.arange(n0, (kk1 + 1) * sy, dtype=onp.int64)
    k, j, i = onp.meshgrid(onp.arange(kk1 - k0, dtype=onp.int64), onp.arange(N, dtype=onp.int64),
                           onp.arange(N, dtype=onp.int64), indexing='ij')
    b = (i + sx * j + sy * k).ravel()
    cells = onp.stack([b, b + 1, b + 1 + sx, b + sx, b + sy, b + 1 + sy, b + 1 + sx + sy, b + sx + sy], axis=1).astype(onp.int32)
    cell_gid = onp.arange(k0 * N * N, kk1 * N * N, dtype=onp.int64)
    plane = (node_gid - n0) // sy + k0                # global node-plane index of each local node
    owner = onp.full(len(node_gid), rank, dtype=onp.int32)
    if rank > 0:
        owner[plane == k0] = rank - 1
    if has_ghost:
        # plane k1 is shared with rank+1 and owned here; the ghost plane k1+1 belongs to rank+1 (its lowest toucher)
        owner[plane == kk1] = rank + 1
    send, recv = {}, {}
    if rank > 0:
        send[rank - 1] = onp.nonzero(plane == k0)[0].astype(onp.int64)
    if has_ghost:
        recv[rank + 1] = onp.nonzero(plane == k1)[0].astype(onp.int64)
    return RankMesh(rank, world, cells, (k1 - k0) * N * N, points, node_gid, owner, cell_gid, send, recv, (N + 1) ** 3)


# ----------------------------------------------------------------------------------------------------------
# exchange plan
# ----------------------------------------------------------------------------------------------------------
def _rows_of_nodes(nodes: torch.Tensor):
    return (3 * nodes[:, None] + torch.arange(3, device=nodes.device)[None, :]).reshape(-1)


def _slots_of_rows(indptr: torch.Tensor, rows: torch.Tensor):
    """Concatenated CSR slot ranges of `rows` (+ per-row lengths)."""
    start = indptr[rows]
    length = indptr[rows + 1] - start
    total = int(length.sum())
    off = torch.cumsum(length, 0) - length
    rid = torch.repeat_interleave(torch.arange(len(rows), device=rows.device), length, output_size=total)
    slots = start[rid] + (torch.arange(total, device=rows.device) - off[rid])
    return slots, length, rid


def _is_contiguous_range(idx: torch.Tensor):
    if idx.numel() == 0:
        return True
    return bool((idx[-1] - idx[0] + 1 == idx.numel()).item()) and bool((idx[1:] - idx[:-1] == 1).all().item())


class ExchangePlan:
    """Send/receive maps of one rank, built once per mesh.  indptr/indices: the LOCAL CSR pattern (local column ids);
    pg: torch.distributed process group (None = default)."""

    def __init__(self, rm: RankMesh, indptr: torch.Tensor, indices: torch.Tensor, pg=None):
        import torch.distributed as dist
        self.rm, self.pg = rm, pg
        dev = indptr.device
        self.device = dev
        gid = torch.as_tensor(rm.node_gid, device=dev)
        self.send_rows, self.send_slots, self.recv_rows, self.recv_slots = {}, {}, {}, {}
        self.send_contig = {}
        ndof_g = 3 * rm.n_global_nodes
        meta_send = {}
        for p, nodes in sorted(rm.send_nodes.items()):
            rows = _rows_of_nodes(torch.as_tensor(nodes, device=dev))
            slots, length, rid = _slots_of_rows(indptr, rows)
            self.send_rows[p], self.send_slots[p] = rows, slots
            self.send_contig[p] = (_is_contiguous_range(rows), _is_contiguous_range(slots))
            # keys = global_row * ndof_g + global_col of every entry sent
            cols_l = indices[slots].to(torch.int64)
            gcol = 3 * gid[cols_l // 3] + cols_l % 3
            grow = (3 * gid[rows // 3] + rows % 3)[rid]
            meta_send[p] = grow * ndof_g + gcol
        # exchange the key lists (sizes first)
        peers_s, peers_r = sorted(rm.send_nodes), sorted(rm.recv_nodes)
        sizes_r = {p: torch.zeros(1, dtype=torch.int64, device=dev) for p in peers_r}
        ops = [dist.P2POp(dist.isend, torch.tensor([meta_send[p].numel()], dtype=torch.int64, device=dev), p, group=pg) for p in peers_s]
        ops += [dist.P2POp(dist.irecv, sizes_r[p], p, group=pg) for p in peers_r]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        keys_r = {p: torch.empty(int(sizes_r[p].item()), dtype=torch.int64, device=dev) for p in peers_r}
        ops = [dist.P2POp(dist.isend, meta_send[p], p, group=pg) for p in peers_s]
        ops += [dist.P2POp(dist.irecv, keys_r[p], p, group=pg) for p in peers_r]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for p in peers_r:
            rows = _rows_of_nodes(torch.as_tensor(rm.recv_nodes[p], device=dev))
            slots, length, rid = _slots_of_rows(indptr, rows)
            cols_l = indices[slots].to(torch.int64)
            mykeys = (3 * gid[rows // 3] + rows % 3)[rid] * ndof_g + (3 * gid[cols_l // 3] + cols_l % 3)
            pos = torch.searchsorted(mykeys, keys_r[p])
            if pos.numel() and (int(pos.max()) >= mykeys.numel() or not bool((mykeys[pos] == keys_r[p]).all())):
                raise RuntimeError('ExchangePlan: a neighbour sends a CSR entry this rank has no slot for '
                                   '(ghost cells missing from the local pattern)')
            self.recv_slots[p] = slots[pos]
            # residual rows arrive in the sender's row order = ascending global dof = our `rows` order
            self.recv_rows[p] = rows
        owned = torch.as_tensor(rm.owned_node_mask, device=dev)
        self.owned_rows = _rows_of_nodes(torch.nonzero(owned).reshape(-1))
        self.owned_contig = _is_contiguous_range(self.owned_rows)
        self._bufs = {}

    # ---- runtime -------------------------------------------------------------------------------------
    def _gather(self, src, idx, contig):
        if contig and idx.numel():
            return src[int(idx[0]):int(idx[0]) + idx.numel()]
        if src.is_cuda:
            from . import api
            out = torch.empty(idx.numel(), dtype=src.dtype, device=src.device)
            api.gather(src, idx, out)
            return out
        return src[idx]

    def _scatter_add(self, dst, idx, val):
        if dst.is_cuda:
            from . import api
            api.scatter_add(val, idx, dst)
        else:
            dst.index_add_(0, idx, val)

    def prepare(self):
        """Cache python ints for the contiguous fast paths (no device syncs at exchange time)."""
        self._s0 = {p: (int(self.send_rows[p][0]), self.send_rows[p].numel(), int(self.send_slots[p][0]), self.send_slots[p].numel())
                    for p in self.send_rows}
        self._own = (int(self.owned_rows[0]), self.owned_rows.numel()) if self.owned_rows.numel() else (0, 0)

    def exchange(self, res: torch.Tensor, csr_data: Optional[torch.Tensor]):
        """Adds the neighbours' interface contributions into this rank's owned rows (res: (nn,3) or flat)."""
        import torch.distributed as dist
        if not hasattr(self, '_s0'):
            self.prepare()
        r = res.reshape(-1)
        ops, recv_bufs = [], {}
        keep = []
        for p in sorted(self.send_rows):
            r0, rn, s0, sn = self._s0[p]
            rc, sc = self.send_contig[p]
            sb = r[r0:r0 + rn] if rc else self._gather(r, self.send_rows[p], False)
            ops.append(dist.P2POp(dist.isend, sb, p, group=self.pg))
            keep.append(sb)
            if csr_data is not None:
                cb = csr_data[s0:s0 + sn] if sc else self._gather(csr_data, self.send_slots[p], False)
                ops.append(dist.P2POp(dist.isend, cb, p, group=self.pg))
                keep.append(cb)
        for p in sorted(self.recv_rows):
            key = (p, csr_data is not None)
            if key not in self._bufs:
                self._bufs[key] = (torch.empty(self.recv_rows[p].numel(), dtype=r.dtype, device=r.device),
                                   torch.empty(self.recv_slots[p].numel(), dtype=r.dtype, device=r.device) if csr_data is not None else None)
            rb, cb = self._bufs[key]
            ops.append(dist.P2POp(dist.irecv, rb, p, group=self.pg))
            if csr_data is not None:
                ops.append(dist.P2POp(dist.irecv, cb, p, group=self.pg))
            recv_bufs[p] = (rb, cb)
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for p, (rb, cb) in recv_bufs.items():
            self._scatter_add(r, self.recv_rows[p], rb)
            if csr_data is not None:
                self._scatter_add(csr_data, self.recv_slots[p], cb)

    # ---- overlap with the assembly (CUDA only) -----------------------------------------------------------------
    def send_cell_prefix(self):
        """Number of leading owned cells after which every row this rank sends is complete (the last owned cell that
        touches a node owned by a peer, + 1; one layer of cells for a z-slab)."""
        rm = self.rm
        if not rm.send_nodes:
            return 0
        mark = onp.zeros(len(rm.node_gid), dtype=bool)
        for nodes in rm.send_nodes.values():
            mark[onp.asarray(nodes)] = True
        touch = onp.nonzero(mark[rm.cells[:rm.n_owned_cells]].any(axis=1))[0]
        return int(touch[-1]) + 1 if len(touch) else 0

    def attach(self, plan):
        """Overlap the exchange with the assembly: `plan.newton_update` signals (cpfem_plan_set_progress_event) when the
        cells feeding this rank's send rows are done - for a slab that is its first layer of cells, i.e. the end of the
        first assembly chunk - and `exchange_overlapped` then runs send / receive / add on a high-priority stream
        beside the remaining chunks.  Contributions of peers are added with atomics (cpfem_scatter_add), like the
        element kernel's own, so the two may interleave."""
        self.prepare()
        self._plan = plan
        with torch.cuda.device(self.device):
            self._comm = torch.cuda.Stream(device=self.device, priority=-1)
            self._ready = torch.cuda.Event()
            self._ready.record()                          # creates the handle the library records from now on
            self._done = torch.cuda.Event()
        plan.set_progress_event(max(1, self.send_cell_prefix()), self._ready)

    def detach(self):
        if getattr(self, '_plan', None) is not None:
            self._plan.set_progress_event(0, None)
            self._plan = None

    def exchange_overlapped(self, res: torch.Tensor, csr_data: Optional[torch.Tensor]):
        """Call right after `plan.newton_update(..., res=res, csr_data=csr_data)` of the attached plan was enqueued on
        the current stream; when this returns the current stream is ordered after the exchange."""
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(self._ready)
            self.exchange(res, csr_data)
            self._done.record(self._comm)
        res.record_stream(self._comm)
        if csr_data is not None:
            csr_data.record_stream(self._comm)
        cur.wait_event(self._done)

    # ---- peer-memory path (CUDA, one node): the sender stores into the owner's mailbox over NVLink --------------------
    def attach_peer(self, with_csr=True, timeout_s=20.0):
        """Collective.  Allocates this rank's mailbox (cpfem_peer_alloc: room for what every neighbour sends, behind two
        arrays of 64-bit epoch flags), ships its IPC handle to the neighbours and maps theirs.  After this
        `exchange_peer` replaces `exchange`: no NCCL call, no rendezvous, no receive-side copy - the sender's copy kernel
        (cpfem_peer_put) writes the rows into the owner's memory and releases a flag the owner's stream waits on."""
        import ctypes
        import torch.distributed as dist
        from . import _lib
        from ._lib import check
        L = _lib.lib()
        dev = self.device
        me, world = self.rm.rank, self.rm.world
        self.prepare()
        peers_s, peers_r = sorted(self.send_rows), sorted(self.recv_rows)
        head = ((2 * world * 8 + 255) // 256) * 256                      # data flags [src], ack flags [dst]
        off, o = {}, head // 8
        for q in peers_r:
            off[q] = o
            o += self.recv_rows[q].numel() + (self.recv_slots[q].numel() if with_csr else 0)
            o = (o + 1) & ~1                                              # 16-byte aligned segments
        nbytes = max(o * 8, head)
        def agree(ok, what):
            # every step that can fail locally is followed by a vote, so that all ranks leave together (a rank that raised
            # on its own would leave the others waiting in the next collective)
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.pg)
            if not bool(flag.item()):
                raise RuntimeError(f'attach_peer: {what} failed on ' + ('this rank: ' + str(err[0]) if not ok else 'another rank'))

        err = [None]
        with torch.cuda.device(dev):
            ptr = ctypes.c_void_p()
            hb = (ctypes.c_uint8 * 64)()
            try:
                check(L.cpfem_peer_alloc(nbytes, ctypes.byref(ptr), hb), 'cpfem_peer_alloc')
            except Exception as e:                                       # noqa: BLE001
                err[0] = e
            if os.environ.get('CPFEM_PEER_FAIL_RANK') == str(me):        # fault injection for tests/multigpu_check.py
                if err[0] is None:
                    L.cpfem_peer_free(ptr.value)
                err[0] = RuntimeError('injected failure (CPFEM_PEER_FAIL_RANK)')
            try:
                agree(err[0] is None, 'mailbox allocation')
            except RuntimeError:
                if err[0] is None:
                    L.cpfem_peer_free(ptr.value)
                raise
            mine = torch.tensor(list(hb), dtype=torch.uint8, device=dev)
            allh = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(allh, mine, group=self.pg)
            # the receiver tells every sender where its segment starts
            offs_s = {q: torch.zeros(1, dtype=torch.int64, device=dev) for q in peers_s}
            ops = [dist.P2POp(dist.isend, torch.tensor([off[q]], dtype=torch.int64, device=dev), q, group=self.pg) for q in peers_r]
            ops += [dist.P2POp(dist.irecv, offs_s[q], q, group=self.pg) for q in peers_s]
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            torch.cuda.synchronize()
            remote = {}
            try:
                for q in sorted(set(peers_s) | set(peers_r)):
                    rp = ctypes.c_void_p()
                    hq = (ctypes.c_uint8 * 64)(*allh[q].cpu().tolist())
                    check(L.cpfem_peer_open(hq, ctypes.byref(rp)), 'cpfem_peer_open')
                    remote[q] = rp.value
            except Exception as e:                                       # noqa: BLE001
                err[0] = e
            try:
                agree(err[0] is None, 'mapping a neighbour\'s mailbox (CUDA IPC)')
            except RuntimeError:
                for rp in remote.values():
                    L.cpfem_peer_close(rp)
                L.cpfem_peer_free(ptr.value)
                raise
            self._peer = dict(ptr=ptr.value, remote=remote, off=off, off_s={q: int(offs_s[q].item()) for q in peers_s},
                              with_csr=with_csr, world=world, me=me, timeout=float(timeout_s),
                              status=torch.zeros(4, dtype=torch.int64, device=dev))
            self._epoch = 0
            dist.barrier(group=self.pg)          # every mailbox is mapped before anybody writes

    def detach_peer(self):
        pe = getattr(self, '_peer', None)
        if pe is None:
            return
        import torch.distributed as dist
        from . import _lib
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.pg)              # nobody still writes into a mailbox that is about to go
        L = _lib.lib()
        for rp in pe['remote'].values():
            L.cpfem_peer_close(rp)
        L.cpfem_peer_free(pe['ptr'])
        self._peer = None

    def peer_timeouts(self):
        """Number of cpfem_peer_wait calls that gave up (must be 0)."""
        return int(self._peer['status'][1])

    def exchange_peer(self, res: torch.Tensor, csr_data: Optional[torch.Tensor]):
        """Same result as `exchange` through the mailboxes of `attach_peer` (enqueue-only, current stream)."""
        import ctypes
        from . import _lib
        from ._lib import check
        from .api import _ptr, _stream
        L, pe = _lib.lib(), self._peer
        assert (csr_data is not None) <= pe['with_csr'], 'attach_peer(with_csr=True) needed to ship CSR rows'
        r = res.reshape(-1)
        self._epoch += 1
        e, me, W = self._epoch, pe['me'], pe['world']
        vp = ctypes.c_void_p
        st, to, stat = _stream(), pe['timeout'], _ptr(pe['status'])
        with torch.cuda.device(self.device):
            for q in sorted(self.send_rows):
                r0, rn, s0, sn = self._s0[q]
                rc, sc = self.send_contig[q]
                base = pe['remote'][q]
                dst = base + 8 * pe['off_s'][q]
                flag = vp(base + 8 * me)                                      # data flag [src = me] in q's mailbox
                if e > 1:                                                     # q has consumed what the last put left there
                    check(L.cpfem_peer_wait(vp(pe['ptr'] + 8 * (W + q)), e - 1, to, stat, st), 'cpfem_peer_wait')
                last = csr_data is None
                check(L.cpfem_peer_put(vp(dst), vp(r.data_ptr() + 8 * r0) if rc else _ptr(r), None if rc else _ptr(self.send_rows[q]), rn,
                                       flag if last else None, e, st), 'cpfem_peer_put')
                if not last:
                    check(L.cpfem_peer_put(vp(dst + 8 * rn), vp(csr_data.data_ptr() + 8 * s0) if sc else _ptr(csr_data),
                                           None if sc else _ptr(self.send_slots[q]), sn, flag, e, st), 'cpfem_peer_put')
            for q in sorted(self.recv_rows):
                src = pe['ptr'] + 8 * pe['off'][q]
                nr = self.recv_rows[q].numel()
                check(L.cpfem_peer_wait(vp(pe['ptr'] + 8 * q), e, to, stat, st), 'cpfem_peer_wait')
                check(L.cpfem_scatter_add(vp(src), _ptr(self.recv_rows[q]), nr, _ptr(r), st), 'cpfem_scatter_add')
                if csr_data is not None:
                    check(L.cpfem_scatter_add(vp(src + 8 * nr), _ptr(self.recv_slots[q]), self.recv_slots[q].numel(), _ptr(csr_data), st),
                          'cpfem_scatter_add')
                check(L.cpfem_peer_signal(vp(pe['remote'][q] + 8 * (W + me)), e, st), 'cpfem_peer_signal')   # ack flag [dst = me] at q

    def owned_sumsq(self, res: torch.Tensor, out: Optional[torch.Tensor] = None):
        """sum of squares of the residual over the rows this rank owns (device scalar)."""
        if not hasattr(self, '_s0'):
            self.prepare()
        r = res.reshape(-1)
        o0, on = self._own
        seg = r[o0:o0 + on] if self.owned_contig else r[self.owned_rows]
        if r.is_cuda:
            from . import api
            if out is None:
                out = torch.zeros(1, dtype=torch.float64, device=r.device)
            else:
                out.zero_()
            api.sumsq(seg.contiguous(), out)
            return out
        return (seg * seg).sum().reshape(1)

    def global_res_norm(self, res: torch.Tensor):
        import torch.distributed as dist
        s = self.owned_sumsq(res).clone()
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=self.pg)
        return torch.sqrt(s)


# ----------------------------------------------------------------------------------------------------------
# distributed linear solve on the row-partitioned matrix (beyond the reference, whose jax_solve is single-device)
# ----------------------------------------------------------------------------------------------------------
class HaloPlan:
    """Before a sparse matrix-vector product every rank needs the vector entries of the local nodes it does not own
    (interface nodes of lower ranks, nodes of the ghost cells): `update(v)` fetches them from their owners.  Built once
    per partition from RankMesh.node_owner / node_gid: each rank tells every owner which global nodes it needs."""

    def __init__(self, rm: RankMesh, device, pg=None):
        import torch.distributed as dist
        self.rm, self.pg, self.device = rm, pg, torch.device(device)
        owner = onp.asarray(rm.node_owner)
        gid = onp.asarray(rm.node_gid)
        self.need = {}                                   # peer -> local dof indices I receive (nodes owned by peer)
        need_gid = {}
        for p in sorted(set(owner.tolist()) - {rm.rank}):
            loc = onp.nonzero(owner == p)[0]
            self.need[int(p)] = _rows_of_nodes(torch.as_tensor(loc, device=self.device))
            need_gid[int(p)] = torch.as_tensor(gid[loc], device=self.device)
        # who needs what from me: every rank announces its request sizes to every other rank, then the gid lists
        world = rm.world
        sizes = torch.zeros(world, dtype=torch.int64, device=self.device)
        for p, g in need_gid.items():
            sizes[p] = g.numel()
        all_sizes = [torch.zeros(world, dtype=torch.int64, device=self.device) for _ in range(world)]
        dist.all_gather(all_sizes, sizes, group=pg)
        wanted = {p: int(all_sizes[p][rm.rank]) for p in range(world) if p != rm.rank and int(all_sizes[p][rm.rank]) > 0}
        bufs = {p: torch.empty(n, dtype=torch.int64, device=self.device) for p, n in wanted.items()}
        ops = [dist.P2POp(dist.isend, need_gid[p], p, group=pg) for p in sorted(need_gid)]
        ops += [dist.P2POp(dist.irecv, bufs[p], p, group=pg) for p in sorted(bufs)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        gid_t = torch.as_tensor(gid, device=self.device)
        self.give = {}                                   # peer -> local dof indices (nodes I own) whose values I send
        for p, g in bufs.items():
            loc = torch.searchsorted(gid_t, g)
            if not bool((gid_t[loc] == g).all()) or not bool(torch.as_tensor(owner, device=self.device)[loc].eq(rm.rank).all()):
                raise RuntimeError('HaloPlan: a neighbour requests a node this rank does not own')
            self.give[p] = _rows_of_nodes(loc)
        self._recv = {p: torch.empty(idx.numel(), dtype=torch.float64, device=self.device) for p, idx in self.need.items()}

    def update(self, v: torch.Tensor):
        import torch.distributed as dist
        ops, keep = [], []
        for p in sorted(self.give):
            sb = v[self.give[p]]
            keep.append(sb)
            ops.append(dist.P2POp(dist.isend, sb, p, group=self.pg))
        for p in sorted(self.need):
            ops.append(dist.P2POp(dist.irecv, self._recv[p], p, group=self.pg))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for p, idx in self.need.items():
            v[idx] = self._recv[p]
        return v


class DistributedBicgstab:
    """jax.scipy.sparse.linalg.bicgstab (solver.py:34-40 semantics: x0, Jacobi, tol, atol, maxiter) on a matrix whose rows
    are partitioned by node owner.  Every rank holds local-length vectors; only the entries of owned rows are meaningful,
    the others are refreshed by the halo exchange before each matrix-vector product.  Dot products are partial sums over
    the owned rows + one all-reduce; the recurrence scalars live on every rank identically.

    matvec(v) -> A_local v (local length; rows of non-owned nodes are ignored).  On the GPU that is
    `lambda v: plan.spmv(csr_data, v)` after ExchangePlan.exchange completed the owned rows."""

    def __init__(self, rm: RankMesh, halo: HaloPlan, pg=None):
        self.rm, self.halo, self.pg = rm, halo, pg
        dev = halo.device
        own = torch.nonzero(torch.as_tensor(rm.owned_node_mask, device=dev)).reshape(-1)
        self.owned = _rows_of_nodes(own)
        self.contig = _is_contiguous_range(self.owned)
        if self.contig and self.owned.numel():
            self.sl = slice(int(self.owned[0]), int(self.owned[0]) + self.owned.numel())

    def _o(self, v):
        return v[self.sl] if self.contig else v[self.owned]

    def _allsum(self, *vals):
        import torch.distributed as dist
        t = torch.stack(list(vals))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
        return t

    def solve(self, matvec, b, x0=None, minv=None, tol=1e-10, atol=1e-10, maxiter=10000):
        """Returns (x, k, err): x local-length (owned entries + fresh halo), k iterations (JAX's negative breakdown codes
        passed through), err = global ||A x - b|| over owned rows."""
        o = self._o
        dot = lambda a, c: torch.dot(o(a), o(c))
        x = torch.zeros_like(b) if x0 is None else x0.clone()
        M = (lambda v: v) if minv is None else (lambda v: v * minv)
        bs = float(self._allsum(dot(b, b))[0])
        atol2 = max(tol ** 2 * bs, atol ** 2)
        self.halo.update(x)
        r = b - matvec(x)
        rhat, p, q = r.clone(), r.clone(), r.clone()
        rho = alpha = omega = 1.0
        t2 = self._allsum(dot(r, r), dot(rhat, r))
        rs, rho_ = float(t2[0]), float(t2[1])
        k = 0
        while (rs > atol2) and (k < maxiter) and (k >= 0):
            beta = rho_ / rho * alpha / omega
            p = r + beta * (p - omega * q)
            phat = M(p)
            self.halo.update(phat)
            q = matvec(phat)
            alpha_ = rho_ / float(self._allsum(dot(rhat, q))[0])
            s = r - alpha_ * q
            exit_early = float(self._allsum(dot(s, s))[0]) < atol2
            shat = M(s)
            self.halo.update(shat)
            t = matvec(shat)
            t2 = self._allsum(dot(t, s), dot(t, t))
            omega_ = float(t2[0]) / float(t2[1])
            if exit_early:
                x = x + alpha_ * phat
                r = s
            else:
                x = x + (alpha_ * phat + omega_ * shat)
                r = s - omega_ * t
            k_ = -11 if (omega_ == 0 or alpha_ == 0) else k + 1
            if rho_ == 0:
                k_ = -10
            alpha, omega, rho, k = alpha_, omega_, rho_, k_
            t2 = self._allsum(dot(r, r), dot(rhat, r))
            rs, rho_ = float(t2[0]), float(t2[1])
        self.halo.update(x)
        res = matvec(x) - b
        err = float(torch.sqrt(self._allsum(dot(res, res))[0]))
        return x, k, err
