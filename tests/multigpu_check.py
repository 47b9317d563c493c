#!/usr/bin/env python
"""Multi-GPU integration check (run under torchrun with >= 2 GPUs; not collected by pytest, which the driver runs on one GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multigpu_check.py

Every rank assembles its slab of a 304-steel polycrystal with the CUDA path, exchanges the interface rows over NCCL
(partition.ExchangePlan.exchange) and through the peer-memory mailboxes (exchange_peer: CUDA IPC + cpfem_peer_* kernels), applies the Dirichlet rows and solves the row-partitioned system with
partition.DistributedBicgstab (node-block SpMV kernel + halo exchange + all-reduced dot products).  Checked against the
single-GPU assembly and cpfem_bicgstab of the whole mesh: CSR rows of the owned nodes to atomic-summation order, the
Newton increment to solver tolerance, the global residual norm."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('jax-cpfem_b200',):
    sys.path.insert(0, os.path.join(ROOT, p))


def main():
    from cpfem_b200 import Plan, make_material, synthetic, slip_systems
    from cpfem_b200.partition import slab_partition_structured, ExchangePlan, HaloPlan, DistributedBicgstab
    from cpfem_b200.problem import get_rot_mat
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    N = int(os.environ.get('CHECK_N', '24'))
    os.environ.setdefault('CPFEM_CHUNK_CELLS', '2048')     # several assembly chunks per slab, so that the overlap path is real
    mat = make_material(2.622e5, 1.120e5, 0.746e5, 392.9772, 7295.1754, 8.0, 1.0 / 120.0, 1.0, 0.001, 1e-8, 8)
    G = (N + 7) // 8
    Rg = torch.as_tensor(get_rot_mat(synthetic.grain_quaternions(G ** 3, 0)), device=dev)
    noise_all = synthetic.noise_field(N)

    def build(rm):
        plan = Plan(rm.cells, rm.points, slip_systems.FCC12)
        plan.set_active_cells(rm.n_owned_cells)
        nc = rm.n_owned_cells
        cg = torch.as_tensor(rm.cell_gid[:nc], device=dev)
        gid = (cg % N) // 8 + G * (((cg // N) % N) // 8) + G * G * ((cg // (N * N)) // 8)
        rot = Rg[gid][:, None].expand(nc, 8, 3, 3).contiguous()
        state = [torch.eye(3, dtype=torch.float64, device=dev).expand(nc, 8, 3, 3).contiguous(),
                 torch.full((nc, 8, 12), 90.0, dtype=torch.float64, device=dev), torch.zeros(nc, 8, 12, dtype=torch.float64, device=dev), rot]
        pts = torch.as_tensor(rm.points, device=dev)
        noise = torch.as_tensor(noise_all[rm.node_gid], device=dev)
        disp = lambda s: (pts * torch.tensor([-0.3, -0.3, 1.0], dtype=torch.float64, device=dev) * (2e-4 * s) + noise).contiguous()
        for s in range(1, 8):
            new = plan.update_state(mat, disp(s), state, 2e-3)
            state = [new[0], new[1], new[2], rot]
        sol = disp(8)
        res, csr, _ = plan.newton_update(mat, sol, state, 2e-3)
        # clamp the bottom face (z = 0), prescribe z on the top face: rows of the local nodes concerned
        z = pts[:, 2]
        bot = torch.nonzero(z < 1e-9).reshape(-1)
        top = torch.nonzero(z > 1 - 1e-9).reshape(-1)
        rows = torch.cat([3 * bot, 3 * bot + 1, 3 * bot + 2, 3 * top + 2])
        vals = torch.cat([torch.zeros(3 * bot.numel(), dtype=torch.float64, device=dev), torch.full((top.numel(),), 2e-4 * 8, dtype=torch.float64, device=dev)])
        return plan, sol, res, csr, rows, vals, state

    # ---- partitioned ----
    # CHECK_GENERIC=1: the same mesh with its node ids permuted, cut by partition_cells - interface rows and CSR slots are
    # then NOT contiguous ranges, which exercises the gather path of the exchange (cpfem_gather / the map of cpfem_peer_put)
    generic = bool(os.environ.get('CHECK_GENERIC'))
    if generic:
        from cpfem_b200.partition import partition_cells
        from cpfem_b200.generate_mesh import box_mesh
        mm = box_mesh(N, N, N, 1., 1., 1.)
        perm = np.random.default_rng(5).permutation(len(mm.points))          # new id of old node i = perm[i]
        gpts = np.empty_like(mm.points)
        gpts[perm] = mm.points
        gcells = perm[mm.cells_dict['hexahedron']]
        noise_all = np.ascontiguousarray(noise_all[np.argsort(perm)])         # noise follows the node, not the id
        part = lambda w, r: partition_cells(gcells, gpts, w, r)
    else:
        part = lambda w, r: slab_partition_structured(N, w, r)
    rm = part(world, rank)
    plan, sol, res, csr, rows, vals, state = build(rm)
    ip, ix = plan.csr_pattern()
    ex = ExchangePlan(rm, ip, ix)
    ex.exchange(res, csr)
    # the same exchange overlapped with the assembly (progress event after the first chunk, high-priority stream):
    # identical up to the summation order of the atomics
    ex.attach(plan)
    e_ov = 0.0
    for _ in range(3):
        res2, csr2, _ = plan.newton_update(mat, sol, state, 2e-3)
        ex.exchange_overlapped(res2, csr2)
        e_ov = max(e_ov, float((res2 - res).abs().max() / csr.abs().max()), float((csr2 - csr).abs().max() / csr.abs().max()))
    ex.detach()
    # the same exchange through the peer-memory mailboxes (CUDA IPC, cpfem_peer_put / _wait / _signal, no NCCL call):
    # several epochs in a row (mailbox reuse + acknowledgement), residual-only and residual + CSR
    ex.attach_peer(with_csr=True)
    e_pm = 0.0
    t_pm = []
    for it in range(4):
        res3, csr3, _ = plan.newton_update(mat, sol, state, 2e-3)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        ex.exchange_peer(res3, csr3)
        t1.record()
        torch.cuda.synchronize()
        t_pm.append(t0.elapsed_time(t1))
        e_pm = max(e_pm, float((res3 - res).abs().max() / csr.abs().max()), float((csr3 - csr).abs().max() / csr.abs().max()))
    res4 = plan.residual(mat, sol, state, 2e-3)
    ex.exchange_peer(res4, None)
    torch.cuda.synchronize()
    e_pm = max(e_pm, float((res4 - res).abs().max() / csr.abs().max()))
    n_to = ex.peer_timeouts()
    ex.detach_peer()
    # a rank that cannot set its mailbox up must take every rank out together (no rank left waiting in a collective)
    os.environ['CPFEM_PEER_FAIL_RANK'] = str(world - 1)
    try:
        ex.attach_peer(with_csr=True)
        raised = False
    except RuntimeError:
        raised = True
    del os.environ['CPFEM_PEER_FAIL_RANK']
    assert raised and getattr(ex, '_peer', None) is None, 'attach_peer: injected failure did not raise on every rank'
    if rows.numel():
        plan.apply_dirichlet(rows, vals, sol.reshape(-1), res=res.reshape(-1), csr_data=csr)
    nrm = float(ex.global_res_norm(res))
    halo = HaloPlan(rm, dev)
    solver = DistributedBicgstab(rm, halo)
    minv = plan.csr_diagonal(csr, invert=True)
    owned = torch.as_tensor(np.repeat(rm.owned_node_mask, 3), device=dev)
    minv = torch.where(owned, minv, torch.zeros_like(minv))
    x, k, err = solver.solve(lambda v: plan.spmv(csr, v), -res.reshape(-1), minv=minv, tol=1e-10, atol=1e-10, maxiter=10000)

    # ---- whole mesh on this GPU ----
    whole = part(1, 0)
    gplan, gsol, gres, gcsr, grows, gvals, _ = build(whole)
    gplan.apply_dirichlet(grows, gvals, gsol.reshape(-1), res=gres.reshape(-1), csr_data=gcsr)
    gx, gk, gerr = gplan.bicgstab(gcsr, -gres.reshape(-1))
    gnrm = float(torch.linalg.norm(gres))
    # compare on the owned nodes
    own = np.nonzero(rm.owned_node_mask)[0]
    gd = torch.as_tensor((3 * rm.node_gid[own][:, None] + np.arange(3)[None, :]).reshape(-1), device=dev)
    ld = torch.as_tensor((3 * own[:, None] + np.arange(3)[None, :]).reshape(-1), device=dev)
    e_x = float((x[ld] - gx[gd]).abs().max() / gx.abs().max())
    e_r = float((res.reshape(-1)[ld] - gres.reshape(-1)[gd]).abs().max() / gcsr.abs().max())
    gip, _ = gplan.csr_pattern()
    worst = 0.0
    for n_l, n_g in list(zip(own, rm.node_gid[own]))[:: max(1, len(own) // 200)]:
        for i in range(3):
            a = csr[int(ip[3 * n_l + i]):int(ip[3 * n_l + i + 1])]
            b = gcsr[int(gip[3 * n_g + i]):int(gip[3 * n_g + i + 1])]
            assert a.numel() == b.numel()
            worst = max(worst, float((a - b).abs().max()))
    e_A = worst / float(gcsr.abs().max())
    ok = e_x < 1e-8 and e_r < 1e-12 and e_A < 1e-12 and e_ov < 1e-12 and e_pm < 1e-12 and n_to == 0 and abs(nrm - gnrm) < 1e-10 * gnrm and k > 0
    print(f'rank {rank}/{world}{" (generic partition, permuted node ids)" if generic else ""}: distributed BiCGStab {k} its (single GPU {gk}), err {err:.2e} | x {e_x:.2e}  res {e_r:.2e}  '
          f'CSR rows {e_A:.2e}  overlapped exchange {e_ov:.2e}  peer-memory exchange {e_pm:.2e} ({min(t_pm):.3f} ms, {n_to} timeouts)  ||res|| {nrm:.12e} vs {gnrm:.12e}  ->  {"OK" if ok else "FAIL"}', flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(int(flag.item() > 0))


if __name__ == '__main__':
    main()
