#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native JAX-CPFEM hot path.

    python bench.py --gpus N --steps K --warmup W [--n 200] [--impl reference]

Workload (BASELINE.json configs[4]): synthetic n^3 hex8 polycrystal (default 200^3 = 8 M cells,
64 M quadrature points, 8^3-cell grains with random orientations), 304-steel parameter set (FCC12, rate exponent
120), advanced 10 load steps from the virgin state with the state-update kernel so that every point flows
plastically; the timed "step" is load step 11: one state-update pass (update_int_vars_gp) + one Newton-iteration
assembly (newton_update: stress + consistent tangent at every point, hex8 integration, residual scatter, CSR
fill; multi-GPU: + interface exchange + residual-norm allreduce).  The mesh is element-partitioned into z-slabs
over the N GPUs (strong scaling: total work fixed).

One JSON line on rank 0.  `value` = quadrature-point updates / s of the update pass with inputs resident in HBM;
`assembly_ms` = ms per Newton-iteration assembly; `e2e` = the same update pass through the public API with HOST
(pinned) buffers: H2D of sol + state and D2H of the new state inside the timed region.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'jax-cpfem_b200'))

import numpy as np
import torch

# ---- work model (DESIGN.md "Work model") ------------------------------------------------------------------------
# CONTRACT figure: algorithmic flops per point of SURVEY.md section 8(d) / Appendix G (hand-derived 9x9 formulation,
# FMA = 2 flops).  Reported under roofline.contract only - the kernels run a leaner algorithm (6x6 symmetric
# crystal-frame form, active slip set, factored tangent), so this figure over-states the work they do.
F_UPDATE_FIXED = 1600.0     # set-up: Schmid rotation, B = F A M, hardening
F_ITER = 5000.0             # one local Newton iteration incl. one line-search residual evaluation
F_ASSEMBLY_FIXED = 16600.0  # set-up + consistent tangent (10 k) + element K_e share (5 k)
# EXECUTED work = what `roofline.achieved` / `roofline.frac` are computed from: FP64 thread instructions per point
# (DFMA + DMUL + DADD, each occupies one issue slot of the FP64 pipe) and DRAM bytes per point, measured by
# `ncu --set full` on the timed step at 128^3 (same workload, same local-Newton iteration count) and committed as
# profiles/fp64_instr.json (made by profiles/fp64_model.py from profiles/r2/*_ncu_summary_n128.txt).
B_UPDATE = 610.0            # algorithmic bytes/point: state in 336 + state out 264 + mesh/sol share 10
B_ASSEMBLY = 1940.0         # algorithmic bytes/point: state 240 + mesh/sol 10 + CSR zero-fill 244 + CSR RMW 244 (+244 read) + scratch 2 x 720


def _fp64_model():
    with open(os.path.join(ROOT, 'profiles', 'fp64_instr.json')) as f:
        return json.load(f)


MESH_N = 200
D_EPS, DT, PRE_STEPS = 2e-4, 2e-3, 10


def workload_name(n):
    return (f'synthetic {n}^3 hex8 polycrystal ({n ** 3} cells, {8 * n ** 3} quad points), 304 steel FCC12 exponent 120, '
            f'load step {PRE_STEPS + 1} (all points plastic), z-slab element partition')


def _peaks():
    p = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
    except Exception:
        pass
    return p


class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid) if not uuid.startswith('GPU-') else uuid)
            except Exception:
                vis = os.environ.get('CUDA_VISIBLE_DEVICES')
                phys = int(vis.split(',')[self.index]) if vis and vis.split(',')[self.index].isdigit() else self.index
                h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap',
                     0x80: 'hw_power_brake_slowdown'}
            while not self._stop.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.02)
        except Exception as e:          # sampling must never break the bench
            self.reasons.add('sampler_error:' + type(e).__name__)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)
        s = sorted(self.samples)
        return {'sm_mhz': (s[len(s) // 2] if s else None), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


# ----------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle (torch fp64 autodiff restatement of the reference, all host threads)
# on a bounded sample of the same workload
# ----------------------------------------------------------------------------------------------------------
def cpu_reference(n_mesh, sample_cells, steps, warmup):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import cpfem_oracle as O
    from cpfem_b200 import synthetic
    mat = O.steel304()
    # the sample = the first `sample_cells` cells of the n^3 mesh (x-fastest numbering), same grains / load history
    sx, sy = n_mesh + 1, (n_mesh + 1) ** 2
    c = np.arange(sample_cells, dtype=np.int64)
    i, j, k = c % n_mesh, (c // n_mesh) % n_mesh, c // (n_mesh * n_mesh)
    b = i + sx * j + sy * k
    cells_g = np.stack([b, b + 1, b + 1 + sx, b + sx, b + sy, b + 1 + sy, b + 1 + sx + sy, b + sx + sy], axis=1)
    node_gid = np.unique(cells_g)
    cells = np.searchsorted(node_gid, cells_g)
    gi, gj, gk = node_gid % sx, (node_gid // sx) % sx, node_gid // sy
    pts = np.stack([gi, gj, gk], axis=1) / float(n_mesh)
    G = (n_mesh + 7) // 8
    gid = (i // 8) + G * (j // 8) + G * G * (k // 8)
    quat = synthetic.grain_quaternions(G ** 3, 0)
    fe = O.FEOracle(pts, cells, O.make_uniform_batch_factory(mat))
    params = O.initial_internal_vars(sample_cells, mat, O.get_rot_mat(quat)[gid])
    noise = synthetic.noise_field(n_mesh)[node_gid]
    disp = lambda s: synthetic.affine_displacement(pts, D_EPS * s) + noise
    for s in range(1, PRE_STEPS + 1):
        params = fe.update_int_vars_gp(disp(s), params, DT)
    sol = disp(PRE_STEPS + 1)
    t_upd, t_asm, t_csr = [], [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        fe.update_int_vars_gp(sol, params, DT)
        t1 = time.perf_counter()
        res, V = fe.newton_update(sol, params, DT)
        tc = time.perf_counter()
        A = O.csr_from_coo(V, fe.I, fe.J, fe.nn * 3)       # the reference's own get_A line (solver.py:281): scipy, one core
        t2 = time.perf_counter()
        if it >= warmup:
            t_upd.append(t1 - t0)
            t_asm.append(t2 - t1)
            t_csr.append(t2 - tc)
    npts = sample_cells * 8
    return {'updates_per_s': npts * len(t_upd) / sum(t_upd), 'assembly_ms_sample': 1e3 * sum(t_asm) / len(t_asm),
            'assembly_us_per_cell': 1e6 * sum(t_asm) / len(t_asm) / sample_cells, 'sample_points': npts,
            'coo_to_csr_us_per_cell': 1e6 * sum(t_csr) / len(t_csr) / sample_cells,
            'ms_per_step': 1e3 * (sum(t_upd) + sum(t_asm)) / len(t_upd), 'cores': torch.get_num_threads()}


def _use_all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU arms run on rank 0 alone and use the box."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def same_algorithm_cpu(state, sol, rm, nthreads, sample_points=1 << 19):
    """Second, labelled CPU baseline: the algorithm the GPU kernels run (csrc/cp_point.cuh: crystal-frame 6x6 form,
    active slip set) compiled for the host (tests/hostcheck, g++ -O2 -fopenmp) on all host cores, on the first
    `sample_points` quadrature points of the benchmark state - so that GPU / CPU compares hardware on the SAME algorithm,
    while the oracle port above compares with the reference's own (autodiff) algorithm."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import hostcheck_build
    import cpfem_oracle as O
    lib = hostcheck_build.load()
    nthreads = hostcheck_build.set_threads(lib, nthreads)
    npts = int(min(sample_points, state[0].shape[0] * 8))
    ncell = npts // 8
    cells = rm.cells[:ncell]
    X = rm.points[cells]                                        # (c, 8, 3)
    g0, g1 = 0.21132486540518713, 0.7886751345948129
    nodes = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
    dN = np.zeros((8, 8, 3))
    for q in range(8):
        b = np.array([(q >> 2) & 1, (q >> 1) & 1, q & 1])
        f = np.where(nodes == b[None, :], g1, g0)
        sgn = np.where(nodes == 1, 1.0, -1.0)
        dN[q, :, 0] = sgn[:, 0] * f[:, 1] * f[:, 2]
        dN[q, :, 1] = sgn[:, 1] * f[:, 0] * f[:, 2]
        dN[q, :, 2] = sgn[:, 2] * f[:, 0] * f[:, 1]
    jac = np.einsum('cai,qaj->cqij', X, dN)
    grads = np.einsum('qaj,cqji->cqai', dN, np.linalg.inv(jac))
    u = sol.cpu().numpy()[cells]
    H = np.einsum('cai,cqaj->cqij', u, grads).reshape(npts, 3, 3)
    A, g, sl, R = (t[:ncell].reshape(npts, -1).cpu().numpy() for t in state[:4])
    mat = O.steel304()
    hostcheck_build.evaluate(lib, mat, DT, H[:4096], A[:4096], g[:4096], sl[:4096], R[:4096], tangent=False, pown=119)     # warm-up
    t0 = time.perf_counter()
    out = hostcheck_build.evaluate(lib, mat, DT, H, A, g, sl, R, tangent=False, pown=119)
    t1 = time.perf_counter()
    hostcheck_build.evaluate(lib, mat, DT, H, A, g, sl, R, tangent=True, pown=119)
    t2 = time.perf_counter()
    return {'value': npts / (t1 - t0), 'unit': 'quad-point updates/s', 'cores': nthreads,
            'kind': 'port (the GPU path\'s own per-point algorithm, csrc/cp_point.cuh built for the host with g++ -O2 -fopenmp)',
            'sample': f'first {npts} quadrature points of the benchmark state (stress + state update per point)',
            'with_tangent_points_per_s': npts / (t2 - t1), 'mean_local_newton_iters': float(out[-1][:, 0].mean())}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    _use_all_host_threads()
    sample_cells = args.cpu_sample_cells
    r = cpu_reference(args.n, sample_cells, args.steps, args.warmup)
    sample = (f'first {sample_cells} cells ({r["sample_points"]} quad points) of the {args.n}^3 polycrystal, 10 load steps '
              f'then the timed step, oracle port (torch fp64 + torch.func.jacfwd) on {r["cores"]} host threads')
    line = {'metric': 'cp_quad_point_updates_per_s', 'value': r['updates_per_s'], 'unit': 'quad-point updates/s',
            'impl': 'reference', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload_name(args.n), 'partition': 'host cores', 'state_layout': 'aos',
                       'sample_cells': sample_cells, 'l2': 'n/a (CPU arm)'},
            'assembly_ms': None, 'assembly_us_per_cell': r['assembly_us_per_cell'],
            'cpu_baseline': {'value': r['updates_per_s'], 'unit': 'quad-point updates/s', 'cores': r['cores'], 'kind': 'port',
                             'sample': sample, 'assembly_us_per_cell': r['assembly_us_per_cell'],
                             'coo_to_csr_us_per_cell': r['coo_to_csr_us_per_cell']},
            'e2e': {'value': r['updates_per_s'], 'unit': 'quad-point updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--n', type=int, default=MESH_N, help='mesh cells per edge (default 200 = BASELINE config)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--exchange', choices=['peer', 'nccl'], default='peer',
                    help='multi-GPU interface exchange: peer-memory mailboxes (CUDA IPC, cpfem_peer_put; default) or NCCL send/recv')
    ap.add_argument('--overlap-exchange', action='store_true',
                    help='multi-GPU: start the interface exchange after the first assembly chunk, beside the rest (measured: no gain, see DESIGN.md)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample-cells', type=int, default=1024)
    ap.add_argument('--layout', default='aos', choices=['aos', 'soa'])
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':      # keeps NCCL's version banner off stdout (one JSON line only)
        os.environ['NCCL_DEBUG'] = 'WARN'
    import torch.distributed as dist
    import cpfem_b200
    from cpfem_b200 import Plan, api, make_material, synthetic, slip_systems
    from cpfem_b200.partition import slab_partition_structured, ExchangePlan

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    N = args.n
    K, W = args.steps, max(args.warmup, 0)

    # ---- setup (untimed) ---------------------------------------------------------------------------------
    rm = slab_partition_structured(N, world, rank)
    plan = Plan(rm.cells, rm.points, slip_systems.FCC12)
    plan.set_active_cells(rm.n_owned_cells)
    nc = rm.n_owned_cells
    npts = nc * 8
    npts_global = 8 * N ** 3
    mat = make_material(2.622e5, 1.120e5, 0.746e5, 392.9772, 7295.1754, 8.0, 1.0 / 120.0, 1.0, 0.001, 1e-8, 8)
    G = (N + 7) // 8
    cg = torch.as_tensor(rm.cell_gid[:nc], device=dev)
    gid = (cg % N) // 8 + G * (((cg // N) % N) // 8) + G * G * ((cg // (N * N)) // 8)
    from cpfem_b200.problem import get_rot_mat
    Rg = torch.as_tensor(get_rot_mat(synthetic.grain_quaternions(G ** 3, 0)), device=dev)
    rot = Rg[gid][:, None].expand(nc, 8, 3, 3).contiguous()
    del gid, cg
    Fp = torch.eye(3, dtype=torch.float64, device=dev).expand(nc, 8, 3, 3).contiguous()
    g = torch.full((nc, 8, 12), 90.0, dtype=torch.float64, device=dev)
    sl = torch.zeros(nc, 8, 12, dtype=torch.float64, device=dev)
    noise = torch.as_tensor(synthetic.noise_field(N)[rm.node_gid], device=dev)
    pts_d = torch.as_tensor(rm.points, device=dev)
    scale = torch.tensor([-0.3, -0.3, 1.0], dtype=torch.float64, device=dev)
    disp = lambda s: (pts_d * scale * (D_EPS * s) + noise).contiguous()
    layout = api.LAYOUT_SOA if args.layout == 'soa' else api.LAYOUT_AOS
    if layout == api.LAYOUT_SOA:
        Fp, g, sl, rot = (api.aos_to_soa(t, c) for t, c in ((Fp, 9), (g, 12), (sl, 12), (rot, 9)))
    cur = [Fp, g, sl, rot]
    nxt = [torch.empty_like(Fp), torch.empty_like(g), torch.empty_like(sl)]
    # all-elastic step (load step 1 from the virgin state: one local Newton iteration per point) - reported separately,
    # the regime where the update pass is closest to its HBM floor
    def _t_elastic(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / 3
    sol1 = disp(1)
    t_el_upd = _t_elastic(lambda: plan.update_state(mat, sol1, cur, DT, out=nxt, layout=layout))
    del sol1
    for s in range(1, PRE_STEPS + 1):
        plan.update_state(mat, disp(s), cur, DT, out=nxt, layout=layout)
        cur, nxt = [nxt[0], nxt[1], nxt[2], rot], [cur[0], cur[1], cur[2]]
    sol = disp(PRE_STEPS + 1)
    state = cur
    out = nxt
    res = torch.empty(plan.nn, 3, dtype=torch.float64, device=dev)
    csr = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
    ex = None
    if world > 1:
        ip, ix = plan.csr_pattern()
        ex = ExchangePlan(rm, ip, ix)
        ex.prepare()
        if args.overlap_exchange:
            ex.attach(plan)              # interface exchange starts after the first assembly chunk, beside the rest
        elif args.exchange == 'peer':
            # peer-memory mailboxes (CUDA IPC + cpfem_peer_* kernels); if the box refuses IPC mappings every rank falls back
            # to the NCCL send/recv transport TOGETHER (attach_peer is collective) and the line says which one ran
            try:
                ex.attach_peer(with_csr=True)        # votes after every step that can fail: all ranks raise together
            except RuntimeError as e:
                print(f'rank {rank}: peer-memory exchange unavailable ({e}); using NCCL send/recv', file=sys.stderr)
                args.exchange = 'nccl'
    if ex is not None and args.overlap_exchange:
        exchange = lambda r, c: ex.exchange_overlapped(r, c)
    elif ex is not None and args.exchange == 'peer':
        exchange = lambda r, c: ex.exchange_peer(r, c)
    else:
        exchange = lambda r, c: ex.exchange(r, c)
    status_u = plan.new_status()
    status_a = plan.new_status()
    norm_buf = torch.zeros(1, dtype=torch.float64, device=dev)

    def do_update():
        plan.update_state(mat, sol, state, DT, out=out, status=status_u, layout=layout)

    def do_assembly():
        plan.newton_update(mat, sol, state, DT, res=res, csr_data=csr, status=status_a, layout=layout)
        if ex is not None:
            exchange(res, csr)
            s = ex.owned_sumsq(res, norm_buf)
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
        else:
            norm_buf.zero_()
            api.sumsq(res.reshape(-1), norm_buf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        do_update()
        do_assembly()
    status_u.zero_()
    status_a.zero_()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = int(cpfem_b200.lib().cpfem_launch_count())         # kernels launched by libcpfem_b200.so so far (this rank)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    for k in range(K):
        ev[k][0].record()
        do_update()
        ev[k][1].record()
        do_assembly()
        ev[k][2].record()
    launches = int(cpfem_b200.lib().cpfem_launch_count()) - launches0
    barrier()
    clocks = sampler.stop()
    t_upd = sum(e[0].elapsed_time(e[1]) for e in ev)
    t_asm = sum(e[1].elapsed_time(e[2]) for e in ev)
    t_tot = ev[0][0].elapsed_time(ev[-1][2])
    times = torch.tensor([t_upd, t_asm, t_tot], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_upd, t_asm, t_tot = times.tolist()
    st_u = status_u.clone()
    st_a = status_a.clone()
    if world > 1:
        dist.all_reduce(st_u[:2]); dist.all_reduce(st_a[:2])
        s3 = torch.stack([st_u[3], st_a[3]]); dist.all_reduce(s3)
        st_u[3], st_a[3] = s3[0], s3[1]
    tel = torch.tensor([t_el_upd], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tel, op=dist.ReduceOp.MAX)
    t_el_upd_max = float(tel[0])
    k_mean_u = float(st_u[3]) / (npts_global * K)
    k_mean_a = float(st_a[3]) / (npts_global * K)
    res_norm = float(torch.sqrt(norm_buf)[0])

    # ---- extra device-resident measurements (not part of `value`): fused update + average stress, F2 solver kernels ---------
    def _time(fn, reps):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    sig = torch.empty(nc, 3, 3, dtype=torch.float64, device=dev)
    t_fused = _time(lambda: plan.update_state_avg_stress(mat, sol, state, DT, out=out, sigma=sig, layout=layout), max(1, min(K, 3)))
    t_avg = _time(lambda: plan.avg_stress(mat, sol, state, DT, out=sig, layout=layout), max(1, min(K, 3)))
    solver_info = None
    if world == 1:
        xs = torch.randn(plan.ndof, dtype=torch.float64, device=dev)
        ys = torch.empty_like(xs)
        t_spmv = _time(lambda: plan.spmv(csr, xs, out=ys), 5)
        spmv_bytes = plan.nnz * 8 + (plan.nnz // 9) * 4 + (plan.nn + 1) * 8 + 2 * plan.ndof * 8
        nit = 20
        plan.bicgstab(csr, res.reshape(-1), tol=0.0, atol=0.0, maxiter=2)      # first use allocates the workspace, captures the graph
        torch.cuda.synchronize()
        tb0 = time.perf_counter()
        _, kit, _ = plan.bicgstab(csr, res.reshape(-1), tol=0.0, atol=0.0, maxiter=nit)
        torch.cuda.synchronize()
        tb1 = time.perf_counter()
        solver_info = {'spmv_ms': t_spmv, 'spmv_gbs': spmv_bytes / (t_spmv * 1e-3) / 1e9,
                       'spmv_bytes': spmv_bytes, 'bicgstab_ms_per_iteration': 1e3 * (tb1 - tb0) / nit, 'bicgstab_iterations_timed': nit,
                       'what': 'node-block SpMV on the assembled CSR (8 + 4/9 B per stored entry + vectors) and one Jacobi-BiCGStab '
                               'iteration (2 SpMV + fused vector kernels, host wall clock incl. the convergence polls); '
                               'row F2 of SURVEY 8(f): the reference does this through scipy -> BCOO on the host'}
        del xs, ys
    elif ex is not None:
        # row-partitioned Jacobi-BiCGStab over all ranks (halo exchange + all-reduced dot products; beyond the reference,
        # whose linear solver is single-device): wall clock of 20 iterations on the assembled, exchanged matrix
        from cpfem_b200.partition import HaloPlan, DistributedBicgstab
        halo = HaloPlan(rm, dev)
        dsolver = DistributedBicgstab(rm, halo)
        owned = torch.as_tensor(np.repeat(rm.owned_node_mask, 3), device=dev)
        minv = torch.where(owned, plan.csr_diagonal(csr, invert=True), torch.zeros(plan.ndof, dtype=torch.float64, device=dev))
        bvec = torch.where(owned, -res.reshape(-1), torch.zeros(plan.ndof, dtype=torch.float64, device=dev))
        nit = 20
        dsolver.solve(lambda v: plan.spmv(csr, v), bvec, minv=minv, tol=0.0, atol=0.0, maxiter=2)
        barrier()
        tb0 = time.perf_counter()
        dsolver.solve(lambda v: plan.spmv(csr, v), bvec, minv=minv, tol=0.0, atol=0.0, maxiter=nit)
        barrier()
        tb1 = time.perf_counter()
        solver_info = {'distributed_bicgstab_ms_per_iteration': 1e3 * (tb1 - tb0) / nit, 'bicgstab_iterations_timed': nit,
                       'what': 'row-partitioned Jacobi-BiCGStab: per iteration 2 node-block SpMVs on the local rows, 2 halo exchanges '
                               '(NCCL send/recv of interface and ghost-plane entries), 4 scalar all-reduces; host wall clock'}

    # ---- e2e: host (pinned) buffers through the public API ---------------------------------------------------
    # update pass: Plan.update_state_host streams the host-resident state through the device (H2D of sol + state, update,
    # D2H of the new state, chunk-pipelined on three streams).  Assembly: H2D of sol + state, newton_update, D2H of the
    # residual (the CSR stays on the device for the device linear solver).
    e2e = None
    if not args.no_e2e:
        if layout != api.LAYOUT_AOS:
            raise SystemExit('e2e uses the reference (AoS) layout')
        hsol = torch.empty(sol.shape, dtype=sol.dtype, pin_memory=True)
        hsol.copy_(sol)
        hstate = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in state]
        for h, t in zip(hstate, state):
            h.copy_(t)
        hnew = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in out]
        hres = torch.empty(res.shape, dtype=res.dtype, pin_memory=True)
        dsol = torch.empty_like(sol)
        h2d_upd = hsol.numel() * 8 + sum(h.numel() * 8 for h in hstate[:3])     # rot_mats stays resident (copied once, never changes)
        d2h_upd = sum(h.numel() * 8 for h in hnew)
        Ke = max(1, min(K, 3))
        te = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(Ke + 1)]
        barrier()
        for k in range(Ke + 1):                      # first pass is warm-up
            te[k][0].record()
            plan.update_state_host(mat, hsol, hstate, DT, out=hnew)            # synchronises before returning
            te[k][1].record()
            dsol.copy_(hsol, non_blocking=True)
            dst = [state[0], state[1], state[3]]
            for d, h in zip(dst, (hstate[0], hstate[1], hstate[3])):           # the assembly does not read the slips
                d.copy_(h, non_blocking=True)
            plan.newton_update(mat, dsol, state, DT, res=res, csr_data=csr, layout=layout)
            if ex is not None:
                exchange(res, csr)
            hres.copy_(res, non_blocking=True)
            te[k][2].record()
        barrier()
        t_u = sum(e[0].elapsed_time(e[1]) for e in te[1:])
        t_a = sum(e[1].elapsed_time(e[2]) for e in te[1:])
        tt = torch.tensor([t_u, t_a], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tb = torch.tensor([h2d_upd, d2h_upd], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb)
        e2e = {'value': npts_global * Ke / (tt[0].item() * 1e-3), 'unit': 'quad-point updates/s',
               'h2d_bytes_per_step': int(tb[0].item()), 'd2h_bytes_per_step': int(tb[1].item()),
               'ms_per_step': tt[0].item() / Ke, 'assembly_ms': tt[1].item() / Ke, 'steps': Ke,
               'what': 'Plan.update_state_host: pinned-host sol + state (Fp_inv, g, slip) in, new state out, chunk-pipelined H2D / '
                       'update / D2H; rot_mats (never modified, models_copper.py:282) is copied on the first call and kept on the '
                       'device, so the timed steps move 264 B/point each way; assembly_ms = H2D of sol + Fp_inv, g, rot, '
                       'newton_update, D2H of the residual (the CSR stays on the device for the device linear solver)'}
        del hsol, hstate, hnew, hres, dsol

    if ex is not None and getattr(ex, '_peer', None) is not None:
        peer_timeouts = ex.peer_timeouts()
        ex.detach_peer()                         # collective: unmap the neighbours' mailboxes, free this rank's
        assert peer_timeouts == 0, f'rank {rank}: {peer_timeouts} peer-memory waits timed out'
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline (update kernel = the metric's kernel; assembly reported beside it) --------------------------------
    # `achieved` / `frac` = EXECUTED FP64 work: FP64 thread instructions per point (ncu at the bench state, 128^3,
    # profiles/fp64_instr.json) x points of this rank / CUDA-event duration, as a share of the FP64 pipe's issue slots
    # (SMs x 64 lanes x clock; written as TFLOP/s with 2 flops per slot so that it compares with the DFMA peak).  `true_flops`
    # counts DFMA = 2, DMUL = DADD = 1.  `contract` is SURVEY 8(d)'s hand-derived 9x9 figure, which this algorithm undercuts.
    peaks = _peaks()
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    hbm_src = 'measured (MEASURED_PEAKS.json, burst copy)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
    fp64_meas = api.dfma_peak()
    # denominator: the larger of the DFMA microbenchmark of this run (which a power-capped moment can depress by 10 %) and
    # the nominal rate SMs x 64 DFMA/clk x max SM clock - the conservative choice for the fractions below
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    fp64_nominal = sm_count * 64 * 2 * (clocks.get('sm_max_mhz') or 1965) * 1e6 / 1e12
    fp64_peak = max(fp64_meas, fp64_nominal)
    upd_s = t_upd * 1e-3 / K
    asm_s = t_asm * 1e-3 / K
    pts_rank = npts_global / world
    model = _fp64_model()
    mk = model['kernels']
    ku, kp, ke = mk['k_update_state'], mk['k_point_tangent'], mk['k_element_tangent']
    tf = lambda f, t: pts_rank * f / t / 1e12
    peak_src = ('max(DFMA microbenchmark of this run (cpfem_dfma_peak_kernel): %.2f TFLOP/s, nominal SMs x 64 DFMA/clk x '
                'max SM clock: %.2f TFLOP/s); MEASURED_PEAKS.json has no FP64 entry' % (fp64_meas, fp64_nominal))
    model_src = ('%s: %s' % (model['source'], model['what']))
    f_upd = F_UPDATE_FIXED + F_ITER * k_mean_u
    f_asm = F_ASSEMBLY_FIXED + F_ITER * k_mean_a
    x_upd = 2.0 * ku['fp64_instr_per_point']
    roof = {'bound': 'fp64', 'kernel': 'k_update_state<12,119>', 'achieved': tf(x_upd, upd_s), 'peak': fp64_peak,
            'unit': 'TFLOP/s', 'frac': tf(x_upd, upd_s) / fp64_peak,
            'what': 'executed FP64 issue slots: (DFMA + DMUL + DADD thread instructions per point, ncu) x 2 x points / '
                    'CUDA-event duration of the kernel, against the DFMA peak = the share of the FP64 pipe the kernel keeps busy',
            'fp64_instr_per_point': ku['fp64_instr_per_point'], 'model_source': model_src,
            'ncu_pipe_fp64_cycles_active_pct_at_capture': ku['pipe_fp64_cycles_active_pct'],
            'true_flops': {'flops_per_point': ku['flops_per_point'], 'achieved': tf(ku['flops_per_point'], upd_s),
                           'frac': tf(ku['flops_per_point'], upd_s) / fp64_peak, 'what': 'DFMA = 2 flops, DMUL = DADD = 1'},
            'contract': {'flops_per_point': f_upd, 'achieved': tf(f_upd, upd_s), 'frac_of_peak': tf(f_upd, upd_s) / fp64_peak,
                         'model': 'SURVEY 8(d): 1.6 k + k x 5.0 k flops per point (hand-derived 9x9 form), k = mean local Newton '
                                  'iterations measured live; the kernel runs a leaner algorithm (6x6 symmetric crystal-frame '
                                  'form, active slip set), so this is NOT the work it does - a value above the peak only says so'},
            'traffic': ku['dram_bytes_per_point'] * pts_rank,
            'traffic_source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum of the same capture = %.0f B/point, scaled to '
                              'this launch; algorithmic bytes %.0f B/point' % (ku['dram_bytes_per_point'], B_UPDATE),
            'peak_source': peak_src, 'mean_local_newton_iters': k_mean_u,
            'hbm': {'achieved': pts_rank * B_UPDATE / upd_s / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': pts_rank * B_UPDATE / upd_s / 1e9 / hbm_peak, 'bytes_per_point': B_UPDATE, 'peak_source': hbm_src},
            'note': 'arithmetic intensity ~%.0f flop/B >> B200 balance (~5.5): the FP64 pipe is the bound, not HBM or tensor '
                    'cores (3x3 / 6x6 / 12-wide algebra per point)' % (ku['flops_per_point'] / B_UPDATE)}
    x_asm = 2.0 * (kp['fp64_instr_per_point'] + ke['fp64_instr_per_point'])
    fl_asm = kp['flops_per_point'] + ke['flops_per_point']
    share = lambda k: k['ms'] / (kp['ms'] + ke['ms'])
    roof_asm = {'bound': 'fp64', 'kernel': 'k_point_tangent<12,119> + k_element_tangent', 'achieved': tf(x_asm, asm_s),
                'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': tf(x_asm, asm_s) / fp64_peak,
                'what': 'executed FP64 issue slots of both kernels x 2 x points / duration of the whole assembly',
                'fp64_instr_per_point': {'k_point_tangent': kp['fp64_instr_per_point'], 'k_element_tangent': ke['fp64_instr_per_point']},
                'kernel_time_share_at_capture': {'k_point_tangent': share(kp), 'k_element_tangent': share(ke)},
                'true_flops': {'flops_per_point': fl_asm, 'achieved': tf(fl_asm, asm_s), 'frac': tf(fl_asm, asm_s) / fp64_peak},
                'contract': {'flops_per_point': f_asm, 'achieved': tf(f_asm, asm_s), 'frac_of_peak': tf(f_asm, asm_s) / fp64_peak},
                'traffic': (kp['dram_bytes_per_point'] + ke['dram_bytes_per_point']) * pts_rank,
                'mean_local_newton_iters': k_mean_a,
                'hbm': {'achieved': pts_rank * B_ASSEMBLY / asm_s / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                        'frac': pts_rank * B_ASSEMBLY / asm_s / 1e9 / hbm_peak, 'bytes_per_point': B_ASSEMBLY},
                'note': 'duration includes both kernels of every chunk, the residual memset and (multi-GPU) the interface exchange'}

    if solver_info is not None and 'spmv_gbs' in solver_info:
        solver_info['spmv_roofline'] = {'bound': 'hbm', 'achieved': solver_info['spmv_gbs'], 'peak': hbm_peak, 'unit': 'GB/s',
                                        'frac': solver_info['spmv_gbs'] / hbm_peak, 'peak_source': hbm_src}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        nthreads = _use_all_host_threads()
        r = cpu_reference(N, args.cpu_sample_cells, 2, 1)
        cpu = {'value': r['updates_per_s'], 'unit': 'quad-point updates/s', 'cores': r['cores'], 'kind': 'port',
               'sample': f'first {args.cpu_sample_cells} cells ({r["sample_points"]} points) of the same {N}^3 workload, oracle port '
                         f'(torch fp64 + torch.func.jacfwd, restatement of the reference\'s algorithm - JAX is not installable here)',
               'assembly_us_per_cell': r['assembly_us_per_cell'],
               'coo_to_csr_us_per_cell': r['coo_to_csr_us_per_cell'],
               'extrapolated_to_workload': {'update_s': npts_global / r['updates_per_s'],
                                            'assembly_s': r['assembly_us_per_cell'] * 1e-6 * (npts_global // 8),
                                            'note': 'per-point / per-cell sample rates x the full mesh (the reference cannot hold '
                                                    'it: ~110 GB of COO triplets at 200^3); assembly includes the scipy '
                                                    'COO->CSR step of solver.py:281 (coo_to_csr_us_per_cell, one core)'},
               'same_algorithm': same_algorithm_cpu(state, sol, rm, nthreads)}

    line = {
        'metric': 'cp_quad_point_updates_per_s', 'value': npts_global * K / (t_upd * 1e-3), 'unit': 'quad-point updates/s',
        'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': t_tot / K, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(N),
                   'partition': f'{world} slab(s)' + (', interface exchange overlapped with the assembly (starts after the first chunk)' if (world > 1 and args.overlap_exchange) else ''), 'state_layout': args.layout,
                   'exchange': (None if world == 1 else ('nccl send/recv (overlapped)' if args.overlap_exchange else
                                                         ('peer-memory mailbox (CUDA IPC, cpfem_peer_put)' if args.exchange == 'peer' else 'nccl send/recv'))),
                   'l2': 'inputs (>= 2 GB of state per pass) are larger than the 126 MB L2; no flush needed',
                   'layout_ab': 'SoA (component-major) state measured slower than the reference AoS layout on B200 (update 66.8 vs '
                                '61.1 ms at 200^3, profiles/r2/a_layout_ab.txt): 42 component streams 0.5 GB apart per warp '
                                'against one contiguous 2.3 kB record block; --layout soa reproduces it'},
        'update_ms': t_upd / K, 'assembly_ms': t_asm / K, 'assembly_metric': 'ms per Newton-iteration assembly '
        '(stress+tangent, hex8 integration, residual scatter, CSR fill, interface exchange, norm allreduce)',
        'update_avg_stress_fused_ms': t_fused, 'avg_stress_ms': t_avg, 'solver': solver_info,
        'elastic_step': {'update_ms': t_el_upd_max, 'updates_per_s': npts_global / (t_el_upd_max * 1e-3),
                         'hbm_gbs': npts_global / world * B_UPDATE / (t_el_upd_max * 1e-3) / 1e9,
                         'what': 'load step 1 from the virgin state (1 local Newton iteration per point): 610 B/point against the HBM peak'},
        'mean_local_newton_iters': k_mean_u, 'points_at_iter_cap': int(st_u[0]) + int(st_a[0]),
        'nonfinite_points': int(st_u[1]) + int(st_a[1]), 'residual_norm': res_norm,
        'roofline': roof, 'roofline_assembly': roof_asm, 'cpu_baseline': cpu, 'e2e': e2e,
        'gpu_launches': launches,
        'gpu_launches_what': 'kernels launched by libcpfem_b200.so inside the timed region on rank 0 (cpfem_launch_count(): '
                             'update 1 + assembly 2 per chunk + norm / exchange helpers per step; memsets and NCCL kernels not counted)',
        'clocks': clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
