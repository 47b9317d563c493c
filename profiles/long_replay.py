#!/usr/bin/env python
"""One-off long validation (not part of the test suite): ALL committed VTU steps of the three driver series of the
reference (50 DP-steel, 16 304-steel, 50 tantalum) replayed on the GPU path.  The committed fixtures under tests/golden
hold the first 8 steps; the full series are decoded into build/long/*.npz (git-ignored, travels with gpurun) by

    python profiles/long_replay.py --make        # in the build container, where /root/reference exists
    gpurun -- python profiles/long_replay.py     # on the B200 box

Result of the r1k code: profiles/r1/k_long_replay.txt."""
import sys
if '--make' in sys.argv:
    import importlib.util, os
    import numpy as np
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('mg', os.path.join(ROOT, 'tests', 'golden', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    REF = '/root/reference'
    os.makedirs(os.path.join(ROOT, 'build', 'long'), exist_ok=True)
    series = lambda d, pat, n: [mg.read_vtu(os.path.join(REF, d, pat % k)) for k in range(n)]
    v = series('polycrystal_DPsteel/data/vtk/polycrystal_DPsteel', 'u_inhomo_%03d.vtu', 50)
    np.savez_compressed(os.path.join(ROOT, 'build/long/dp.npz'), sol=np.stack([x['sol'] for x in v]),
                        sigma_zz=np.stack([x['sigma_zz'] for x in v]), sigma_xx=np.stack([x['sigma_xx'] for x in v]))
    v = series('polycrystal_304steel/data/vtk/polycrystal_304steel', 'u_%03d.vtu', 16)
    np.savez_compressed(os.path.join(ROOT, 'build/long/s304.npz'), sigma_zz=np.stack([x['sigma_zz'] for x in v]),
                        sigma_xx=np.stack([x['sigma_xx'] for x in v]))
    v = series('singlecrystal_tantalum/data/vtk/singlecrystal_tantalum', 'u_%03d.vtu', 50)
    np.savez_compressed(os.path.join(ROOT, 'build/long/ta.npz'), sigma_zz=np.stack([x['sigma_zz'] for x in v]))
    print('wrote build/long/{dp,s304,ta}.npz')
    sys.exit(0)

import os, sys, time, numpy as np, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('jax-cpfem_b200','oracle','tests'): sys.path.insert(0, os.path.join(ROOT,p))
from cpfem_b200.generate_mesh import Mesh
from cpfem_b200.solver import solver
GOLD=os.path.join(ROOT,'tests','golden'); LONG=os.path.join(ROOT,'build','long')
close=lambda a,b: np.isclose(a,b,atol=1e-5)
def cubic(C11,C12,C44):
    C=np.zeros((3,3,3,3))
    for i in range(3):
        C[i,i,i,i]=C11
        for j in range(3):
            if i!=j: C[i,i,j,j]=C12; C[i,j,i,j]=C44; C[i,j,j,i]=C44
    return C
def run(name, Model, g, long, quat, ori, extra, mk, disps, ts, opts, fix=None):
    pts, cells = g['points'], g['cells']
    problem = Model(Mesh(pts, cells), vec=3, dim=3, ele_type='HEX8', dirichlet_bc_info=mk(0.), additional_info=(quat, ori) + extra)
    params = list(problem.internal_vars)
    if fix: params = fix(params)
    sol = torch.zeros(len(pts), 3, dtype=torch.float64, device='cuda')
    n = long['sigma_zz'].shape[0]
    t0=time.time(); out=[]
    for i in range(n):
        problem.dt = ts[i+1]-ts[i]
        problem.fes[0].update_Dirichlet_boundary_conditions(mk(disps[i+1]))
        problem.set_params(params)
        sol = solver(problem, dict(opts, initial_guess=[sol]))[0]
        sg = problem.compute_avg_stress(sol, params).cpu().numpy()
        params = problem.update_int_vars_gp(sol, params)
        ref = long['sigma_zz'][i].astype(np.float64)
        e_mean = abs(sg[:,2,2].mean()/ref.mean()-1); e_cell = np.abs(sg[:,2,2]-ref).max()/np.abs(ref).max()
        e_sol = np.abs(sol.cpu().numpy()-long['sol'][i]).max()/np.abs(long['sol'][i]).max() if 'sol' in long else float('nan')
        out.append((e_mean,e_cell,e_sol))
    o=np.array(out)
    print(f'{name}: {n} steps in {time.time()-t0:.1f} s | mean sigma_zz: max rel err {o[:,0].max():.2e} | per-cell sigma_zz: max {o[:,1].max():.2e} | sol: max {np.nanmax(o[:,2]) if "sol" in long else float("nan"):.2e} | local Newton max its {int(problem.last_status[2])}')
    print('   per-step mean-stress error:', ' '.join('%.1e'%v for v in o[:,0]))
from cpfem_b200.models_DPsteel_inhomo import CrystalPlasticity as DP
from cpfem_b200.models_304steel import CrystalPlasticity as S304
from cpfem_b200.models_tantalum import CrystalPlasticity as Ta
g=np.load(os.path.join(GOLD,'dpsteel_vtu.npz')); L=g['points'].max(0)
mk=lambda d: [[lambda p: close(p[0],0.), lambda p: close(p[1],0.), lambda p: close(p[2],0.), lambda p: close(p[2],L[2])],[0,1,2,2],[lambda p:0.,lambda p:0.,lambda p:0.,lambda p:d]]
keys=np.stack([g['C11'],g['C12'],g['C44']],1); uniq,inv=np.unique(keys,axis=0,return_inverse=True)
Cgp=np.repeat(np.array([cubic(*u) for u in uniq])[inv.reshape(-1)][:,None],8,axis=1)
def fix(params): params[9]=torch.as_tensor(Cgp,device='cuda'); return params
run('DP steel 10^3 (polycrystal_DPsteel_inhomo.py)', DP, g, np.load(os.path.join(LONG,'dp.npz')), np.loadtxt(os.path.join(GOLD,'quat_dp.txt'))[:20,1:], g['cell_ori_inds'].astype(int), (g['phase_inds'].astype(int),), mk, np.linspace(0.,0.01*L[0],51), np.linspace(0.,10.,51), {'jax_solver':{}, 'line_search_flag':True}, fix)
g=np.load(os.path.join(GOLD,'steel304_vtu.npz')); L=g['points'].max(0)
corner=lambda p: close(p[0],0.)&close(p[1],0.)&close(p[2],L[2])
mk=lambda d: [[corner,corner,lambda p: close(p[2],0.),lambda p: close(p[2],L[2])],[0,1,2,2],[lambda p:0.,lambda p:0.,lambda p:0.,lambda p:d]]
run('304 steel 16^3 (polycrystal_304steel.py)', S304, g, np.load(os.path.join(LONG,'s304.npz')), np.loadtxt(os.path.join(GOLD,'quat_304.txt'))[:8,1:], g['cell_ori_inds'].astype(int), (), mk, np.linspace(0.,0.01*L[0],51), np.linspace(0.,0.1,51), {'jax_solver':{}})
g=np.load(os.path.join(GOLD,'tantalum_vtu.npz')); L=g['points'].max(0)
corner=lambda p: close(p[0],0.)&close(p[1],0.)&close(p[2],L[2])
mk=lambda d: [[corner,corner,lambda p: close(p[2],0.),lambda p: close(p[2],L[2])],[0,1,2,2],[lambda p:0.,lambda p:0.,lambda p:0.,lambda p:d]]
run('Ta 10^3 (singlecrystal_tantalum.py)', Ta, g, np.load(os.path.join(LONG,'ta.npz')), np.array([[1.,0,0,0]]), np.zeros(len(g['cells']),int), (), mk, np.linspace(0.,-0.0125*L[0],51), np.linspace(0.,12.5,51), {'jax_solver':{}})
