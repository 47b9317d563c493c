"""Thin Python layer over the C ABI: device memory and streams come from torch, the work is done by the
sm_100a kernels in libcpfem_b200.so.  Arrays are torch CUDA tensors (float64 / int32 / int64).
"""
from __future__ import annotations

import ctypes
import weakref
from typing import List, Optional, Sequence

import numpy as onp
import torch

from . import _lib
from ._lib import Material, State, StateGrad, StateOut, check

LAYOUT_AOS, LAYOUT_SOA = 0, 1


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f64(x, device):
    """float64 contiguous tensor on `device` (copies host arrays; no-op for resident tensors)."""
    if isinstance(x, torch.Tensor):
        if x.dtype != torch.float64 or not x.is_contiguous() or x.device != device:
            x = x.to(device=device, dtype=torch.float64).contiguous()
        return x
    return torch.as_tensor(onp.ascontiguousarray(x, dtype=onp.float64)).to(device)


def make_material(C11, C12, C44, h, t_sat, gss_a, xm, r=1.0, ao=0.001, tol=1e-8, max_sub_step=5, max_iter=200) -> Material:
    return Material(C11, C12, C44, h, t_sat, gss_a, ao, xm, r, tol, int(max_sub_step), int(max_iter))


class Plan:
    """Per-mesh plan (cpfem_plan): device connectivity + coordinates, scipy-identical CSR pattern, slot map."""

    def __init__(self, cells, points, slip, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('cpfem_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        L = _lib.lib()
        with torch.cuda.device(self.device):
            cells_t = torch.as_tensor(onp.ascontiguousarray(cells, dtype=onp.int32)) if not isinstance(cells, torch.Tensor) else cells
            self.cells = cells_t.to(device=self.device, dtype=torch.int32).contiguous()
            self.points = _dev_f64(points, self.device)
            slip = onp.ascontiguousarray(slip, dtype=onp.float64)
            assert slip.ndim == 2 and slip.shape[1] == 6
            self.slip = slip
            self.ns = slip.shape[0]
            self.nc = int(self.cells.shape[0])
            self.nn = int(self.points.shape[0])
            h = ctypes.c_void_p()
            check(L.cpfem_plan_create(_ptr(self.cells), self.nc, _ptr(self.points), self.nn,
                                      slip.ctypes.data_as(ctypes.c_void_p), self.ns, _stream(), ctypes.byref(h)),
                  'cpfem_plan_create')
            self._h = h
            info = (ctypes.c_int64 * 6)()
            check(L.cpfem_plan_info(self._h, info), 'cpfem_plan_info')
            self.nnz = int(info[3])
            self.max_valence = int(info[4])
            self.chunk_cells = int(info[5])
        self.ndof = 3 * self.nn
        self.nc_active = self.nc
        self.np = 8 * self.nc
        self._indptr = None
        self._indices = None

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                _lib.lib().cpfem_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_active_cells(self, n_active):
        """Element-partitioned runs: kernels loop over the first n_active (owned) cells; ghost cells keep their slots."""
        check(_lib.lib().cpfem_plan_set_active_cells(self._h, int(n_active)), 'cpfem_plan_set_active_cells')
        self.nc_active = int(n_active)
        self.np = 8 * self.nc_active

    def set_progress_event(self, cell_prefix, event: Optional[torch.cuda.Event]):
        """cpfem_plan_set_progress_event: newton_update records `event` once cells [0, cell_prefix) are assembled
        (None switches it off).  The event must have been recorded once already (torch creates the handle lazily)."""
        handle = None if event is None else ctypes.c_void_p(event.cuda_event)
        check(_lib.lib().cpfem_plan_set_progress_event(self._h, int(cell_prefix), handle), 'cpfem_plan_set_progress_event')
        self._progress_event = event                     # keep it alive as long as the plan may record it

    # ---- CSR pattern -------------------------------------------------------------------------
    def csr_pattern(self):
        """(indptr int64 (ndof+1), indices int32 (nnz)) as device tensors (copied once from the plan)."""
        if self._indptr is None:
            with torch.cuda.device(self.device):
                ip = torch.empty(self.ndof + 1, dtype=torch.int64, device=self.device)
                ix = torch.empty(self.nnz, dtype=torch.int32, device=self.device)
                check(_lib.lib().cpfem_plan_csr_copy(self._h, _ptr(ip), _ptr(ix), _stream()), 'cpfem_plan_csr_copy')
            self._indptr, self._indices = ip, ix
        return self._indptr, self._indices

    # ---- helpers -----------------------------------------------------------------------------
    def _uniform_value(self, t: torch.Tensor):
        """The value of a per-point parameter array if it is the same at every point, else None.  One reduction + host
        read per distinct tensor (identity + version counter), remembered afterwards."""
        # keyed on the tensor OBJECT (weak reference) + its version counter: an address alone can be recycled by the
        # allocator for a new array with other values
        key = (id(t), t._version)
        cache = self.__dict__.setdefault('_uniform_cache', {})
        hit = cache.get(key)
        if hit is not None and hit[0]() is t:
            return hit[1]
        if len(cache) > 64:
            cache.clear()
        lo, hi = torch.aminmax(t)
        val = float(lo) if float(lo) == float(hi) else None
        cache[key] = (weakref.ref(t), val)
        return val

    def _state(self, params: Sequence[torch.Tensor], layout=LAYOUT_AOS, mat: Optional[Material] = None, validate=True):
        """cpfem_state from the reference's internal_vars list (4, 9 or 10 arrays).  With `mat` given, per-point parameter
        arrays that hold one value everywhere (the calibration drivers scale whole arrays, calibration_case4_...py:238-251;
        DP steel shares the rate sensitivity between its phases) are demoted to scalars of a copy of the material, so the
        kernels take their uniform-parameter / compile-time-exponent paths; returns (state, tensors, material)."""
        n = len(params)
        if n not in (4, 9, 10):
            raise ValueError('internal_vars must have 4 (uniform), 9 (calibration) or 10 (DP steel) arrays')
        ts = [_dev_f64(p, self.device) for p in params]
        st = State()
        st.Fp_inv, st.g, st.slip, st.rot = (t.data_ptr() for t in ts[:4])
        m = mat
        if n >= 9:
            names = ('gss_a', 'h', 't_sat', 'xm', 'r')
            vals = [self._uniform_value(t) for t in ts[4:9]] if mat is not None else [None] * 5
            if any(v is not None for v in vals):
                m = Material.from_buffer_copy(mat)
            for name, t, v in zip(names, ts[4:9], vals):
                if v is None:
                    setattr(st, name, t.data_ptr())
                else:
                    setattr(m, name, v)
        if n == 10:
            if validate:
                self._check_cubic(ts[9])
            st.C = ts[9].data_ptr()
        st.layout = layout
        return (st, ts) if mat is None else (st, ts, m)

    def _check_cubic(self, C: torch.Tensor, rtol=1e-12):
        """The kernels read three entries of every point's (3,3,3,3) elastic tensor and assume the cubic pattern in the
        crystal frame (what the reference builds, models_DPsteel_inhomo.py:121-147).  Anything else - a pre-rotated or a
        lower-symmetry tensor - would give wrong answers silently, so it is refused: one reduction per distinct tensor
        (identity + version counter), remembered afterwards."""
        key = (id(C), C._version)
        cache = self.__dict__.setdefault('_cubic_cache', {})
        hit = cache.get(key)
        if hit is not None and hit() is C:
            return
        if C.numel() % 81 != 0:
            raise ValueError('C_gp must have shape (..., 3, 3, 3, 3)')
        bad = torch.zeros(1, dtype=torch.int64, device=self.device)
        check(_lib.lib().cpfem_check_cubic(_ptr(C), int(C.numel() // 81), float(rtol), _ptr(bad), _stream()), 'cpfem_check_cubic')
        nbad = int(bad.item())
        if nbad:
            raise ValueError(f'C_gp: {nbad} of {C.numel() // 81} points hold an elastic tensor that is not cubic in the crystal '
                             'frame (C11 on iiii, C12 on iijj, C44 on ijij/ijji, zero elsewhere); the crystal-plasticity kernels '
                             'support the cubic form of the reference only (models_DPsteel_inhomo.py:121-147)')
        if len(cache) > 16:
            cache.clear()
        cache[key] = weakref.ref(C)

    def new_status(self):
        return torch.zeros(4, dtype=torch.int64, device=self.device)

    # ---- hot-path calls ----------------------------------------------------------------------
    def update_state(self, mat: Material, sol, params, dt, out=None, status=None, layout=LAYOUT_AOS):
        """cpfem_update_state.  Returns (Fp_inv_new, g_new, slip_new) with the shapes of the inputs."""
        with torch.cuda.device(self.device):
            st, ts, mat = self._state(params, layout, mat)
            sol = _dev_f64(sol, self.device)
            if out is None:
                out = [torch.empty_like(ts[0]), torch.empty_like(ts[1]), torch.empty_like(ts[2])]
            so = StateOut(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), layout)
            check(_lib.lib().cpfem_update_state(self._h, ctypes.byref(mat), _ptr(sol), ctypes.byref(st), ctypes.byref(so),
                                                float(dt), _ptr(status), _stream()), 'cpfem_update_state')
        return out

    def update_state_avg_stress(self, mat: Material, sol, params, dt, out=None, sigma=None, status=None, layout=LAYOUT_AOS):
        """cpfem_update_state_avg_stress: update_int_vars_gp and compute_avg_stress from ONE local solve per point.
        Returns ((Fp_inv_new, g_new, slip_new), sigma_cell (nc, 3, 3))."""
        with torch.cuda.device(self.device):
            st, ts, mat = self._state(params, layout, mat)
            sol = _dev_f64(sol, self.device)
            if out is None:
                out = [torch.empty_like(ts[0]), torch.empty_like(ts[1]), torch.empty_like(ts[2])]
            if sigma is None:
                sigma = torch.empty(self.nc_active, 3, 3, dtype=torch.float64, device=self.device)
            so = StateOut(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), layout)
            check(_lib.lib().cpfem_update_state_avg_stress(self._h, ctypes.byref(mat), _ptr(sol), ctypes.byref(st),
                                                           ctypes.byref(so), float(dt), _ptr(sigma), _ptr(status), _stream()),
                  'cpfem_update_state_avg_stress')
        return out, sigma

    def update_state_cells(self, mat: Material, sol, params, dt, cell0, ncells, out, status=None, layout=LAYOUT_AOS,
                           demote=True):
        """cpfem_update_state_cells: `params` / `out` hold only the points of cells [cell0, cell0+ncells).  `demote`:
        per-point parameter arrays of the chunk that hold one value everywhere become scalars of the material (as in
        update_state); update_state_host resolves that once for the whole state and passes demote=False."""
        if demote:
            st, ts, mat = self._state(params, layout, mat)
        else:
            st, ts = self._state(params, layout)
        so = StateOut(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), layout)
        check(_lib.lib().cpfem_update_state_cells(self._h, ctypes.byref(mat), _ptr(sol), ctypes.byref(st), ctypes.byref(so),
                                                  float(dt), int(cell0), int(ncells), _ptr(status), _stream()),
              'cpfem_update_state_cells')
        return out

    def update_state_host(self, mat: Material, sol, params, dt, out=None, status=None, chunk_cells=None, cache_rot=True):
        """update_int_vars_gp for a HOST-resident state (reference layout, torch CPU tensors; pinned memory makes the
        copies asynchronous): the state streams through the device in chunks of `chunk_cells` cells on three streams,
        so that the H2D copy of chunk k+1, the update of chunk k and the D2H copy of chunk k-1 overlap (PCIe is full
        duplex).  `out` = [Fp_inv_new, g_new, slip_new] host tensors (allocated pinned when None).  Returns `out` after
        synchronising; the per-point material arrays of the DP-steel form (params[4:]) are streamed with the state.
        `cache_rot`: rot_mats_gp never changes during a simulation (models_copper.py:282 passes it through), so a device
        copy is kept between calls - keyed on the host tensor's identity and version counter - and only the 264 B/point
        that do change cross the bus in each direction."""
        hs = [p if isinstance(p, torch.Tensor) else torch.as_tensor(onp.ascontiguousarray(p, dtype=onp.float64)) for p in params]
        nc = self.nc_active
        if any(h.is_cuda or h.dtype != torch.float64 or not h.is_contiguous() or h.shape[0] != nc for h in hs):
            raise ValueError('update_state_host: params must be contiguous float64 CPU tensors with leading dimension nc')
        if out is None:
            out = [torch.empty(h.shape, dtype=torch.float64, pin_memory=nc > 0) for h in hs[:3]]
        if nc == 0:                                     # a rank that owns no cells
            return out
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            s_in, s_out = self._host_streams()
            sol_d = _dev_f64(sol if isinstance(sol, torch.Tensor) else onp.asarray(sol), self.device)
            if chunk_cells is None:                     # at least 8 chunks in flight, at most 2 Mi points each
                chunk_cells = max(1, min(1 << 18, -(-nc // 8)))
            cc = int(min(chunk_cells, nc))
            nbuf = 2
            key = (cc, tuple(tuple(h.shape[1:]) for h in hs))
            if getattr(self, '_host_key', None) != key:
                self._host_in = [[torch.empty((cc,) + tuple(h.shape[1:]), dtype=torch.float64, device=self.device) for h in hs]
                                 for _ in range(nbuf)]
                self._host_out = [[torch.empty((cc,) + tuple(h.shape[1:]), dtype=torch.float64, device=self.device) for h in hs[:3]]
                                  for _ in range(nbuf)]
                self._host_key = key
            # per-point parameter arrays (DP / calibration forms) that hold one value everywhere: demote them to scalars
            # of the material ONCE, on the host tensors, and stream only the arrays that really vary - every chunk then
            # runs the same (uniform-parameter / compile-time exponent) kernels as the device-resident path
            skip = set()
            if len(hs) >= 9:
                names = ('gss_a', 'h', 't_sat', 'xm', 'r')
                vals = [self._uniform_value(h) for h in hs[4:9]]
                if any(v is not None for v in vals):
                    mat = Material.from_buffer_copy(mat)
                    for j, (name, v) in enumerate(zip(names, vals)):
                        if v is not None:
                            setattr(mat, name, v)
                            skip.add(4 + j)
            rot_dev = None
            if cache_rot:
                rk = getattr(self, '_rot_key', None)             # (weak reference to the host tensor, its version counter)
                if rk is None or rk[0]() is not hs[3] or rk[1] != hs[3]._version:
                    self._rot_dev = hs[3].to(self.device, non_blocking=True)
                    self._rot_key = (weakref.ref(hs[3]), hs[3]._version)
                rot_dev = self._rot_dev
            self._host_bad = torch.zeros(1, dtype=torch.int64, device=self.device)
            ev_in = [torch.cuda.Event() for _ in range(nbuf)]
            ev_cmp = [torch.cuda.Event() for _ in range(nbuf)]
            ev_out = [torch.cuda.Event() for _ in range(nbuf)]
            start = torch.cuda.Event()
            start.record(cur)
            s_in.wait_event(start)
            s_out.wait_event(start)
            k = 0
            for c0 in range(0, nc, cc):
                n = min(cc, nc - c0)
                b = k % nbuf
                din = [t[:n] for t in self._host_in[b]]
                dout = [t[:n] for t in self._host_out[b]]
                with torch.cuda.stream(s_in):
                    if k >= nbuf:
                        s_in.wait_event(ev_cmp[b])            # the update that read this input buffer is done
                    for j, (d, h) in enumerate(zip(din, hs)):
                        if (j == 3 and rot_dev is not None) or j in skip:
                            continue                          # resident on the device already / demoted to a scalar
                        d.copy_(h[c0:c0 + n], non_blocking=True)
                    ev_in[b].record(s_in)
                if rot_dev is not None:
                    din = din[:3] + [rot_dev[c0:c0 + n]] + din[4:]
                cur.wait_event(ev_in[b])
                if k >= nbuf:
                    cur.wait_event(ev_out[b])                 # the D2H copy that read this output buffer is done
                self._update_cells_demoted(mat, sol_d, din, dt, c0, n, dout, status, skip)
                ev_cmp[b].record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[b])
                    for h, d in zip(out, dout):
                        h[c0:c0 + n].copy_(d, non_blocking=True)
                    ev_out[b].record(s_out)
                k += 1
            for b in range(min(k, nbuf)):
                cur.wait_event(ev_out[b])
            cur.synchronize()
            if len(hs) == 10 and int(self._host_bad.item()):
                raise ValueError(f'C_gp: {int(self._host_bad.item())} points hold an elastic tensor that is not cubic in the '
                                 'crystal frame (see Plan._check_cubic); the returned state is invalid')
        return out

    def _update_cells_demoted(self, mat, sol_d, din, dt, c0, n, dout, status, skip):
        """One chunk of update_state_host: the arrays in `skip` were demoted to scalars of `mat` (NULL pointers)."""
        st, ts = self._state(din, LAYOUT_AOS, validate=False)
        if len(din) == 10:          # C_gp of this chunk: violations are counted on the device, read once at the end of the pass
            check(_lib.lib().cpfem_check_cubic(_ptr(ts[9]), int(n) * 8, 1e-12, _ptr(self._host_bad), _stream()), 'cpfem_check_cubic')
        for j, name in enumerate(('gss_a', 'h', 't_sat', 'xm', 'r')):
            if 4 + j in skip:
                setattr(st, name, None)
        so = StateOut(dout[0].data_ptr(), dout[1].data_ptr(), dout[2].data_ptr(), LAYOUT_AOS)
        check(_lib.lib().cpfem_update_state_cells(self._h, ctypes.byref(mat), _ptr(sol_d), ctypes.byref(st), ctypes.byref(so),
                                                  float(dt), int(c0), int(n), _ptr(status), _stream()),
              'cpfem_update_state_cells')

    def _host_streams(self):
        if getattr(self, '_s_in', None) is None:
            self._s_in = torch.cuda.Stream(device=self.device)
            self._s_out = torch.cuda.Stream(device=self.device)
        return self._s_in, self._s_out

    def residual(self, mat: Material, sol, params, dt, res=None, status=None, layout=LAYOUT_AOS):
        with torch.cuda.device(self.device):
            st, ts, mat = self._state(params, layout, mat)
            sol = _dev_f64(sol, self.device)
            if res is None:
                res = torch.empty(self.nn, 3, dtype=torch.float64, device=self.device)
            check(_lib.lib().cpfem_residual(self._h, ctypes.byref(mat), _ptr(sol), ctypes.byref(st), float(dt), _ptr(res),
                                            _ptr(status), _stream()), 'cpfem_residual')
        return res

    def newton_update(self, mat: Material, sol, params, dt, res=None, csr_data=None, coo_V=None, want_csr=True,
                      want_V=False, status=None, layout=LAYOUT_AOS):
        with torch.cuda.device(self.device):
            st, ts, mat = self._state(params, layout, mat)
            sol = _dev_f64(sol, self.device)
            if res is None:
                res = torch.empty(self.nn, 3, dtype=torch.float64, device=self.device)
            if csr_data is None and want_csr:
                csr_data = torch.empty(self.nnz, dtype=torch.float64, device=self.device)
            if coo_V is None and want_V:
                coo_V = torch.empty(self.nc_active * 576, dtype=torch.float64, device=self.device)
            check(_lib.lib().cpfem_newton_update(self._h, ctypes.byref(mat), _ptr(sol), ctypes.byref(st), float(dt),
                                                 _ptr(res), _ptr(csr_data), _ptr(coo_V), _ptr(status), _stream()),
                  'cpfem_newton_update')
        return res, csr_data, coo_V

    def avg_stress(self, mat: Material, sol, params, dt, out=None, status=None, layout=LAYOUT_AOS):
        with torch.cuda.device(self.device):
            st, ts, mat = self._state(params, layout, mat)
            sol = _dev_f64(sol, self.device)
            if out is None:
                out = torch.empty(self.nc_active, 3, 3, dtype=torch.float64, device=self.device)
            check(_lib.lib().cpfem_avg_stress(self._h, ctypes.byref(mat), _ptr(sol), ctypes.byref(st), float(dt), _ptr(out),
                                              _ptr(status), _stream()), 'cpfem_avg_stress')
        return out

    def point_stress_tangent(self, mat: Material, u_grads, params, dt, want_tangent=True, status=None):
        """tensor_map (and its jacfwd) on explicit u_grads (np, 3, 3); state arrays have leading size np."""
        with torch.cuda.device(self.device):
            st, ts, mat = self._state(params, LAYOUT_AOS, mat)
            ug = _dev_f64(u_grads, self.device)
            n = int(ug.numel() // 9)
            P = torch.empty(n, 3, 3, dtype=torch.float64, device=self.device)
            A = torch.empty(n, 3, 3, 3, 3, dtype=torch.float64, device=self.device) if want_tangent else None
            check(_lib.lib().cpfem_point_stress_tangent(self._h, ctypes.byref(mat), _ptr(ug), n, ctypes.byref(st), float(dt),
                                                        _ptr(P), _ptr(A), _ptr(status), _stream()),
                  'cpfem_point_stress_tangent')
        return P, A

    def point_update_state(self, mat: Material, u_grads, params, dt, status=None):
        """update_int_vars_map under vmap (models_copper.py:164-169,267-269): explicit u_grads (np, 3, 3) + state arrays
        of np points -> (Fp_inv_new, g_new, slip_new)."""
        with torch.cuda.device(self.device):
            st, ts, mat = self._state(params, LAYOUT_AOS, mat)
            ug = _dev_f64(u_grads, self.device)
            n = int(ug.numel() // 9)
            out = [torch.empty_like(ts[0]), torch.empty_like(ts[1]), torch.empty_like(ts[2])]
            so = StateOut(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), LAYOUT_AOS)
            check(_lib.lib().cpfem_point_update_state(self._h, ctypes.byref(mat), _ptr(ug), n, ctypes.byref(st), ctypes.byref(so),
                                                      float(dt), _ptr(status), _stream()), 'cpfem_point_update_state')
        return out

    def point_eval(self, mat: Material, u_grads, params, dt, want_tangent=True, want_state=True, status=None):
        """cpfem_point_eval: stress, tangent, new state and the per-point solve account (iterations, residual
        evaluations, status bits) from one local solve per point.  Returns (P, A or None, new_state or None, info (np, 3))."""
        with torch.cuda.device(self.device):
            st, ts, mat = self._state(params, LAYOUT_AOS, mat)
            ug = _dev_f64(u_grads, self.device)
            n = int(ug.numel() // 9)
            P = torch.empty(n, 3, 3, dtype=torch.float64, device=self.device)
            A = torch.empty(n, 3, 3, 3, 3, dtype=torch.float64, device=self.device) if want_tangent else None
            out = [torch.empty_like(ts[0]), torch.empty_like(ts[1]), torch.empty_like(ts[2])] if want_state else None
            so = StateOut(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), LAYOUT_AOS) if want_state else None
            info = torch.zeros(n, 3, dtype=torch.int32, device=self.device)
            check(_lib.lib().cpfem_point_eval(self._h, ctypes.byref(mat), _ptr(ug), n, ctypes.byref(st), float(dt), _ptr(P), _ptr(A),
                                              ctypes.byref(so) if so is not None else None, _ptr(info), _ptr(status), _stream()),
                  'cpfem_point_eval')
        return P, A, out, info

    # ---- adjoint columns (F5) ----------------------------------------------------------------------
    @staticmethod
    def _nextra(params):
        return {4: 0, 9: 5, 10: 6}[len(params)]

    def point_jac_x(self, mat: Material, u_grads, params, dt, want_jac_y=True, status=None):
        """f_jvp's jac_x (np, 9, nx), jac_y (np, 9, 9) and y = S (np, 9) at the converged local solution
        (models_copper.py:251-259); x in the reference's ravel order, see include/cpfem.h."""
        with torch.cuda.device(self.device):
            st, ts = self._state(params, LAYOUT_AOS)
            ug = _dev_f64(u_grads, self.device)
            n, ne = int(ug.numel() // 9), self._nextra(params)
            nx = 27 + 2 * self.ns + (5 if ne >= 5 else 0) + (81 if ne >= 6 else 0)
            jx = torch.empty(n, 9, nx, dtype=torch.float64, device=self.device)
            jy = torch.empty(n, 9, 9, dtype=torch.float64, device=self.device) if want_jac_y else None
            S = torch.empty(n, 9, dtype=torch.float64, device=self.device)
            check(_lib.lib().cpfem_point_jac_x(self._h, ctypes.byref(mat), _ptr(ug), n, ctypes.byref(st), float(dt), ne, _ptr(jx),
                                               _ptr(jy), _ptr(S), _ptr(status), _stream()), 'cpfem_point_jac_x')
        return jx, jy, S

    def point_vjp(self, mat: Material, u_grads, params, dt, W, status=None):
        """W : d tensor_map / dx through the local solve: (np, nx)."""
        with torch.cuda.device(self.device):
            st, ts = self._state(params, LAYOUT_AOS)
            ug, W = _dev_f64(u_grads, self.device), _dev_f64(W, self.device)
            n, ne = int(ug.numel() // 9), self._nextra(params)
            nx = 27 + 2 * self.ns + (5 if ne >= 5 else 0) + (81 if ne >= 6 else 0)
            grad = torch.empty(n, nx, dtype=torch.float64, device=self.device)
            check(_lib.lib().cpfem_point_vjp(self._h, ctypes.byref(mat), _ptr(ug), n, ctypes.byref(st), float(dt), ne, _ptr(W), _ptr(grad),
                                             _ptr(status), _stream()), 'cpfem_point_vjp')
        return grad

    def vjp_params(self, mat: Material, sol, params, dt, adjoint, status=None):
        """vjp_linear_fn of implicit_vjp (solver.py:832-848): adjoint (nnodes, 3) . d(residual)/d(internal_vars) as a list of
        arrays shaped like `params` (zero the adjoint on the Dirichlet dofs first; the caller applies the final minus sign)."""
        with torch.cuda.device(self.device):
            st, ts = self._state(params, LAYOUT_AOS)
            sol, adj = _dev_f64(sol, self.device), _dev_f64(adjoint, self.device)
            out = [torch.empty_like(t) for t in ts]
            sg = StateGrad(*[o.data_ptr() for o in out], *([None] * (10 - len(out))))
            check(_lib.lib().cpfem_vjp_params(self._h, ctypes.byref(mat), _ptr(sol), ctypes.byref(st), float(dt), _ptr(adj),
                                              ctypes.byref(sg), _ptr(status), _stream()), 'cpfem_vjp_params')
        return out

    def csr_transpose(self, csr_data, out=None):
        """Values of A^T on the plan's (structurally symmetric) pattern."""
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty_like(csr_data)
            check(_lib.lib().cpfem_csr_transpose(self._h, _ptr(csr_data), _ptr(out), _stream()), 'cpfem_csr_transpose')
        return out

    def apply_dirichlet(self, rows, vals, sol, res=None, csr_data=None):
        with torch.cuda.device(self.device):
            check(_lib.lib().cpfem_apply_dirichlet(self._h, _ptr(rows), _ptr(vals), int(rows.numel()), _ptr(sol), _ptr(res),
                                                   _ptr(csr_data), _stream()), 'cpfem_apply_dirichlet')


    # ---- device linear solver (F2) -----------------------------------------------------------------
    def spmv(self, csr_data, x, out=None):
        """y = A x on the plan's pattern."""
        with torch.cuda.device(self.device):
            x = _dev_f64(x, self.device).reshape(-1)
            if out is None:
                out = torch.empty(self.ndof, dtype=torch.float64, device=self.device)
            check(_lib.lib().cpfem_spmv(self._h, _ptr(csr_data), _ptr(x), _ptr(out), _stream()), 'cpfem_spmv')
        return out

    def csr_diagonal(self, csr_data, invert=False):
        with torch.cuda.device(self.device):
            out = torch.empty(self.ndof, dtype=torch.float64, device=self.device)
            check(_lib.lib().cpfem_csr_diagonal(self._h, _ptr(csr_data), _ptr(out), int(bool(invert)), _stream()), 'cpfem_csr_diagonal')
        return out

    def bicgstab(self, csr_data, b, x0=None, precond=True, tol=1e-10, atol=1e-10, maxiter=10000):
        """jax.scipy.sparse.linalg.bicgstab(A, b, x0, M=Jacobi, tol, atol, maxiter) on the device-resident matrix
        (solver.py:34-40).  Returns (x, iterations, ||A x - b||)."""
        with torch.cuda.device(self.device):
            b = _dev_f64(b, self.device).reshape(-1)
            x = torch.zeros(self.ndof, dtype=torch.float64, device=self.device) if x0 is None else \
                _dev_f64(x0, self.device).reshape(-1).clone()
            info = (ctypes.c_int64 * 2)()
            resid = ctypes.c_double()
            check(_lib.lib().cpfem_bicgstab(self._h, _ptr(csr_data), _ptr(b), _ptr(x), int(bool(precond)), float(tol), float(atol),
                                            int(maxiter), info, ctypes.byref(resid), _stream()), 'cpfem_bicgstab')
        return x, int(info[0]), float(resid.value)


    def bicgstab_enqueue(self, csr_data, b, x, iters, precond=True, tol=1e-10, atol=1e-10, maxiter=10000, info=None, resid=None):
        """cpfem_bicgstab_enqueue: no host synchronisation.  `x` holds the start vector and receives the solution; `info`
        (int64[2]) and `resid` (float64[1]) are device tensors written when the stream gets there.  Returns (x, info, resid)."""
        with torch.cuda.device(self.device):
            if info is None:
                info = torch.zeros(2, dtype=torch.int64, device=self.device)
            if resid is None:
                resid = torch.zeros(1, dtype=torch.float64, device=self.device)
            check(_lib.lib().cpfem_bicgstab_enqueue(self._h, _ptr(csr_data), _ptr(b), _ptr(x), int(bool(precond)), float(tol),
                                                    float(atol), int(maxiter), int(iters), _ptr(info), _ptr(resid), _stream()),
                  'cpfem_bicgstab_enqueue')
        return x, info, resid


# ---- free helpers ------------------------------------------------------------------------------
def scatter_add(src, index_map, dst):
    check(_lib.lib().cpfem_scatter_add(_ptr(src), _ptr(index_map), int(src.numel()), _ptr(dst), _stream()), 'cpfem_scatter_add')


def gather(src, index_map, dst):
    check(_lib.lib().cpfem_gather(_ptr(src), _ptr(index_map), int(index_map.numel()), _ptr(dst), _stream()), 'cpfem_gather')


def sumsq(x, out):
    check(_lib.lib().cpfem_sumsq(_ptr(x), int(x.numel()), _ptr(out), _stream()), 'cpfem_sumsq')


def aos_to_soa(aos: torch.Tensor, comps: int) -> torch.Tensor:
    n = aos.numel() // comps
    out = torch.empty(comps, n, dtype=torch.float64, device=aos.device)
    check(_lib.lib().cpfem_aos_to_soa(_ptr(aos), n, comps, _ptr(out), _stream()), 'cpfem_aos_to_soa')
    return out


def soa_to_aos(soa: torch.Tensor, comps: int) -> torch.Tensor:
    n = soa.numel() // comps
    out = torch.empty(n, comps, dtype=torch.float64, device=soa.device)
    check(_lib.lib().cpfem_soa_to_aos(_ptr(soa), n, comps, _ptr(out), _stream()), 'cpfem_soa_to_aos')
    return out


def dfma_peak(iters=20000, repeats=3):
    """Measured FP64 FMA throughput of the current device in TFLOP/s (CUDA events, best of `repeats`)."""
    sink = torch.zeros(1, dtype=torch.float64, device='cuda')
    flops = ctypes.c_double()
    best = 0.0
    for _ in range(repeats + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_lib.lib().cpfem_dfma_peak_kernel(int(iters), _ptr(sink), ctypes.byref(flops), _stream()), 'cpfem_dfma_peak_kernel')
        e1.record()
        e1.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best
