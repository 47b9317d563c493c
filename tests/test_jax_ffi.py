"""XLA-FFI adapter (row G1: "thin C-ABI exposed as JAX FFI custom calls").  The C++ side is type-checked against a
stand-in header everywhere (no JAX needed); the JAX side runs only where `import jax` works - it cannot in the
container this repo is built in (no index), so those tests skip there."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ffi_adapter_type_checks():
    """csrc/cpfem_ffi.cc parses and every handler's signature matches its Bind() chain (tests/ffi_stub stands in for
    xla/ffi/api/ffi.h; the static_assert in XLA_FFI_DEFINE_HANDLER_SYMBOL is the check a real build performs)."""
    cuda_inc = '/usr/local/cuda/include'
    if not os.path.exists(os.path.join(cuda_inc, 'cuda_runtime_api.h')):
        pytest.skip('CUDA headers not found')
    r = subprocess.run(['g++', '-std=c++17', '-fsyntax-only', '-Wall', '-I', os.path.join(ROOT, 'tests', 'ffi_stub'), '-I', cuda_inc,
                        '-I', os.path.join(ROOT, 'include'), os.path.join(ROOT, 'jax-cpfem_b200', 'csrc', 'cpfem_ffi.cc')],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_ffi_targets_match_the_source():
    """Every target jax_ffi.register() asks for is defined in cpfem_ffi.cc, and vice versa."""
    from cpfem_b200 import jax_ffi
    src = open(os.path.join(ROOT, 'jax-cpfem_b200', 'csrc', 'cpfem_ffi.cc')).read()
    import re
    defined = set(re.findall(r'XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),', src))
    assert defined == set(jax_ffi.TARGETS)


def test_param_presets_match_the_torch_mirror():
    """param_sets.PRESETS (used by the JAX shim) hold the numbers of the models_*.py mirrors (and so of the reference)."""
    pytest.importorskip('torch')
    from cpfem_b200 import param_sets
    from cpfem_b200 import models_copper, models_tantalum, models_304steel, models_DPsteel_inhomo
    for name, mod in (('copper', models_copper), ('tantalum', models_tantalum), ('304steel', models_304steel)):
        c, p = mod.CrystalPlasticity, param_sets.PRESETS[name]
        m = p['material']
        assert (m['C11'], m['C12'], m['C44'], m['h'], m['t_sat'], m['gss_a'], m['xm']) == (c.C11, c.C12, c.C44, c.h, c.t_sat, c.gss_a, c.xm)
        assert m['max_sub_step'] == c.max_sub_step and p['gss_initial'] == c.gss_initial and m['r'] == c.r and m['ao'] == c.ao
        assert np.array_equal(p['slip'], c.slip_file)
    d, ph = param_sets.PRESETS['dpsteel'], models_DPsteel_inhomo.CrystalPlasticity.phase
    m = d['material']
    assert (m['C11'], m['C12'], m['C44'], m['h'], m['t_sat'], m['gss_a'], m['xm']) == tuple(ph[k][0] for k in ('C11', 'C12', 'C44', 'h0', 't_sat0', 'gss_a0', 'xm0'))
    assert list(param_sets.MATERIAL_FIELDS) == [f[0] for f in __import__('cpfem_b200._lib', fromlist=['Material']).Material._fields_]


@pytest.mark.gpu
def test_jax_ffi_round_trip():
    """On a box with JAX + a GPU: build and register the handlers, run update_state / newton_update through
    jax.ffi.ffi_call and compare with the ctypes path."""
    jax = pytest.importorskip('jax')
    import torch
    from cpfem_b200 import Plan, jax_ffi, make_material
    from cpfem_b200.param_sets import PRESETS
    import cpfem_oracle as O
    jax.config.update('jax_enable_x64', True)
    import jax.numpy as jnp
    jax_ffi.register()
    ps = PRESETS['304steel']
    pts, cells = O.box_mesh(3, 3, 3)
    jp = jax_ffi.JaxPlan(cells, pts, ps['slip'])
    nc = len(cells)
    rng = np.random.default_rng(0)
    q = rng.normal(size=(4, 4)); q /= np.linalg.norm(q, axis=1)[:, None]
    R = O.get_rot_mat(q)[rng.integers(0, 4, size=nc)]
    state = [np.tile(np.eye(3)[None, None], (nc, 8, 1, 1)), np.full((nc, 8, 12), 90.0), np.zeros((nc, 8, 12)),
             np.repeat(R[:, None], 8, axis=1)]
    eps = 3e-3
    sol = np.stack([-0.3 * eps * pts[:, 0], -0.3 * eps * pts[:, 1], eps * pts[:, 2]], 1)
    ffi = jax_ffi.jax_ffi_module()
    out = tuple(jax.ShapeDtypeStruct(s.shape, jnp.float64) for s in state[:3]) + (jax.ShapeDtypeStruct((4,), jnp.int64),)
    Fp, g, sl, status = ffi.ffi_call('cpfem_update_state_ffi', out)(jnp.asarray(sol), *[jnp.asarray(s) for s in state],
                                                                   plan=jp.handle, dt=np.float64(2e-3), mat=jax_ffi.material_attr(ps['material']))
    m = ps['material']
    plan = Plan(cells, pts, ps['slip'])
    mat = make_material(m['C11'], m['C12'], m['C44'], m['h'], m['t_sat'], m['gss_a'], m['xm'], m['r'], m['ao'], m['tol'], m['max_sub_step'])
    ref = plan.update_state(mat, sol, state, 2e-3)
    for a, b in zip((Fp, g, sl), ref):
        assert np.array_equal(np.asarray(a), b.cpu().numpy())
    assert int(status[0]) == 0 and int(status[2]) > 3
