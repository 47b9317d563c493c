"""Synthetic polycrystal workload of SURVEY.md section 8(d) / BASELINE.json configs[4].

Mesh: N^3 unit-cube hex8 (generate_mesh.box_mesh numbering).  Grains: cubic blocks of `grain` cells per edge,
one orientation per grain from Bunge Euler angles (phi1, phi2 ~ U[0, 2 pi), cos Phi ~ U[-1, 1],
numpy default_rng(0)) -> quaternion (w,x,y,z) -> get_rot_mat (models_copper.py:37-45).
Load: affine uniaxial field u_z = eps z, u_x = -0.3 eps x, u_y = -0.3 eps y plus nodal noise
U(-1,1) * 1e-6 / N (default_rng(1)).  A run advances `steps` load steps of d_eps, dt with the state-update
kernel from the virgin state; the timed step is the next one.
"""
from __future__ import annotations

import numpy as onp

from .generate_mesh import Mesh, box_mesh


def bunge_to_quat(phi1, Phi, phi2):
    return onp.stack([onp.cos(Phi / 2) * onp.cos((phi1 + phi2) / 2), onp.sin(Phi / 2) * onp.cos((phi1 - phi2) / 2),
                      onp.sin(Phi / 2) * onp.sin((phi1 - phi2) / 2), onp.cos(Phi / 2) * onp.sin((phi1 + phi2) / 2)], axis=-1)


def grain_quaternions(n_grains, seed=0):
    rng = onp.random.default_rng(seed)
    phi1 = rng.uniform(0, 2 * onp.pi, n_grains)
    phi2 = rng.uniform(0, 2 * onp.pi, n_grains)
    Phi = onp.arccos(rng.uniform(-1, 1, n_grains))
    return bunge_to_quat(phi1, Phi, phi2)


def polycrystal(N, grain=8, seed=0, z_range=None):
    """Returns (Mesh, quat (n_grains,4), cell_ori_inds (nc,)).  z_range=(k0,k1) keeps only cell layers
    k0 <= k < k1 (slab of an element partition) with local node numbering; global node ids are returned too."""
    m = box_mesh(N, N, N, 1., 1., 1.)
    points, cells = m.points, m.cells_dict['hexahedron']
    G = (N + grain - 1) // grain
    k, j, i = onp.meshgrid(onp.arange(N), onp.arange(N), onp.arange(N), indexing='ij')
    gid = ((i // grain) + G * (j // grain) + G * G * (k // grain)).ravel()
    quat = grain_quaternions(G ** 3, seed)
    node_gid = None
    if z_range is not None:
        k0, k1 = z_range
        sel = slice(k0 * N * N, k1 * N * N)
        cells = cells[sel]
        gid = gid[sel]
        n0, n1 = k0 * (N + 1) ** 2, (k1 + 1) * (N + 1) ** 2
        node_gid = onp.arange(n0, n1)
        points = points[n0:n1]
        cells = cells - n0
    mesh = Mesh(points, cells)
    mesh.node_gid = node_gid
    return mesh, quat, gid


def noise_field(N, seed=1):
    """Nodal noise U(-1,1) * 1e-6 / N for the whole (N+1)^3 node set (slabs index it with their global node ids)."""
    rng = onp.random.default_rng(seed)
    return rng.uniform(-1, 1, size=((N + 1) ** 3, 3)) * (1e-6 / N)


def affine_displacement(points, eps):
    return onp.stack([-0.3 * eps * points[:, 0], -0.3 * eps * points[:, 1], eps * points[:, 2]], axis=1)


def displacement(points, eps, N, seed=1, node_gid=None):
    """Affine uniaxial field + nodal noise (same noise values whether the mesh is whole or cut into slabs)."""
    noise = noise_field(N, seed)
    if node_gid is not None:
        noise = noise[node_gid]
    return affine_displacement(points, eps) + noise
