"""Mesh helpers with the names the reference drivers import from `jax_fem.generate_mesh`
(`Mesh`, `box_mesh`, `get_meshio_cell_type`; e.g. singlecrystal_copper/singlecrystal_copper.py:14,69-78)
plus a Gmsh-2.2 ASCII hex8 reader for the Neper meshes the reference ships (no meshio in this image).
"""
from __future__ import annotations

import numpy as onp


class Mesh:
    """jax_fem.generate_mesh.Mesh: points (nnodes, 3), cells (nc, 8) in meshio/Gmsh hex8 node order."""

    def __init__(self, points, cells, ele_type='HEX8'):
        self.points = onp.ascontiguousarray(points, dtype=onp.float64)
        self.cells = onp.ascontiguousarray(cells, dtype=onp.int32)
        self.ele_type = ele_type


def get_meshio_cell_type(ele_type):
    if ele_type != 'HEX8':
        raise NotImplementedError('only HEX8 is on the crystal-plasticity hot path')
    return 'hexahedron'


class _MeshioLike:
    def __init__(self, points, cells, cell_data=None):
        self.points = points
        self.cells_dict = {'hexahedron': cells}
        self.cell_data = cell_data or {}


def box_mesh(Nx, Ny, Nz, Lx=1., Ly=1., Lz=1.):
    """Structured hex8 box: node id = ix + (Nx+1) iy + (Nx+1)(Ny+1) iz, cell id x-fastest, Gmsh node order
    [n000,n100,n110,n010,n001,n101,n111,n011] (the numbering of the Neper files, e.g. mesh2.msh:39)."""
    xs, ys, zs = onp.linspace(0, Lx, Nx + 1), onp.linspace(0, Ly, Ny + 1), onp.linspace(0, Lz, Nz + 1)
    Z, Y, X = onp.meshgrid(zs, ys, xs, indexing='ij')
    points = onp.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    sx, sy = Nx + 1, (Nx + 1) * (Ny + 1)
    k, j, i = onp.meshgrid(onp.arange(Nz, dtype=onp.int64), onp.arange(Ny, dtype=onp.int64),
                           onp.arange(Nx, dtype=onp.int64), indexing='ij')
    n0 = (i + sx * j + sy * k).ravel()
    cells = onp.stack([n0, n0 + 1, n0 + 1 + sx, n0 + sx, n0 + sy, n0 + 1 + sy, n0 + 1 + sx + sy, n0 + sx + sy], axis=1)
    return _MeshioLike(points, cells.astype(onp.int32))


def read_gmsh22_hex(path):
    """Minimal Gmsh 2.2 ASCII reader: $Nodes and the hex8 (type 5) records of $Elements.
    Returns a meshio-like object with .points, .cells_dict['hexahedron'], .cell_data['gmsh:physical'][0]."""
    with open(path) as f:
        lines = f.read().split('\n')
    it = iter(range(len(lines)))
    points, cells, phys = None, [], []
    i = 0
    while i < len(lines):
        s = lines[i].strip()
        if s == '$Nodes':
            n = int(lines[i + 1])
            arr = onp.array([lines[i + 2 + k].split() for k in range(n)], dtype=onp.float64)
            ids = arr[:, 0].astype(onp.int64)
            points = onp.zeros((ids.max(), 3))
            points[ids - 1] = arr[:, 1:4]
            i += n + 2
        elif s == '$Elements':
            n = int(lines[i + 1])
            for k in range(n):
                t = lines[i + 2 + k].split()
                if int(t[1]) == 5:
                    ntags = int(t[2])
                    phys.append(int(t[3]) if ntags > 0 else 0)
                    cells.append([int(v) - 1 for v in t[3 + ntags:3 + ntags + 8]])
            i += n + 2
        else:
            i += 1
    return _MeshioLike(points, onp.array(cells, dtype=onp.int32), {'gmsh:physical': [onp.array(phys, dtype=onp.int64)]})
