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
    """The attributes of a meshio mesh that the reference drivers touch (points, cells_dict, cell_data; e.g.
    polycrystal_304steel.py:83-92) plus what Neper adds to its Gmsh files and meshio drops: node sets and the per-grain
    orientation block."""

    def __init__(self, points, cells, cell_data=None, field_data=None, nsets=None, orientations=None):
        self.points = points
        self.cells_dict = {'hexahedron': cells}
        self.cell_data = cell_data or {}
        self.field_data = field_data or {}          # physical name -> [tag, dimension] (meshio convention)
        self.nsets = nsets or {}                    # Neper $NSets: name -> 0-based node indices
        self.orientations = orientations            # Neper $ElsetOrientations: dict(descriptor, ids, values) or None


def rodrigues_to_quat(r):
    """Rodrigues vectors (n, 3) -> unit quaternions (w, x, y, z), the input convention of get_rot_mat
    (models_copper.py:37-45): r = tan(theta/2) axis, q = (1, r) / sqrt(1 + |r|^2)."""
    r = onp.atleast_2d(onp.asarray(r, dtype=onp.float64))
    q = onp.concatenate([onp.ones((len(r), 1)), r], axis=1)
    return q / onp.linalg.norm(q, axis=1, keepdims=True)


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
    """Gmsh 2.2 ASCII reader for the hex8 meshes the reference ships (Neper output, e.g.
    singlecrystal_copper/data/neper/singlecrystal_copper/mesh2.msh; Gmsh's own box.msh): $Nodes, the hex8 (type 5)
    records of $Elements with their physical / elementary tags (mesh2.msh:37-47), $PhysicalNames, and Neper's $NSets
    and $ElsetOrientations blocks (mesh2.msh:121-124).  Other element types (points, lines, quads) are skipped.
    Returns a meshio-like object: .points, .cells_dict['hexahedron'], .cell_data['gmsh:physical'][0] (what
    polycrystal_304steel.py:86 reads), .cell_data['gmsh:geometrical'][0], .field_data, .nsets, .orientations."""
    with open(path) as f:
        lines = f.read().split('\n')
    points, cells, phys, geom = None, [], [], []
    field_data, nsets, orientations = {}, {}, None
    node_index = None
    i = 0
    while i < len(lines):
        s = lines[i].strip()
        if s == '$MeshFormat':
            ver = lines[i + 1].split()
            if not ver[0].startswith('2') or int(ver[1]) != 0:
                raise ValueError(f'{path}: only Gmsh 2.x ASCII files are supported (got "{lines[i + 1].strip()}")')
            i += 2
        elif s == '$Nodes':
            n = int(lines[i + 1])
            arr = onp.array([lines[i + 2 + k].split() for k in range(n)], dtype=onp.float64)
            ids = arr[:, 0].astype(onp.int64)
            # node ids are 1-based and normally dense; a sparse numbering is compacted in file order
            if ids.min() == 1 and ids.max() == n:
                points = onp.zeros((n, 3))
                points[ids - 1] = arr[:, 1:4]
                node_index = None
            else:
                points = arr[:, 1:4].copy()
                node_index = {int(v): k for k, v in enumerate(ids)}
            i += n + 2
        elif s == '$Elements':
            n = int(lines[i + 1])
            for k in range(n):
                t = lines[i + 2 + k].split()
                if int(t[1]) == 5:
                    ntags = int(t[2])
                    phys.append(int(t[3]) if ntags > 0 else 0)
                    geom.append(int(t[4]) if ntags > 1 else 0)
                    nodes = [int(v) for v in t[3 + ntags:3 + ntags + 8]]
                    cells.append([v - 1 for v in nodes] if node_index is None else [node_index[v] for v in nodes])
            i += n + 2
        elif s == '$PhysicalNames':
            n = int(lines[i + 1])
            for k in range(n):
                t = lines[i + 2 + k].split(None, 2)
                field_data[t[2].strip().strip('"')] = onp.array([int(t[1]), int(t[0])])
            i += n + 2
        elif s == '$NSets':
            n = int(lines[i + 1])
            j = i + 2
            for _ in range(n):
                name = lines[j].strip()
                m = int(lines[j + 1])
                ids = onp.array([int(lines[j + 2 + k]) for k in range(m)], dtype=onp.int64)
                nsets[name] = ids - 1 if node_index is None else onp.array([node_index[int(v)] for v in ids])
                j += m + 2
            i = j
        elif s == '$ElsetOrientations':
            head = lines[i + 1].split()
            n = int(head[0])
            arr = onp.array([lines[i + 2 + k].split() for k in range(n)], dtype=onp.float64)
            orientations = {'descriptor': head[1] if len(head) > 1 else '', 'ids': arr[:, 0].astype(onp.int64), 'values': arr[:, 1:]}
            i += n + 2
        else:
            i += 1
    if points is None or not cells:
        raise ValueError(f'{path}: no $Nodes / hex8 elements found')
    return _MeshioLike(points, onp.array(cells, dtype=onp.int32),
                       {'gmsh:physical': [onp.array(phys, dtype=onp.int64)], 'gmsh:geometrical': [onp.array(geom, dtype=onp.int64)]},
                       field_data, nsets, orientations)
