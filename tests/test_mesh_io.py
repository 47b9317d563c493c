"""I/O edge of the drivers: the Gmsh-2.2 reader against the Neper / Gmsh meshes the reference ships
(fixtures copied by tests/golden/make_golden.py) and against the structured generator.  No GPU needed."""
import os

import numpy as np
import pytest

from cpfem_b200.generate_mesh import box_mesh, read_gmsh22_hex, rodrigues_to_quat

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_gmsh_reader_mesh2_single_crystal():
    """singlecrystal_copper/data/neper/singlecrystal_copper/mesh2.msh: 27 nodes, 8 hex8 records with three tags
    (:37-47), six Neper node sets, one physical name, one Rodrigues orientation (:121-124)."""
    m = read_gmsh22_hex(os.path.join(GOLD, 'mesh2.msh'))
    cells = m.cells_dict['hexahedron']
    assert m.points.shape == (27, 3) and cells.shape == (8, 8) and cells.dtype == np.int32
    # the file is a structured 2^3 box of edge 0.1 in Gmsh node order: identical to the generator
    b = box_mesh(2, 2, 2, 0.1, 0.1, 0.1)
    assert np.allclose(m.points, b.points, atol=1e-12) and np.array_equal(cells, b.cells_dict['hexahedron'])
    assert np.array_equal(m.cell_data['gmsh:physical'][0], np.ones(8, dtype=np.int64))
    assert np.array_equal(m.cell_data['gmsh:geometrical'][0], np.ones(8, dtype=np.int64))
    assert list(m.field_data) == ['poly1'] and list(m.field_data['poly1']) == [1, 3]
    assert sorted(m.nsets) == ['x0', 'x1', 'y0', 'y1', 'z0', 'z1']
    for name, axis, val in (('x0', 0, 0.0), ('x1', 0, 0.1), ('y0', 1, 0.0), ('y1', 1, 0.1), ('z0', 2, 0.0), ('z1', 2, 0.1)):
        want = np.where(np.isclose(m.points[:, axis], val))[0]
        assert np.array_equal(np.sort(m.nsets[name]), want), name
    o = m.orientations
    assert o['descriptor'] == 'rodrigues:active' and list(o['ids']) == [1]
    assert np.allclose(o['values'], [[1.385351469941, -0.456168566806, -0.469768220257]])
    # positive Jacobian for every cell (node order is the one the kernels expect)
    X = m.points[cells]
    e1, e2, e3 = X[:, 1] - X[:, 0], X[:, 3] - X[:, 0], X[:, 4] - X[:, 0]
    assert (np.einsum('ci,ci->c', np.cross(e1, e2), e3) > 0).all()


def test_gmsh_reader_polycrystal_tags_and_orientations():
    """polycrystal_304steel/data/neper/polycrystal_304steel/domain0_mesh5.msh: 5^3 cells in 8 grains; the drivers take
    cell_data['gmsh:physical'][0] - 1 as the grain index (polycrystal_304steel.py:86)."""
    m = read_gmsh22_hex(os.path.join(GOLD, 'domain0_mesh5.msh'))
    cells = m.cells_dict['hexahedron']
    assert m.points.shape == (216, 3) and cells.shape == (125, 8)
    phys = m.cell_data['gmsh:physical'][0]
    assert phys.shape == (125,) and set(phys.tolist()) == set(range(1, 9))
    assert phys[0] == 1 and phys[10] == 7                       # records 1 and 11 of the file (:228, :238)
    assert sorted(m.field_data) == [f'poly{i}' for i in range(1, 9)]
    o = m.orientations
    assert o['descriptor'] == 'rodrigues:active' and list(o['ids']) == list(range(1, 9)) and o['values'].shape == (8, 3)
    assert np.allclose(o['values'][7], [0.183166726246, -11.170774630101, -2.534524671934])
    q = rodrigues_to_quat(o['values'])
    assert q.shape == (8, 4) and np.allclose(np.linalg.norm(q, axis=1), 1.0)
    assert np.allclose(q[:, 1:] / q[:, :1], o['values'])
    b = box_mesh(5, 5, 5, *m.points.max(axis=0))
    assert np.allclose(m.points, b.points, atol=1e-12) and np.array_equal(cells, b.cells_dict['hexahedron'])


def test_gmsh_reader_skips_lower_dimensional_elements():
    """calibration/data/msh/box.msh (Gmsh's own output): 27 element records of which one is a hex8 with two tags."""
    m = read_gmsh22_hex(os.path.join(GOLD, 'box.msh'))
    cells = m.cells_dict['hexahedron']
    assert m.points.shape == (8, 3) and cells.shape == (1, 8)
    assert list(cells[0]) == [0, 1, 3, 2, 4, 5, 6, 7]
    assert list(m.cell_data['gmsh:physical'][0]) == [0] and list(m.cell_data['gmsh:geometrical'][0]) == [1]
    assert m.orientations is None and m.nsets == {}


def test_gmsh_reader_sparse_ids_and_errors(tmp_path):
    p = tmp_path / 'sparse.msh'
    p.write_text('$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n8\n' +
                 ''.join(f'{10 * (k + 1)} {k & 1} {(k >> 1) & 1} {(k >> 2) & 1}\n' for k in range(8)) +
                 '$EndNodes\n$Elements\n1\n1 5 2 4 9 10 20 40 30 50 60 80 70\n$EndElements\n')
    m = read_gmsh22_hex(str(p))
    assert list(m.cells_dict['hexahedron'][0]) == [0, 1, 3, 2, 4, 5, 7, 6] and list(m.cell_data['gmsh:physical'][0]) == [4]
    bad = tmp_path / 'v4.msh'
    bad.write_text('$MeshFormat\n4.1 0 8\n$EndMeshFormat\n')
    with pytest.raises(ValueError):
        read_gmsh22_hex(str(bad))
    empty = tmp_path / 'empty.msh'
    empty.write_text('$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n')
    with pytest.raises(ValueError):
        read_gmsh22_hex(str(empty))
