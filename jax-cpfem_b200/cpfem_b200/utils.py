"""I/O edge of the drivers: `save_sol` with the signature the reference imports from
`jax_fem.utils` (e.g. singlecrystal_tantalum/singlecrystal_tantalum.py:16,251), writing the same kind of file the
reference commits under */data/vtk (VTK XML UnstructuredGrid, binary DataArrays = base64(uint32 header) +
base64(zlib blocks), `sol` and cell data as Float32, `Points` Float64, `connectivity` Int32), and `read_vtu` to read
such files back (no meshio in this image)."""
from __future__ import annotations

import base64
import re
import struct
import zlib

import numpy as onp

_VTK_TYPES = {'Float32': onp.float32, 'Float64': onp.float64, 'Int32': onp.int32, 'Int64': onp.int64, 'UInt8': onp.uint8}
_VTK_NAMES = {onp.dtype(v): k for k, v in _VTK_TYPES.items()}


def _to_host(x):
    if hasattr(x, 'detach'):
        x = x.detach().cpu().numpy()
    return onp.asarray(x)


def _encode(arr):
    raw = onp.ascontiguousarray(arr).tobytes()
    comp = zlib.compress(raw)
    hdr = struct.pack('<4I', 1, len(raw), len(raw), len(comp))
    return (base64.b64encode(hdr) + base64.b64encode(comp)).decode()


def _data_array(name, arr, ncomp=None):
    arr = onp.ascontiguousarray(arr)
    attrs = f'type="{_VTK_NAMES[arr.dtype]}" Name="{name}" format="binary"'
    if ncomp:
        attrs += f' NumberOfComponents="{ncomp}"'
    return f'<DataArray {attrs}>{_encode(arr)}</DataArray>\n'


def save_sol(fe, sol, sol_file, cell_infos=None, point_infos=None, cell_type='hexahedron'):
    """jax_fem.utils.save_sol(fe, sol, sol_file, cell_infos=[(name, per-cell array), ...], point_infos=[...])."""
    points = onp.asarray(fe.points, dtype=onp.float64)
    cells = onp.asarray(fe.cells, dtype=onp.int32)
    sol = _to_host(sol).astype(onp.float32)
    nc = len(cells)
    out = ['<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian" '
           'header_type="UInt32" compressor="vtkZLibDataCompressor">\n<UnstructuredGrid>\n'
           f'<Piece NumberOfPoints="{len(points)}" NumberOfCells="{nc}">\n<Points>\n',
           _data_array('Points', points, 3), '</Points>\n<Cells>\n',
           _data_array('connectivity', cells.reshape(-1)),
           _data_array('offsets', (8 * onp.arange(1, nc + 1)).astype(onp.int32)),
           _data_array('types', onp.full(nc, 12, dtype=onp.int64)), '</Cells>\n<PointData>\n',
           _data_array('sol', sol, sol.shape[1] if sol.ndim > 1 else None)]
    for name, v in (point_infos or []):
        v = _to_host(v).astype(onp.float32)
        out.append(_data_array(name, v, v.shape[1] if v.ndim > 1 else None))
    out.append('</PointData>\n<CellData>\n')
    for name, v in (cell_infos or []):
        v = _to_host(v).astype(onp.float32)
        assert len(v) == nc, f'cell data {name}: {len(v)} values for {nc} cells'
        out.append(_data_array(name, v, v.shape[1] if v.ndim > 1 else None))
    out.append('</CellData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n')
    with open(sol_file, 'w') as f:
        f.write(''.join(out))


def read_vtu(path):
    """Binary DataArrays of a VTU file as {name: array} (the encoding described in the module docstring)."""
    txt = open(path).read()
    out = {}
    for m in re.finditer(r'<DataArray([^>]*)>(.*?)</DataArray>', txt, re.S):
        attrs = dict(re.findall(r'(\w+)="([^"]*)"', m.group(1)))
        if attrs.get('format') != 'binary':
            continue
        raw = m.group(2).strip().encode()
        nblocks = struct.unpack('<I', base64.b64decode(raw[:24])[:4])[0]
        hbytes = 4 * (3 + nblocks)
        hlen = ((hbytes + 2) // 3) * 4
        hdr = onp.frombuffer(base64.b64decode(raw[:hlen])[:hbytes], dtype=onp.uint32)
        data = base64.b64decode(raw[hlen:])
        buf, off = b'', 0
        for c in hdr[3:]:
            buf += zlib.decompress(data[off:off + int(c)])
            off += int(c)
        arr = onp.frombuffer(buf, dtype=_VTK_TYPES[attrs['type']])
        ncomp = int(attrs.get('NumberOfComponents', 1))
        out[attrs.get('Name', 'noname')] = arr.reshape(-1, ncomp) if ncomp > 1 else arr
    return out
