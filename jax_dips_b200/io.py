"""VTK output of grid fields: `write_vtk_manual(gstate, field_dict, filename)` as the reference's drivers call
it (jax_dips/utils/io.py:80-89; e.g. examples/benchmark_LPBE/main.py, tests/test_poisson.py:262-274).

The reference goes through `pyevtk.hl.structuredToVTK`, which is not in this image; the writer below emits
the same file: a VTK XML StructuredGrid (`<filename>.vts`, version 1.0, little endian, UInt64 block headers,
raw appended data), points = meshgrid(x, y, z, indexing="ij"), one PointData array per field, arrays laid out
with x fastest as VTK requires (pyevtk ravels in Fortran order).  Host-side file format only: nothing here is
on the training path.
"""
from __future__ import annotations

import os
import struct
from typing import Dict

import numpy as np

_VTK_TYPES = {np.dtype("float32"): "Float32", np.dtype("float64"): "Float64", np.dtype("int32"): "Int32",
              np.dtype("int64"): "Int64", np.dtype("uint8"): "UInt8", np.dtype("int8"): "Int8"}


def _host(a) -> np.ndarray:
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a)


def structuredToVTK(path: str, x: np.ndarray, y: np.ndarray, z: np.ndarray, pointData: Dict[str, np.ndarray]) -> str:
    """pyevtk.hl.structuredToVTK for point data on a structured grid given by 3-D coordinate arrays."""
    x, y, z = (np.ascontiguousarray(_host(a)) for a in (x, y, z))
    assert x.ndim == 3 and x.shape == y.shape == z.shape
    nx, ny, nz = x.shape
    extent = f"0 {nx - 1} 0 {ny - 1} 0 {nz - 1}"
    blocks, offset = [], 0

    def add(arrays, ncomp):
        nonlocal offset
        arrays = [np.ascontiguousarray(a) for a in arrays]
        data = (np.stack([a.ravel(order="F") for a in arrays], axis=1).ravel() if ncomp > 1
                else arrays[0].ravel(order="F"))
        raw = data.tobytes()
        off = offset
        blocks.append(struct.pack("<Q", len(raw)) + raw)
        offset += 8 + len(raw)
        return off

    lines = ['<?xml version="1.0"?>',
             '<VTKFile type="StructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">',
             f'<StructuredGrid WholeExtent="{extent}">', f'<Piece Extent="{extent}">']
    names = list(pointData.keys())
    lines.append(f'<PointData scalars="{names[0]}">' if names else "<PointData>")
    for name in names:
        a = _host(pointData[name])
        if a.ndim == 1:
            a = a.reshape(nx, ny, nz)
        assert a.shape == (nx, ny, nz), f"field {name!r} has shape {a.shape}, grid is {(nx, ny, nz)}"
        if a.dtype not in _VTK_TYPES:
            a = a.astype(np.float32)
        off = add([a], 1)
        lines.append(f'<DataArray type="{_VTK_TYPES[a.dtype]}" Name="{name}" NumberOfComponents="1" '
                     f'format="appended" offset="{off}"/>')
    lines.append("</PointData>")
    lines.append("<CellData>")
    lines.append("</CellData>")
    lines.append("<Points>")
    coords = [a if a.dtype in (np.float32, np.float64) else a.astype(np.float64) for a in (x, y, z)]
    off = add(coords, 3)
    lines.append(f'<DataArray type="{_VTK_TYPES[coords[0].dtype]}" Name="points" NumberOfComponents="3" '
                 f'format="appended" offset="{off}"/>')
    lines += ["</Points>", "</Piece>", "</StructuredGrid>", '<AppendedData encoding="raw">']
    out = path + ".vts"
    with open(out, "wb") as f:
        f.write(("\n".join(lines) + "\n_").encode("ascii"))
        for b in blocks:
            f.write(b)
        f.write(b"\n</AppendedData>\n</VTKFile>\n")
    return out


def write_vtk_manual(gstate, field_dict, filename="results/manual_dump"):
    """jax_dips/utils/io.py:80-89: every entry of `field_dict` (flat, z-fastest like gstate.R, or (Nx,Ny,Nz))
    as point data on the grid of `gstate`; writes `<filename>.vts`."""
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
    X, Y, Z = np.meshgrid(_host(gstate.x), _host(gstate.y), _host(gstate.z), indexing="ij")
    host = {}
    for name in field_dict.keys():
        a = _host(field_dict[name])
        host[name] = a.reshape(X.shape) if a.size == X.size else a
    return structuredToVTK(filename, X, Y, Z, pointData=host)


def write_vtk_solution(gstate, log, address="results/", maxsteps=None):
    """jax_dips/utils/io.py:60-77"""
    os.makedirs(address, exist_ok=True)
    X, Y, Z = np.meshgrid(_host(gstate.x), _host(gstate.y), _host(gstate.z), indexing="ij")
    for i in range(len(log["U"])):
        structuredToVTK(address + "/solution" + str(i).zfill(4), X, Y, Z,
                        pointData={"sol": _host(log["U"][i]).reshape(X.shape)})
        if maxsteps and i >= maxsteps - 1:
            break


def read_vts(path: str):
    """Minimal reader of the files written above (tests and quick inspection): returns (points (n,3), {name: array})."""
    import re
    raw = open(path, "rb").read()
    head, _, tail = raw.partition(b'<AppendedData encoding="raw">')
    start = tail.index(b"_") + 1
    body = tail[start:]
    text = head.decode("ascii")
    ext = [int(v) for v in re.search(r'WholeExtent="([^"]+)"', text).group(1).split()]
    shape = (ext[1] + 1, ext[3] + 1, ext[5] + 1)
    out, points = {}, None
    inv = {v: k for k, v in _VTK_TYPES.items()}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="([^"]+)" NumberOfComponents="(\d+)" format="appended" '
                         r'offset="(\d+)"/>', text):
        typ, name, nc, off = m.group(1), m.group(2), int(m.group(3)), int(m.group(4))
        n = struct.unpack("<Q", body[off:off + 8])[0]
        a = np.frombuffer(body[off + 8: off + 8 + n], dtype=inv[typ])
        if name == "points":
            points = a.reshape(-1, 3)
        else:
            out[name] = a.reshape(shape, order="F")
    return points, out
