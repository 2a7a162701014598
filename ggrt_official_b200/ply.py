"""3DGS-style .ply export / import of a Gaussian set.

Mirrors /root/reference/ggrt/model/pixelsplat/ply_export.py:26-92 (`export_ply`): same scene
normalisation (median to the origin, 95 % quantile to unit range), same viewer rotation
(+Z up, -45 degrees about Z, camera space of `extrinsics`), same vertex properties in the same
order (x y z nx ny nz f_dc_0..2 opacity scale_0..2 rot_0..3, all float32, scales as logs,
quaternions as wxyz) and -- like the reference, :75-77 -- only the DC band of the harmonics.
The reference writes through the `plyfile` package; this module writes the identical
`binary_little_endian 1.0` layout directly with numpy and can read it back (`import_ply`), so
exported scenes can be fed to the rasterizer benchmark without extra dependencies.
CPU-side I/O: not on the B200 hot path.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List

import numpy as np
import torch
from torch import Tensor


def construct_list_of_attributes(num_rest: int) -> List[str]:
    """ply_export.py:13-23."""
    attributes = ["x", "y", "z", "nx", "ny", "nz"]
    attributes += [f"f_dc_{i}" for i in range(3)]
    attributes += [f"f_rest_{i}" for i in range(num_rest)]
    attributes.append("opacity")
    attributes += [f"scale_{i}" for i in range(3)]
    attributes += [f"rot_{i}" for i in range(4)]
    return attributes


def _viewer_rotation(extrinsics: Tensor) -> Tensor:
    """ply_export.py:43-63: +Z up, the viewer's 45 degree start angle, then the view's world-to-camera rotation."""
    from scipy.spatial.transform import Rotation as R

    rotation = torch.tensor([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], dtype=torch.float32, device=extrinsics.device)
    adjustment = torch.tensor(R.from_rotvec([0, 0, -45], True).as_matrix(), dtype=torch.float32,
                              device=extrinsics.device)
    return (adjustment @ rotation) @ extrinsics[:3, :3].inverse()


def vertex_elements(extrinsics: Tensor, means: Tensor, scales: Tensor, rotations: Tensor, harmonics: Tensor,
                    opacities: Tensor) -> np.ndarray:
    """The structured vertex array `export_ply` writes (one record per Gaussian)."""
    from scipy.spatial.transform import Rotation as R

    means = means - means.median(dim=0).values
    scale_factor = means.abs().quantile(0.95, dim=0).max()
    means = means / scale_factor
    scales = scales / scale_factor
    rotation = _viewer_rotation(extrinsics)
    means = torch.einsum("ij,...j->...i", rotation, means)
    rot = R.from_quat(rotations.detach().cpu().numpy()).as_matrix()
    rot = rotation.detach().cpu().numpy() @ rot
    x, y, z, w = np.moveaxis(R.from_matrix(rot).as_quat(), -1, 0)
    quat_wxyz = np.stack((w, x, y, z), axis=-1)
    dc = harmonics[..., 0]  # DC band only (ply_export.py:75-77)
    columns = np.concatenate((
        means.detach().cpu().numpy(),
        np.zeros_like(means.detach().cpu().numpy()),
        dc.detach().cpu().contiguous().numpy(),
        opacities[..., None].detach().cpu().numpy(),
        scales.log().detach().cpu().numpy(),
        quat_wxyz,
    ), axis=1)
    names = construct_list_of_attributes(0)
    elements = np.empty(columns.shape[0], dtype=[(n, "<f4") for n in names])
    for k, n in enumerate(names):
        elements[n] = columns[:, k].astype(np.float32)
    return elements


def write_ply(elements: np.ndarray, path: Path) -> None:
    """binary_little_endian PLY with a single `vertex` element of float32 properties."""
    path = Path(path)
    path.parent.mkdir(exist_ok=True, parents=True)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {elements.shape[0]}"]
    header += [f"property float {n}" for n in elements.dtype.names]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(np.ascontiguousarray(elements).tobytes())


def export_ply(extrinsics: Tensor, means: Tensor, scales: Tensor, rotations: Tensor, harmonics: Tensor,
               opacities: Tensor, path: Path) -> None:
    """Same signature and file contents as the reference's `export_ply`."""
    write_ply(vertex_elements(extrinsics, means, scales, rotations, harmonics, opacities), Path(path))


_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
              "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
              "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


def read_ply(path: Path) -> np.ndarray:
    """Structured array of the `vertex` element of a binary_little_endian or ascii PLY without list properties."""
    data = Path(path).read_bytes()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").splitlines()
    if lines[0].strip() != "ply":
        raise ValueError(f"{path}: not a PLY file")
    fmt, count, props, in_vertex = None, 0, [], False
    for ln in lines[1:]:
        tok = ln.split()
        if not tok or tok[0] == "comment":
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            in_vertex = tok[1] == "vertex"
            if in_vertex:
                count = int(tok[2])
            elif count and props:
                break  # elements after `vertex` are ignored
        elif tok[0] == "property" and in_vertex:
            if tok[1] == "list":
                raise ValueError(f"{path}: list properties are not supported")
            props.append((tok[2], _PLY_TYPES[tok[1]]))
    dtype = np.dtype(props)
    if fmt == "binary_little_endian":
        return np.frombuffer(data, dtype=dtype, count=count, offset=end).copy()
    if fmt == "ascii":
        rows = np.loadtxt(data[end:].decode("ascii").splitlines()[:count], dtype=np.float64, ndmin=2)
        out = np.empty(count, dtype=dtype)
        for k, (n, _) in enumerate(props):
            out[n] = rows[:, k]
        return out
    raise ValueError(f"{path}: unsupported PLY format {fmt!r}")


def import_ply(path: Path, device="cpu") -> Dict[str, Tensor]:
    """Reads a 3DGS-style PLY (this module's, the reference's, or a stock 3DGS checkpoint with f_rest_*) into the
    tensors the rasterizer consumes: means [P,3], scales [P,3] (exp of the stored logs), rotations [P,4] xyzw,
    harmonics [P,3,K] (DC band plus any f_rest_* bands, channel-major as 3DGS stores them), opacities [P] (as
    stored: the reference writes probabilities, stock 3DGS writes logits) and covariances [P,3,3] = R S S^T R^T
    (gaussians.py:33-44)."""
    v = read_ply(path)
    names = v.dtype.names
    col = lambda *ns: np.stack([v[n].astype(np.float32) for n in ns], axis=-1)
    means = col("x", "y", "z")
    dc = col("f_dc_0", "f_dc_1", "f_dc_2")
    n_rest = sum(1 for n in names if n.startswith("f_rest_"))
    if n_rest % 3:
        raise ValueError(f"{path}: {n_rest} f_rest_* properties is not a multiple of 3")
    harm = dc[:, :, None]
    if n_rest:
        rest = col(*[f"f_rest_{i}" for i in range(n_rest)]).reshape(-1, 3, n_rest // 3)
        harm = np.concatenate((harm, rest), axis=2)
    scales = np.exp(col("scale_0", "scale_1", "scale_2"))
    w, x, y, z = (v[f"rot_{i}"].astype(np.float32) for i in range(4))
    quat = np.stack((x, y, z, w), axis=-1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    out = dict(means=t(means), scales=t(scales), rotations=t(quat), harmonics=t(harm), opacities=t(v["opacity"].astype(np.float32)))
    q = out["rotations"] / out["rotations"].norm(dim=-1, keepdim=True).clamp_min(1e-12)
    i, j, k, r = q.unbind(-1)
    rot = torch.stack((1 - 2 * (j * j + k * k), 2 * (i * j - k * r), 2 * (i * k + j * r),
                       2 * (i * j + k * r), 1 - 2 * (i * i + k * k), 2 * (j * k - i * r),
                       2 * (i * k - j * r), 2 * (j * k + i * r), 1 - 2 * (i * i + j * j)), -1).reshape(-1, 3, 3)
    rs = rot * out["scales"][:, None, :]
    out["covariances"] = rs @ rs.transpose(-1, -2)
    return out
