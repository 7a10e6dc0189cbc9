"""Counterpart of the reference's src/common/obj_utils.py: `load_obj(filename, normalize=True)` and
`save_obj(filename, vertices, faces, normals=None)` -- same names, arguments, results and file format.

Host-side IO.  The one computation in it, the vertex normals of a file that brings none (obj_utils.py:54-56),
runs on the CUDA kernels of meshes.compute_vertex_normals (no CPU path: such a file needs a CUDA device).
"""
import warnings

import numpy as np
import torch

from . import meshes


def _parse(filename):
    """-> (vertices list, vn list, faces list, face-vertex normal references as (vertex id, normal id) in file
    order).  `v`, `vn` and `f` records only; the first three vertices of a face (obj_utils.py:27-46)."""
    vertices, all_normals, faces, normal_refs = [], [], [], []
    warned = False
    with open(filename) as f:
        for line in f:
            parts = line.split()
            if not parts:
                continue
            if parts[0] == "v":
                vertices.append([float(v) for v in parts[1:4]])
            elif parts[0] == "vn":
                all_normals.append([float(v) for v in parts[1:4]])
            elif parts[0] == "f":
                corners = parts[1:]
                if len(corners) > 3 and not warned:
                    warnings.warn("%s: faces with more than 3 vertices; the extra vertices are skipped" % filename)
                    warned = True
                fields = [c.split("/") for c in corners[:3]]
                faces.append([int(fl[0]) for fl in fields])
                if len(fields[0]) > 2:               # `f v1//vn1 v2//vn2 v3//vn3`
                    normal_refs.extend((int(fl[0]) - 1, int(fl[2]) - 1) for fl in fields)
    return vertices, all_normals, faces, normal_refs


def load_obj(filename, normalize=True):
    """Loads a Wavefront .obj file: vertices [V,3] float32, faces [T,3] int32 (zero-based), normals [V,3]
    float32.  Face-vertex normals are averaged to one normal per vertex (a vertex without any gets (1,1,1)
    before normalisation, obj_utils.py:58-68); a file without normals gets compute_vertex_normals.
    With normalize=True the mesh is moved into a cube of side 2 around the origin (obj_utils.py:70-75)."""
    v_list, vn_list, f_list, normal_refs = _parse(filename)
    vertices = torch.tensor(v_list, dtype=torch.float32).reshape(-1, 3)
    faces = torch.tensor(f_list, dtype=torch.int32).reshape(-1, 3) - 1
    if not normal_refs:
        normals = meshes.compute_vertex_normals(vertices[None, :, :], faces)[0]
    else:
        all_normals = np.asarray(vn_list, dtype=np.float32).reshape(-1, 3)
        refs = np.asarray(normal_refs, dtype=np.int64)
        count = np.bincount(refs[:, 0], minlength=len(vertices)).astype(np.float32)
        summed = np.zeros((len(vertices), 3), dtype=np.float32)
        # unbuffered, in file order: the reference's `normals[i] += all_normals[j] / n` loop
        np.add.at(summed, refs[:, 0], all_normals[refs[:, 1]] / count[refs[:, 0], None])
        summed[count == 0] = 1.0
        normals = torch.nn.functional.normalize(torch.from_numpy(summed), p=2.0, dim=1)
    if normalize and len(vertices):
        vertices -= vertices.min(0)[0][None, :]
        vertices /= torch.abs(vertices).max()
        vertices *= 2
        vertices -= vertices.max(0)[0][None, :] / 2
    return vertices, faces, normals


def save_obj(filename, vertices, faces, normals=None):
    """Writes `v`, `f` (as `f a//a b//b c//c` when normals are given) and `vn` records, one-based indices,
    numbers in Python's shortest round-trip form (obj_utils.py:79-110)."""
    if len(vertices.shape) != 2 or vertices.shape[1] != 3:
        raise ValueError("vertices must have shape [vertex_count, 3]")
    if len(faces.shape) != 2 or faces.shape[1] != 3:
        raise ValueError("faces must have shape [triangle_count, 3]")
    if normals is not None:
        if len(normals.shape) != 2 or normals.shape[1] != 3:
            raise ValueError("normals must have shape [vertex_count, 3]")
    rows = vertices.detach().cpu().tolist()
    ids = (faces.detach().cpu().to(torch.int64) + 1).tolist()
    with open(filename, "w") as f:
        f.writelines("v {} {} {}\n".format(*row) for row in rows)
        if normals is not None:
            f.writelines("f {0}//{0} {1}//{1} {2}//{2}\n".format(*face) for face in ids)
            f.writelines("vn {} {} {}\n".format(*row) for row in normals.detach().cpu().tolist())
        else:
            f.writelines("f {} {} {}\n".format(*face) for face in ids)
