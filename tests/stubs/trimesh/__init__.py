"""Stand-in for the `trimesh` package (not installed in this image; test infrastructure): the handful of operations robogym's
rearrange mesh utilities use (robogym/envs/rearrange/common/utils.py:244-281, 320-340, 485-530) -- load a binary STL, concatenate
meshes, the centre of mass (volume integral over the triangles, as trimesh computes it), affine transforms, uniform surface samples."""
import struct
import types

import numpy as np


class Trimesh:
    def __init__(self, vertices=None, faces=None, **_):
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)

    @property
    def triangles(self):
        return self.vertices[self.faces]

    @property
    def volume(self):
        a, b, c = (self.triangles[:, i] for i in range(3))
        return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)

    @property
    def center_mass(self):
        a, b, c = (self.triangles[:, i] for i in range(3))
        vol6 = np.einsum("ij,ij->i", a, np.cross(b, c))
        return ((a + b + c) * vol6[:, None]).sum(0) / (4.0 * vol6.sum())

    @property
    def bounds(self):
        return np.stack([self.vertices.min(0), self.vertices.max(0)])

    @property
    def extents(self):
        return self.vertices.max(0) - self.vertices.min(0)

    @property
    def bounding_box(self):
        return types.SimpleNamespace(extents=self.vertices.max(0) - self.vertices.min(0), bounds=self.bounds)

    @property
    def area_faces(self):
        t = self.triangles
        return 0.5 * np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1)

    def apply_transform(self, matrix):
        m = np.asarray(matrix, dtype=np.float64)
        self.vertices = self.vertices @ m[:3, :3].T + m[:3, 3]
        if np.linalg.det(m[:3, :3]) < 0:
            self.faces = self.faces[:, ::-1]
        return self

    def apply_scale(self, scale):
        self.vertices = self.vertices * np.asarray(scale, dtype=np.float64)
        return self

    def copy(self):
        return Trimesh(self.vertices.copy(), self.faces.copy())


def load(path, *_, **__):
    with open(path, "rb") as f:
        data = f.read()
    (ntri,) = struct.unpack_from("<I", data, 80)
    if 84 + 50 * ntri != len(data):
        raise ValueError(f"{path}: not a binary STL")
    rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", 9), ("a", "<u2")]), count=ntri, offset=84)
    verts = rec["v"].reshape(-1, 3).astype(np.float64)
    return Trimesh(verts, np.arange(len(verts)).reshape(-1, 3))


def _concatenate(meshes):
    meshes = list(meshes) if not isinstance(meshes, Trimesh) else [meshes]
    verts, faces, off = [], [], 0
    for m in meshes:
        verts.append(m.vertices)
        faces.append(m.faces + off)
        off += len(m.vertices)
    return Trimesh(np.concatenate(verts), np.concatenate(faces))


def _sample_surface(mesh, count):
    """`count` points uniformly distributed over the surface, and the index of the face each lies on"""
    rng = np.random.default_rng()
    area = mesh.area_faces
    idx = rng.choice(len(area), size=count, p=area / area.sum())
    t = mesh.triangles[idx]
    u, v = rng.random(count), rng.random(count)
    flip = u + v > 1.0
    u[flip], v[flip] = 1.0 - u[flip], 1.0 - v[flip]
    return t[:, 0] + u[:, None] * (t[:, 1] - t[:, 0]) + v[:, None] * (t[:, 2] - t[:, 0]), idx


def _subdivide(vertices, faces, **_):
    """one step of midpoint subdivision (every triangle into four)"""
    v = np.asarray(vertices, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    mids, new_faces = {}, []
    verts = list(v)

    def mid(i, j):
        key = (min(i, j), max(i, j))
        if key not in mids:
            mids[key] = len(verts)
            verts.append(0.5 * (v[i] + v[j]))
        return mids[key]

    for a, b, c in f:
        ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
        new_faces += [(a, ab, ca), (ab, b, bc), (ca, bc, c), (ab, bc, ca)]
    return np.array(verts), np.array(new_faces)


def _subdivide_to_size(vertices, faces, max_edge, max_iter=10, **_):
    """midpoint subdivision of the triangles that have an edge longer than max_edge, until none is left"""
    v = [np.asarray(x, dtype=np.float64) for x in np.asarray(vertices, dtype=np.float64)]
    f = [tuple(int(i) for i in t) for t in np.asarray(faces, dtype=np.int64)]
    for _ in range(max_iter):
        out, again, mids = [], False, {}

        def mid(i, j):
            key = (min(i, j), max(i, j))
            if key not in mids:
                mids[key] = len(v)
                v.append(0.5 * (v[i] + v[j]))
            return mids[key]

        for a, b, c in f:
            if max(np.linalg.norm(v[a] - v[b]), np.linalg.norm(v[b] - v[c]), np.linalg.norm(v[c] - v[a])) > max_edge:
                ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
                out += [(a, ab, ca), (ab, b, bc), (ca, bc, c), (ab, bc, ca)]
                again = True
            else:
                out.append((a, b, c))
        f = out
        if not again:
            break
    return np.array(v), np.array(f)


util = types.SimpleNamespace(concatenate=_concatenate)
sample = types.SimpleNamespace(sample_surface=_sample_surface)
remesh = types.SimpleNamespace(subdivide=_subdivide, subdivide_to_size=_subdivide_to_size)
