"""Pack / unpack the compiled-model blob described by include/rg_model_fields.h.

The blob is what crosses the C ABI (`rg_model_load`, include/robogym_b200.h); it
plays the role of the `mjModel` that `mujoco_py.load_model_from_xml` returns in
the reference (robogym/mujoco/mujoco_xml.py:259).
"""
import os
import re
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS_H = os.path.join(_HERE, "..", "include", "rg_model_fields.h")
MAGIC = b"RGMODEL1"
NAMES_MAGIC = b"RGNAMES1"   # optional trailing section: the name tables behind rg_model_name2id (mjModel.*_name2id)

_pat = re.compile(r"^\s*RG_(DIM|IB|FB|I|F)\(\s*(\w+)\s*(?:,\s*(.+?)\s*)?\)\s*(?:/\*.*)?$")


def field_list(path=FIELDS_H):
    """Return (dims, arrays): dims = [name], arrays = [(kind, name, count_expr)]."""
    dims, arrays = [], []
    with open(path) as f:
        for line in f:
            m = _pat.match(line)
            if not m:
                continue
            kind, name, cnt = m.groups()
            kind = {"IB": "I", "FB": "F"}.get(kind, kind)
            if kind == "DIM":
                dims.append(name)
            else:
                arrays.append((kind, name, cnt))
    return dims, arrays


DIMS, ARRAYS = field_list()


def _count(expr, dims):
    return int(eval(expr, {"__builtins__": {}}, dict(dims)))


def pack(model, names=None):
    """model: dict name -> int (dims) / ndarray (arrays); names: optional dict objtype -> [name | None].  Returns bytes."""
    dims = {d: int(model[d]) for d in DIMS}
    out = bytearray()
    out += MAGIC
    out += struct.pack("<i", len(DIMS))
    for d in DIMS:
        out += struct.pack("<i", dims[d])
    for kind, name, cnt in ARRAYS:
        n = _count(cnt, dims)
        while len(out) % 8:
            out += b"\0"
        arr = np.asarray(model[name]).reshape(-1)
        if arr.size != n:
            raise ValueError(f"field {name}: expected {n} values, got {arr.size}")
        if kind == "I":
            out += arr.astype("<i4").tobytes()
        else:
            out += arr.astype("<f8").tobytes()
    while len(out) % 8:
        out += b"\0"
    if names:
        # RGNAMES1, int32 ntypes, then per type: type name, NUL, int32 count, count NUL-terminated names ("" = unnamed)
        out += NAMES_MAGIC
        out += struct.pack("<i", len(names))
        for typ in sorted(names):
            out += typ.encode() + b"\0"
            out += struct.pack("<i", len(names[typ]))
            for n in names[typ]:
                out += (n or "").encode() + b"\0"
    return bytes(out)


def unpack_names(blob):
    """The name tables of a blob (dict objtype -> [name | None]) or None when the blob carries none."""
    blob = bytes(blob)
    k = blob.find(NAMES_MAGIC)
    if k < 0:
        return None
    off = k + len(NAMES_MAGIC)
    (nt,) = struct.unpack_from("<i", blob, off)
    off += 4
    out = {}
    for _ in range(nt):
        e = blob.index(b"\0", off)
        typ = blob[off:e].decode()
        off = e + 1
        (cnt,) = struct.unpack_from("<i", blob, off)
        off += 4
        lst = []
        for _ in range(cnt):
            e = blob.index(b"\0", off)
            lst.append(blob[off:e].decode() or None)
            off = e + 1
        out[typ] = lst
    return out


def unpack(blob):
    """Inverse of pack(): returns dict of ints and (writable) flat numpy arrays."""
    blob = bytes(blob)
    if blob[:8] != MAGIC:
        raise ValueError("bad model blob magic")
    (ndim,) = struct.unpack_from("<i", blob, 8)
    if ndim != len(DIMS):
        raise ValueError("model blob was built against a different rg_model_fields.h")
    vals = struct.unpack_from("<%di" % ndim, blob, 12)
    model = dict(zip(DIMS, vals))
    off = 12 + 4 * ndim
    for kind, name, cnt in ARRAYS:
        n = _count(cnt, model)
        off = (off + 7) & ~7
        if kind == "I":
            model[name] = np.frombuffer(blob, dtype="<i4", count=n, offset=off).copy()
            off += 4 * n
        else:
            model[name] = np.frombuffer(blob, dtype="<f8", count=n, offset=off).copy()
            off += 8 * n
    return model


def pid_stride(m):
    """Floats of controller state per actuator in `userdata` / RG_FIELD_PID: 3 for mujoco-py's PID (integral, last error, last
    derivative); 6 for every actuator of a model that has a cascaded-PI actuator (actuator_user[0] = 1: position-loop integral,
    last error, last derivative, velocity-loop integral, smoothed set-point, "a step has been taken" flag)."""
    import numpy as np

    casc = (np.asarray(m["actuator_user0"]).reshape(-1) == 1.0) & (np.asarray(m["actuator_biastype"]).reshape(-1) == 3)
    return 6 if m["nu"] and bool(casc.any()) else 3
