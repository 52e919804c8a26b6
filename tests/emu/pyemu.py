"""ctypes front end of the CPU emulation build of the CUDA device code (tests only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "librg_emu.so")
NCON = 32
CON_STRIDE = 24
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = ctypes.CDLL(_LIB)
        L.rge_create_ex.restype = ctypes.c_void_p
        L.rge_create_ex.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.rge_destroy.argtypes = [ctypes.c_void_p]
        for f in ("rge_dbg_size", "rge_scratch_floats", "rge_small_bytes", "rge_ncon", "rge_pidw"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.rge_name2id.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        L.rge_model_field.restype = ctypes.c_void_p
        L.rge_model_field.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
        L.rge_set_const.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 6
        L.rge_set_sensordata.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.rge_set_mocap.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.rge_step.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 18 + [ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


class EmuBatch:
    """Same state layout as robogym_b200.engine.Batch, backed by the emulation library."""

    def __init__(self, blob, dims, nenv, contact_capacity=0, row_capacity=0, dofs_per_contact=0):
        self.h = lib().rge_create_ex(bytes(blob), len(blob), contact_capacity, row_capacity, dofs_per_contact)
        if not self.h:
            raise RuntimeError("rge_create failed")
        self.ncon_cap = lib().rge_ncon(self.h)
        self.nenv = nenv
        self.d = dims
        f = np.float32
        self.qpos = np.zeros((nenv, dims["nq"]), f)
        self.qvel = np.zeros((nenv, dims["nv"]), f)
        self.ctrl = np.zeros((nenv, dims["nu"]), f)
        self.pid = np.zeros((nenv, lib().rge_pidw(self.h) * dims["nu"]), f)
        self.warm = np.zeros((nenv, dims["nv"]), f)
        self.time = np.zeros(nenv, f)
        self.xfrc = None
        self.timestep = None
        self.site_xpos = np.zeros((nenv, dims["nsite"], 3), f)
        self.body_xpos = np.zeros((nenv, dims["nbody"], 3), f)
        self.body_xquat = np.zeros((nenv, dims["nbody"], 4), f)
        self.geom_xpos = np.zeros((nenv, dims["ngeom"], 3), f)
        self.act_force = np.zeros((nenv, dims["nu"]), f)
        self.qacc = np.zeros((nenv, dims["nv"]), f)
        self.contact = np.zeros((nenv, self.ncon_cap, 4), f)
        self.ncon = np.zeros(nenv, np.int32)
        self.warn = np.zeros(nenv, np.int32)
        self.dbg = np.zeros((nenv, lib().rge_dbg_size(self.h)), f)
        self.sensordata = np.zeros((nenv, dims.get("nsensordata", 0)), f)
        self.mocap_pos = np.zeros((nenv, dims.get("nmocap", 0), 3), f) if dims.get("nmocap", 0) else None
        self.mocap_quat = np.zeros((nenv, dims.get("nmocap", 0), 4), f) if dims.get("nmocap", 0) else None

    def model_field(self, name, dtype):
        n = ctypes.c_int()
        p = lib().rge_model_field(self.h, name.encode(), ctypes.byref(n))
        ct = ctypes.c_int32 if dtype == np.int32 else ctypes.c_float
        return np.frombuffer((ct * n.value).from_address(p), dtype=dtype)

    def step(self, nsub, final_forward=1):
        lib().rge_set_mocap(self.h, _p(self.mocap_pos), _p(self.mocap_quat))
        lib().rge_set_sensordata(self.h, _p(self.sensordata) if self.sensordata.size else None)
        lib().rge_step(self.h, self.nenv, _p(self.qpos), _p(self.qvel), _p(self.ctrl), _p(self.pid), _p(self.warm), _p(self.time),
                       _p(self.xfrc), _p(self.timestep), _p(self.site_xpos), _p(self.body_xpos), _p(self.body_xquat),
                       _p(self.geom_xpos), _p(self.act_force), _p(self.qacc), _p(self.contact), _p(self.ncon), _p(self.warn),
                       _p(self.dbg), nsub, final_forward)

    def forward(self):
        self.step(0, 1)

    def set_const(self):
        """mj_setConst of the (edited) model through the kernel code: dict of float32 arrays"""
        d = self.d
        out = dict(dof_invweight0=np.zeros(d["nv"], np.float32), body_invweight0=np.zeros(2 * d["nbody"], np.float32),
                   tendon_invweight0=np.zeros(d["ntendon"], np.float32), tendon_length0=np.zeros(d["ntendon"], np.float32),
                   body_subtreemass=np.zeros(d["nbody"], np.float32), opt_meaninertia=np.zeros(1, np.float32))
        lib().rge_set_const(self.h, *[_p(out[k]) if out[k].size else None for k in ("dof_invweight0", "body_invweight0", "tendon_invweight0", "tendon_length0", "body_subtreemass", "opt_meaninertia")])
        return out

    def dbg_view(self, env=0):
        d = self.d
        nv, nt, nu = d["nv"], d["ntendon"], d["nu"]
        g = self.dbg[env]
        o = 0
        out = {}
        out["M"] = g[o:o + nv * nv].reshape(nv, nv); o += nv * nv
        for k in ("bias", "passive", "qfa", "smooth", "qacc", "qfc"):
            out[k] = g[o:o + nv]; o += nv
        out["tlen"] = g[o:o + nt]; o += nt
        out["alen"] = g[o:o + nu]; o += nu
        out["aforce"] = g[o:o + nu]; o += nu
        out["ncon"], out["nel"], out["niter"], out["warn"] = [int(x) for x in g[o:o + 4]]; o += 4
        out["con"] = g[o:o + self.ncon_cap * CON_STRIDE].reshape(self.ncon_cap, CON_STRIDE); o += self.ncon_cap * CON_STRIDE
        out["tJ"] = g[o:o + nt * nv].reshape(nt, nv)
        return out

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.rge_destroy(self.h)
            self.h = None
