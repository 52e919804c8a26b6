"""ctypes front end of the CPU ORACLE (oracle/rgo_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/rgo_oracle.h).  PARITY UNPINNED vs mujoco-py (not installable) except for the one documented real-MuJoCo
output the reference holds (block height 0.51167315, tests/test_rearrange_reset_pin.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "librgo_oracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("rgo_oracle.c", "rgo_collision.inc", "rgo_constraint.inc", "rgo_oracle.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "rg_model_fields.h"))
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = ctypes.CDLL(_LIB)
        L.rgo_model_load.restype = ctypes.c_void_p
        L.rgo_model_load.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.rgo_model_free.argtypes = [ctypes.c_void_p]
        L.rgo_model_field.restype = ctypes.c_void_p
        L.rgo_model_field.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        L.rgo_model_dim.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.rgo_data_new.restype = ctypes.c_void_p
        L.rgo_data_new.argtypes = [ctypes.c_void_p]
        L.rgo_data_free.argtypes = [ctypes.c_void_p]
        L.rgo_data_field.restype = ctypes.c_void_p
        L.rgo_data_field.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        for f in ("rgo_reset", "rgo_forward", "rgo_step"):
            getattr(L, f).argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.rgo_env_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.rgo_pid_stride.argtypes = [ctypes.c_void_p]
        L.rgo_set_casc_gravcomp.argtypes = [ctypes.c_int]
        L.rgo_tendon_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_void_p] * 3
        _lib = L
    return _lib


def _view(ptr, count, is_int):
    ct = ctypes.c_int32 if is_int else ctypes.c_double
    buf = (ct * count).from_address(ptr)
    return np.frombuffer(buf, dtype=np.int32 if is_int else np.float64)


class OracleModel:
    def __init__(self, blob):
        self._blob = bytes(blob)
        self.ptr = lib().rgo_model_load(self._blob, len(self._blob))
        if not self.ptr:
            raise ValueError("rgo_model_load failed")
        self._cache = {}

    def dim(self, name):
        return lib().rgo_model_dim(self.ptr, name.encode())

    def field(self, name):
        """Writable numpy view aliasing the oracle's model memory."""
        if name not in self._cache:
            n, isint = ctypes.c_int(), ctypes.c_int()
            p = lib().rgo_model_field(self.ptr, name.encode(), ctypes.byref(n), ctypes.byref(isint))
            if not p:
                raise KeyError(name)
            self._cache[name] = _view(p, n.value, isint.value)
        return self._cache[name]

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.rgo_model_free(self.ptr)
            self.ptr = None


class OracleData:
    def __init__(self, model):
        self.model = model
        self.ptr = lib().rgo_data_new(model.ptr)
        self._cache = {}

    def field(self, name):
        if name not in self._cache:
            n, isint = ctypes.c_int(), ctypes.c_int()
            p = lib().rgo_data_field(self.ptr, name.encode(), ctypes.byref(n), ctypes.byref(isint))
            if not p:
                raise KeyError(name)
            self._cache[name] = _view(p, n.value, isint.value)
        return self._cache[name]

    def __getattr__(self, name):
        if name.startswith("_") or name in ("model", "ptr"):
            raise AttributeError(name)
        try:
            return self.field(name)
        except KeyError:
            raise AttributeError(name)

    def reset(self):
        lib().rgo_reset(self.model.ptr, self.ptr)

    def forward(self):
        lib().rgo_forward(self.model.ptr, self.ptr)

    def step(self):
        lib().rgo_step(self.model.ptr, self.ptr)

    def env_step(self, nsub):
        lib().rgo_env_step(self.model.ptr, self.ptr, nsub)

    def tendon_eval(self, qpos):
        m = self.model
        L = np.zeros(m.dim("ntendon"))
        J = np.zeros((m.dim("ntendon"), m.dim("nv")))
        q = np.ascontiguousarray(qpos, dtype=np.float64)
        lib().rgo_tendon_eval(m.ptr, self.ptr, q.ctypes.data, L.ctypes.data, J.ctypes.data)
        return L, J

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.rgo_data_free(self.ptr)
            self.ptr = None
