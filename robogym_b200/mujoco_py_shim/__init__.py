"""A `mujoco_py` look-alike backed by the B200 step engine: the drop-in boundary of SURVEY.md 8(b).

robogym imports exactly: `mujoco_py.MjSim` (subclassed in robogym/mujoco/mujoco_xml.py:38-91 WITHOUT
calling `__init__`, so everything is set up in `__new__`), `load_model_from_xml` (:259), `MjSimState`,
`cymj.set_pid_control` & warning/error callbacks (robogym/mujoco/simulation_interface.py:86-88,
robogym/mujoco/warning_buffer.py:15-70), `const` / `generated.const`, `ignore_mujoco_warnings`.

    import robogym_b200.mujoco_py_shim as shim
    shim.install()                      # registers sys.modules["mujoco_py"]
    from robogym.envs.dactyl.locked import make_env     # unchanged reference code

One `MjSim` = one environment of a `BatchedSim` of size 1 (the single-env Python objects of robogym are
1-1 with a sim, README.md:101).  The engine factory is pluggable so the CPU-only test tier can drive the
same shim with the fp64 oracle (tests/stubs/oracle_engine.py); the package itself only ships the CUDA
engine and raises without a GPU.
"""
import sys
import types
from collections import namedtuple

import numpy as np

from .. import mjcf, modelblob
from . import const as _const

MjSimState = namedtuple("MjSimState", "time qpos qvel act udd_state")


class MujocoException(Exception):
    pass


_ENGINE_FACTORY = None


def set_engine_factory(fn):
    """fn(compiled_model: mjcf.CompiledModel) -> engine with the interface of CudaEngine below."""
    global _ENGINE_FACTORY
    _ENGINE_FACTORY = fn


class CudaEngine:
    """batch-of-one BatchedSim behind the single-env mujoco_py API (product engine; needs a GPU)."""

    MODEL_FIELDS_PUSHED = None  # all model arrays are pushed when they change

    def __init__(self, cm):
        from .. import engine

        self.cm = cm
        self.model = engine.DeviceModel(cm.blob(), 0)
        self.sim = engine.BatchedSim(self.model, 1, 1, outputs=("site_xpos", "body_xpos", "body_xquat", "body_xvel", "geom_xpos", "act_force", "qacc", "contact", "ncon", "warn", "sensordata"))
        self.sim.enable_xfrc()

    def push_model(self, name, arr):
        self.model.set_field(name, arr)

    def push_state(self, qpos, qvel, ctrl, pid, warm, xfrc):
        t = self.sim.torch
        f = lambda a: t.as_tensor(np.asarray(a, dtype=np.float32)).to(self.sim.device).reshape(1, -1)
        self.sim.qpos.copy_(f(qpos)); self.sim.qvel.copy_(f(qvel)); self.sim.ctrl.copy_(f(ctrl))
        if np.size(pid) == self.sim.pid.numel():      # a model without room for the PID state in userdata does not use the PID path
            self.sim.pid.copy_(f(pid))
        self.sim.qacc_warmstart.copy_(f(warm))
        self.sim.xfrc_applied.copy_(f(xfrc).reshape(self.sim.xfrc_applied.shape))

    def push_mocap(self, pos, quat):
        if self.sim.mocap_pos is not None:
            t = self.sim.torch
            self.sim.mocap_pos.copy_(t.as_tensor(np.asarray(pos, dtype=np.float32)).to(self.sim.device).reshape(self.sim.mocap_pos.shape))
            self.sim.mocap_quat.copy_(t.as_tensor(np.asarray(quat, dtype=np.float32)).to(self.sim.device).reshape(self.sim.mocap_quat.shape))

    def run(self, nsub, final_forward):
        self.sim.step(nsub, final_forward)

    def pull(self):
        s = self.sim
        g = lambda x: x[0].detach().cpu().numpy().astype(np.float64)
        ncon = int(s.ncon[0].item())
        con = g(s.contact)[:ncon]
        return dict(qpos=g(s.qpos), qvel=g(s.qvel), pid=g(s.pid), warm=g(s.qacc_warmstart), site_xpos=g(s.site_xpos),
                    body_xpos=g(s.body_xpos), body_xquat=g(s.body_xquat), geom_xpos=g(s.geom_xpos), act_force=g(s.act_force),
                    qacc=g(s.qacc), ncon=ncon, contact=con, warn=int(s.warn[0].item()), body_xvel=g(s.body_xvel), sensordata=g(s.sensordata))


def _default_factory(cm):
    return CudaEngine(cm)


# ------------------------------------------------------------------------------------------------
class _Opt:
    FIELDS = dict(timestep="opt_timestep", gravity="opt_gravity", tolerance="opt_tolerance", impratio="opt_impratio",
                  iterations="opt_iterations", mpr_tolerance="opt_mpr_tolerance", mpr_iterations="opt_mpr_iterations", cone="opt_cone",
                  disableflags="opt_disableflags")

    def __init__(self, m):
        object.__setattr__(self, "_m", m)

    def __getattr__(self, k):
        f = _Opt.FIELDS.get(k)
        if f is None:
            raise AttributeError(k)
        a = self._m[f]
        return a if a.size > 1 else a.dtype.type(a[0])

    def __setattr__(self, k, v):
        f = _Opt.FIELDS.get(k)
        if f is None:
            raise AttributeError(k)
        self._m[f][...] = v


_SHAPES = dict(body_pos=3, body_quat=4, body_ipos=3, body_iquat=4, body_inertia=3, body_invweight0=2, jnt_pos=3, jnt_axis=3, jnt_range=2,
               jnt_solref=2, jnt_solimp=5, geom_size=3, geom_pos=3, geom_quat=4, geom_friction=3, geom_solref=2, geom_solimp=5,
               site_pos=3, site_quat=4, site_size=3, mesh_vert=3, mesh_face=3, tendon_range=2, actuator_gainprm=10, actuator_biasprm=10, actuator_ctrlrange=2,
               actuator_forcerange=2, actuator_gear=6, eq_data=7, eq_solref=2, eq_solimp=5, dof_solref=2, dof_solimp=5)
_OBJ = dict(body="body", joint="joint", geom="geom", site="site", tendon="tendon", actuator="actuator", mesh="mesh", sensor="sensor")


class PyMjModel:
    """mjModel view: every array of include/rg_model_fields.h as a WRITABLE float64/int32 numpy array
    (robogym edits them in place); edits are pushed to the engine before the next step."""

    def __init__(self, cm):
        self._cm = cm
        self._m = cm.m
        self._names = cm.names
        self.opt = _Opt(self._m)
        n = self._m
        # rendering-only fields robogym touches (cube_env.py:250-274): harmless dummies
        self.geom_rgba = np.ones((n["ngeom"], 4))
        self.site_rgba = np.ones((n["nsite"], 4))
        self.mat_rgba = np.ones((1, 4))
        self.geom_matid = np.zeros(n["ngeom"], dtype=np.int32)
        self.geom_group = np.zeros(n["ngeom"], dtype=np.int32)
        # cameras / lights / headlight: rendering only (robogym/envs/rearrange/simulation/base.py:187-189,781-803 saves and
        # randomises them); kept as writable arrays with the XML's values so that code runs, never sent to the engine
        import types
        import xml.etree.ElementTree as ET

        cams, lights = [], []
        try:
            root = ET.fromstring(cm.xml) if getattr(cm, "xml", None) else None
        except ET.ParseError:
            root = None
        self.nuser_actuator = 0
        if root is not None:
            cams, lights = list(root.iter("camera")), list(root.iter("light"))
            for sz in root.iter("size"):             # mjModel.nuser_actuator: width of actuator_user (only column 0 is carried)
                self.nuser_actuator = int(sz.get("nuser_actuator", self.nuser_actuator))
        vec = lambda e, k, d: np.array([float(x) for x in e.get(k, d).split()], dtype=np.float64)
        self.cam_fovy = np.array([float(c.get("fovy", "45")) for c in cams], dtype=np.float64)
        self.cam_pos = np.array([vec(c, "pos", "0 0 0") for c in cams], dtype=np.float64).reshape(-1, 3)
        self.cam_quat = np.array([vec(c, "quat", "1 0 0 0") for c in cams], dtype=np.float64).reshape(-1, 4)
        self.camera_names = tuple(c.get("name", "") for c in cams)
        self.light_pos = np.array([vec(l, "pos", "0 0 0") for l in lights], dtype=np.float64).reshape(-1, 3)
        self.light_dir = np.array([vec(l, "dir", "0 0 -1") for l in lights], dtype=np.float64).reshape(-1, 3)
        self.light_castshadow = np.array([l.get("castshadow", "true") == "true" for l in lights], dtype=np.float64)
        self.light_ambient = np.array([vec(l, "ambient", "0 0 0") for l in lights], dtype=np.float64).reshape(-1, 3)
        self.light_diffuse = np.array([vec(l, "diffuse", "0.7 0.7 0.7") for l in lights], dtype=np.float64).reshape(-1, 3)
        self.vis = types.SimpleNamespace(headlight=types.SimpleNamespace(ambient=np.full(3, 0.1), diffuse=np.full(3, 0.4), specular=np.full(3, 0.5)))

    def __getattr__(self, k):
        m = self.__dict__.get("_m")
        if m is None:
            raise AttributeError(k)
        if k in m:
            v = m[k]
            if isinstance(v, np.ndarray):
                w = _SHAPES.get(k)
                return v.reshape(-1, w) if w else v
            return v
        for obj in _OBJ:
            if k == obj + "_names":
                return tuple(x if x is not None else "" for x in self._names[obj])
            if k == obj + "_name2id":
                return lambda name, _o=obj: self._name2id(_o, name)
            if k == obj + "_id2name":
                return lambda i, _o=obj: self._names[_o][i]
        raise AttributeError(k)

    def _name2id(self, obj, name):
        try:
            return self._names[obj].index(name)
        except ValueError:
            raise ValueError(f'No "{obj}" with name {name} exists. Available "{obj}" names = {self._names[obj]}.')

    @property
    def actuator_user(self):
        return self._m["actuator_user0"].reshape(-1, 1)

    def get_joint_qpos_addr(self, name):
        j = self._name2id("joint", name)
        a = int(self._m["jnt_qposadr"][j])
        n = {0: 7, 1: 4, 2: 1, 3: 1}[int(self._m["jnt_type"][j])]
        return a if n == 1 else (a, a + n)

    def get_joint_qvel_addr(self, name):
        j = self._name2id("joint", name)
        a = int(self._m["jnt_dofadr"][j])
        n = {0: 6, 1: 3, 2: 1, 3: 1}[int(self._m["jnt_type"][j])]
        return a if n == 1 else (a, a + n)

    def get_xml(self):
        return self._cm.xml

    def camera_name2id(self, name):
        if name not in self.camera_names:
            raise ValueError(f'No "camera" with name {name} exists.')
        return self.camera_names.index(name)


class _Contact:
    __slots__ = ("geom1", "geom2", "dist", "dim")

    def __init__(self, g1=-1, g2=-1, dist=0.0, dim=0):
        self.geom1, self.geom2, self.dist, self.dim = int(g1), int(g2), float(dist), int(dim)


class PyMjData:
    def __init__(self, model):
        m = model._m
        self._model = model
        self.qpos = m["qpos0"].astype(np.float64).copy()
        self.qvel = np.zeros(m["nv"])
        self.ctrl = np.zeros(m["nu"])
        self.qacc = np.zeros(m["nv"])
        self.qacc_warmstart = np.zeros(m["nv"])
        self.userdata = np.zeros(m["nuserdata"])
        self.xfrc_applied = np.zeros((m["nbody"], 6))
        self.site_xpos = np.zeros((m["nsite"], 3))
        self.body_xpos = np.zeros((m["nbody"], 3))
        self.body_xquat = np.zeros((m["nbody"], 4))
        self.geom_xpos = np.zeros((m["ngeom"], 3))
        self.actuator_force = np.zeros(m["nu"])
        self.sensordata = np.zeros(m["nsensordata"])
        self.body_xvelp = np.zeros((m["nbody"], 3))
        self.body_xvelr = np.zeros((m["nbody"], 3))
        mocap = sorted((int(k), b) for b, k in enumerate(m["body_mocapid"]) if k >= 0)
        self.mocap_pos = np.array([m["body_pos"].reshape(-1, 3)[b] for _, b in mocap], dtype=np.float64).reshape(-1, 3)
        self.mocap_quat = np.array([m["body_quat"].reshape(-1, 4)[b] for _, b in mocap], dtype=np.float64).reshape(-1, 4)
        self.time = 0.0
        self.ncon = 0
        self.contact = [_Contact() for _ in range(32)]

    # accessors robogym uses
    def get_site_xpos(self, name): return self.site_xpos[self._model._name2id("site", name)]
    def get_body_xpos(self, name): return self.body_xpos[self._model._name2id("body", name)]
    def get_body_xquat(self, name): return self.body_xquat[self._model._name2id("body", name)]
    def get_body_xmat(self, name): return mjcf.quat2mat(self.get_body_xquat(name))
    def get_geom_xpos(self, name): return self.geom_xpos[self._model._name2id("geom", name)]
    def get_body_xvelp(self, name): return self.body_xvelp[self._model._name2id("body", name)]
    def get_body_xvelr(self, name): return self.body_xvelr[self._model._name2id("body", name)]

    def _mocapid(self, name):
        k = int(self._model._m["body_mocapid"][self._model._name2id("body", name)])
        if k < 0:
            raise ValueError(f'Body "{name}" is not a mocap body.')
        return k

    def get_mocap_pos(self, name): return self.mocap_pos[self._mocapid(name)]
    def get_mocap_quat(self, name): return self.mocap_quat[self._mocapid(name)]
    def set_mocap_pos(self, name, value): self.mocap_pos[self._mocapid(name)] = value
    def set_mocap_quat(self, name, value): self.mocap_quat[self._mocapid(name)] = value

    def get_joint_qpos(self, name):
        a = self._model.get_joint_qpos_addr(name)
        return self.qpos[a] if isinstance(a, int) else self.qpos[a[0]:a[1]]

    def set_joint_qpos(self, name, value):
        a = self._model.get_joint_qpos_addr(name)
        if isinstance(a, int):
            self.qpos[a] = value
        else:
            self.qpos[a[0]:a[1]] = value

    def get_joint_qvel(self, name):
        a = self._model.get_joint_qvel_addr(name)
        return self.qvel[a] if isinstance(a, int) else self.qvel[a[0]:a[1]]


class MjSim:
    """robogym subclasses this and never calls __init__ (mujoco-py initialises in __cinit__), so all
    state is created in __new__ (robogym/mujoco/mujoco_xml.py:50-55)."""

    def __new__(cls, model, *args, nsubsteps=1, **kwargs):
        self = object.__new__(cls)
        factory = _ENGINE_FACTORY or _default_factory
        self._rg_model = model
        self._rg_data = PyMjData(model)
        self._rg_engine = factory(model._cm)
        self._rg_pushed = {k: np.array(v, copy=True) for k, v in model._m.items() if isinstance(v, np.ndarray)}
        self.nsubsteps = nsubsteps
        self.render_callback = None
        self.udd_state = {}
        return self

    def __init__(self, model, *args, **kwargs):
        pass

    # robogym's subclass overrides these two as properties and calls super().data / super().model
    @property
    def model(self):
        return self._rg_model

    @property
    def data(self):
        return self._rg_data

    # ---- engine plumbing
    def _push(self):
        m = self._rg_model._m
        for k, old in self._rg_pushed.items():
            cur = m[k]
            if cur.shape != old.shape or not np.array_equal(cur, old):
                self._rg_engine.push_model(k, cur)
                self._rg_pushed[k] = np.array(cur, copy=True)
        d = self._rg_data
        nu = m["nu"]
        self._rg_engine.push_state(d.qpos, d.qvel, d.ctrl, d.userdata[:modelblob.pid_stride(m) * nu], d.qacc_warmstart, d.xfrc_applied)
        if len(d.mocap_pos):
            self._rg_engine.push_mocap(d.mocap_pos, d.mocap_quat)

    def _pull(self, nsub):
        d, m = self._rg_data, self._rg_model._m
        out = self._rg_engine.pull()
        d.qpos[:] = out["qpos"]; d.qvel[:] = out["qvel"]; d.qacc[:] = out["qacc"]; d.qacc_warmstart[:] = out["warm"]
        npid = modelblob.pid_stride(m) * m["nu"]
        if len(d.userdata) >= npid:
            d.userdata[:npid] = out["pid"]
        d.site_xpos[:] = out["site_xpos"].reshape(-1, 3); d.body_xpos[:] = out["body_xpos"].reshape(-1, 3)
        d.body_xquat[:] = out["body_xquat"].reshape(-1, 4); d.geom_xpos[:] = out["geom_xpos"].reshape(-1, 3)
        d.actuator_force[:] = out["act_force"]
        if "body_xvel" in out:     # [nbody][6]: angular, then linear velocity of the body frame in world axes
            xv = np.asarray(out["body_xvel"]).reshape(-1, 6)
            d.body_xvelr[:] = xv[:, :3]; d.body_xvelp[:] = xv[:, 3:]
        if "sensordata" in out:
            d.sensordata[:] = out["sensordata"]
        d.ncon = out["ncon"]
        while len(d.contact) < d.ncon:
            d.contact.append(_Contact())
        for i in range(d.ncon):
            c = out["contact"][i]
            d.contact[i] = _Contact(c[0], c[1], c[2], c[3])
        d.time += float(m["opt_timestep"][0]) * nsub
        if out["warn"] and _callbacks["warning"] is not None:
            if out["warn"] & 1:
                _callbacks["warning"](b"Pre-allocated contact buffer is full. Increase nconmax above %d." % 32)
            if out["warn"] & 2:
                _callbacks["warning"](b"Pre-allocated constraint buffer is full. Increase njmax above %d." % 64)
            if out["warn"] & 32:
                _callbacks["warning"](b"A contact touched more degrees of freedom than the engine batch allows and was dropped.")
            if out["warn"] & 16:
                _callbacks["warning"](b"A tendon Jacobian row has more non-zeros than the engine keeps.")
            if out["warn"] & 4:
                _callbacks["warning"](b"Nan, Inf or huge value in QACC at DOF 0. The simulation is unstable. Time = %.4f." % d.time)

    def step(self, with_udd=True):
        self._push()
        self._rg_engine.run(self.nsubsteps, False)
        self._pull(self.nsubsteps)

    def forward(self):
        self._push()
        self._rg_engine.run(0, True)
        self._pull(0)

    def reset(self):
        d, m = self._rg_data, self._rg_model._m
        d.qpos[:] = m["qpos0"]; d.qvel[:] = 0; d.ctrl[:] = 0; d.qacc[:] = 0; d.qacc_warmstart[:] = 0
        d.userdata[:] = 0; d.xfrc_applied[:] = 0; d.time = 0.0; d.ncon = 0
        for b, k in enumerate(m["body_mocapid"]):
            if k >= 0:
                d.mocap_pos[k] = m["body_pos"].reshape(-1, 3)[b]; d.mocap_quat[k] = m["body_quat"].reshape(-1, 4)[b]

    def set_constants(self):
        """mj_setConst after model edits (robogym/mujoco/simulation_interface.py:197-201)."""
        mjcf.set_const(self._rg_model._m)

    def get_state(self):
        d = self._rg_data
        return MjSimState(d.time, d.qpos.copy(), d.qvel.copy(), None, dict(self.udd_state))

    def set_state(self, state):
        d = self._rg_data
        d.time = state.time; d.qpos[:] = state.qpos; d.qvel[:] = state.qvel

    def render(self, *args, **kwargs):
        raise NotImplementedError("rendering is outside the step path (SURVEY.md section 2, row 14)")


def load_model_from_xml(xml_string):
    return PyMjModel(mjcf.compile_mjcf(xml_string))


# ---- cymj: PID switch + callbacks (process-global, like mujoco-py)
_callbacks = dict(warning=None, error=None)


class _Cymj:
    @staticmethod
    def set_pid_control(model, data):
        """robogym/mujoco/simulation_interface.py:86-88.  Zeroes the PID state and enables the PID bias path."""
        m = model._m
        if m["nuserdata"] < modelblob.pid_stride(m) * m["nu"]:
            raise MujocoException("nuserdata is too small for the PID controller state")
        data.userdata[:] = 0
        m["opt_pid"][0] = 1

    @staticmethod
    def set_warning_callback(fn): _callbacks["warning"] = fn
    @staticmethod
    def get_warning_callback(): return _callbacks["warning"]
    @staticmethod
    def set_error_callback(fn): _callbacks["error"] = fn
    @staticmethod
    def wrap_mujoco_warning(): return ignore_mujoco_warnings()


cymj = _Cymj()


class ignore_mujoco_warnings:
    def __enter__(self):
        self._prev = _callbacks["warning"]
        _callbacks["warning"] = None
        return self

    def __exit__(self, *a):
        _callbacks["warning"] = self._prev


class _Functions:
    """The handful of `mujoco_py.functions` the reference's tests call."""

    @staticmethod
    def mju_error(msg):
        if _callbacks["error"] is not None:
            _callbacks["error"](msg.encode() if isinstance(msg, str) else msg)
        else:
            raise MujocoException(msg)

    @staticmethod
    def mju_mat2Quat(res, mat):
        res[:] = mjcf.mat2quat(np.asarray(mat, dtype=float).reshape(3, 3))

    @staticmethod
    def mju_quat2Mat(res, quat):
        res[:] = mjcf.quat2mat(np.asarray(quat, dtype=float)).reshape(-1)


functions = _Functions()
const = _const


def install():
    """Make `import mujoco_py` resolve to this shim (only if the real one is absent)."""
    me = sys.modules[__name__]
    sys.modules.setdefault("mujoco_py", me)
    gen = types.ModuleType("mujoco_py.generated")
    gen.const = _const
    sys.modules.setdefault("mujoco_py.generated", gen)
    sys.modules.setdefault("mujoco_py.generated.const", _const)
    sys.modules.setdefault("mujoco_py.cymj", cymj)
    sys.modules.setdefault("mujoco_py.const", _const)
    sys.modules.setdefault("mujoco_py.functions", functions)
    me.generated = gen
    if not hasattr(np, "float"):       # robogym (2020) uses the alias numpy removed (hand_interface.py:371)
        np.float = float
    return me
