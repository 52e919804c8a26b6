"""Engine for robogym_b200.mujoco_py_shim backed by the fp64 CPU oracle -- TEST INFRASTRUCTURE ONLY.
Lets the CPU-only tier drive unmodified robogym envs through the shim (the product engine is CUDA)."""
import numpy as np

from oracle import pyoracle
from robogym_b200 import modelblob


class OracleEngine:
    def __init__(self, cm):
        self.cm = cm
        self.om = pyoracle.OracleModel(cm.blob())
        self.d = pyoracle.OracleData(self.om)
        self.nu = cm.m["nu"]
        self.npid = modelblob.pid_stride(cm.m) * self.nu

    def push_model(self, name, arr):
        self.om.field(name)[:] = np.asarray(arr).reshape(-1)

    def push_state(self, qpos, qvel, ctrl, pid, warm, xfrc):
        d = self.d
        d.qpos[:] = qpos; d.qvel[:] = qvel; d.ctrl[:] = ctrl; d.userdata[:self.npid] = pid
        d.qacc_warmstart[:] = warm; d.xfrc_applied[:] = np.asarray(xfrc).reshape(-1)

    def push_mocap(self, pos, quat):
        self.d.mocap_pos[:] = np.asarray(pos).reshape(-1); self.d.mocap_quat[:] = np.asarray(quat).reshape(-1)

    def run(self, nsub, final_forward):
        for _ in range(nsub):
            self.d.step()
        if final_forward:
            self.d.forward()

    def pull(self):
        d = self.d
        ncon = int(d.ncon[0])
        con = d.contact.reshape(-1, 24)[:ncon]
        contact = np.stack([con[:, 20], con[:, 21], con[:, 0], con[:, 19]], axis=1) if ncon else np.zeros((0, 4))
        xq = d.xquat.copy()
        cv, xp = d.cvel.reshape(-1, 6), d.xpos.reshape(-1, 3)          # [w, v at the world origin] -> velocity of the body frame
        xvel = np.concatenate([cv[:, :3], cv[:, 3:] + np.cross(cv[:, :3], xp)], axis=1)
        return dict(qpos=d.qpos.copy(), qvel=d.qvel.copy(), pid=d.userdata[:self.npid].copy(), warm=d.qacc_warmstart.copy(),
                    site_xpos=d.site_xpos.copy(), body_xpos=d.xpos.copy(), body_xquat=xq, geom_xpos=d.geom_xpos.copy(),
                    act_force=d.actuator_force.copy(), qacc=d.qacc.copy(), ncon=ncon, contact=contact, warn=int(d.warning[0]), body_xvel=xvel,
                    sensordata=d.sensordata[:self.cm.m["nsensordata"]].copy())
