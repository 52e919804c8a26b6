"""CPU stand-in for robogym_b200.engine.BatchedSim backed by the fp64 oracle -- TEST INFRASTRUCTURE ONLY.
Same attribute / method surface (qpos, qvel, ctrl, pid, qacc_warmstart, time, site_xpos, act_force, ncon, warn,
step(), forward(), reset(mask)), float64 torch tensors on the CPU, one oracle instance per environment."""
import numpy as np
import torch

from oracle import pyoracle


class OracleBatchedSim:
    def __init__(self, blob, nenv, n_substeps=10):
        self.om = pyoracle.OracleModel(blob)
        self.ds = [pyoracle.OracleData(self.om) for _ in range(nenv)]
        self.nenv, self.n_substeps = nenv, n_substeps
        d = self.ds[0]
        self.nu = d.ctrl.shape[0]
        f = dict(dtype=torch.float64)
        self.qpos0 = torch.tensor(np.array(self.om.field("qpos0")), **f)
        self.qpos = self.qpos0.repeat(nenv, 1).contiguous()
        self.qvel = torch.zeros(nenv, d.qvel.shape[0], **f)
        self.ctrl = torch.zeros(nenv, self.nu, **f)
        self.pid = torch.zeros(nenv, 3 * self.nu, **f)
        self.qacc_warmstart = torch.zeros(nenv, d.qvel.shape[0], **f)
        self.time = torch.zeros(nenv, **f)
        self.site_xpos = torch.zeros(nenv, d.site_xpos.size // 3, 3, **f)
        self.act_force = torch.zeros(nenv, self.nu, **f)
        self.ncon = torch.zeros(nenv, dtype=torch.int32)
        self.warn = torch.zeros(nenv, dtype=torch.int32)

    def _run(self, nsub, final_forward, mask=None):
        for e, d in enumerate(self.ds):
            if mask is not None and not bool(mask[e]):
                continue
            d.qpos[:] = self.qpos[e].numpy(); d.qvel[:] = self.qvel[e].numpy(); d.ctrl[:] = self.ctrl[e].numpy()
            d.userdata[:3 * self.nu] = self.pid[e].numpy(); d.qacc_warmstart[:] = self.qacc_warmstart[e].numpy()
            for _ in range(nsub):
                d.step()
            for _ in range(int(final_forward)):
                d.forward()
            self.qpos[e] = torch.from_numpy(d.qpos.copy()); self.qvel[e] = torch.from_numpy(d.qvel.copy())
            self.pid[e] = torch.from_numpy(d.userdata[:3 * self.nu].copy()); self.qacc_warmstart[e] = torch.from_numpy(d.qacc_warmstart.copy())
            self.site_xpos[e] = torch.from_numpy(d.site_xpos.reshape(-1, 3).copy()); self.act_force[e] = torch.from_numpy(d.actuator_force.copy())
            self.ncon[e] = int(d.ncon[0]); self.warn[e] |= int(d.warning[0])
        sel = torch.ones(self.nenv, dtype=torch.bool) if mask is None else mask.bool()
        self.time[sel] += nsub * float(self.om.field("opt_timestep")[0])

    def step(self, n_substeps=None, final_forward=True, mask=None):
        self._run(self.n_substeps if n_substeps is None else n_substeps, final_forward, mask)

    def forward(self, mask=None, count=1):
        self._run(0, count, mask)

    def reset(self, mask=None):
        idx = torch.arange(self.nenv) if mask is None else mask.nonzero().squeeze(1)
        self.qpos[idx] = self.qpos0
        for t in (self.qvel, self.ctrl, self.pid, self.qacc_warmstart, self.time):
            t[idx] = 0
        self.warn[idx] = 0
