"""BatchedSim look-alike on the fp64 CPU oracle -- TEST INFRASTRUCTURE ONLY (the product class is robogym_b200.engine.BatchedSim
on CUDA).  Carries what robogym_b200.rearrange_arm needs: state tensors, mocap poses, body frames, step / forward, name lookups
and model edits."""
import numpy as np
import torch

from oracle import pyoracle
from robogym_b200 import modelblob


class _Model:
    def __init__(self, blob):
        self.blob = bytes(blob)
        self.host = modelblob.unpack(self.blob)
        self.names = modelblob.unpack_names(self.blob)
        self.om = pyoracle.OracleModel(self.blob)

    def name2id(self, objtype, name):
        try:
            return self.names[objtype].index(name)
        except ValueError:
            raise ValueError(f'No "{objtype}" with name {name} exists.')

    def id2name(self, objtype, i):
        return self.names[objtype][i]

    def set_field(self, name, values):
        arr = self.host[name]
        arr[...] = np.asarray(values, dtype=arr.dtype).reshape(arr.shape)
        self.om.field(name)[:] = np.asarray(arr).reshape(-1)


class OracleGenericSim:
    def __init__(self, blob, nenv, n_substeps):
        self.torch = torch
        self.model = _Model(blob)
        self.nenv, self.n_substeps = int(nenv), int(n_substeps)
        self.ds = [pyoracle.OracleData(self.model.om) for _ in range(nenv)]
        m = self.model.host
        f = dict(dtype=torch.float64)
        self.qpos = torch.tensor(np.array(m["qpos0"]), **f).repeat(nenv, 1).contiguous()
        self.qvel = torch.zeros(nenv, m["nv"], **f)
        self.ctrl = torch.zeros(nenv, m["nu"], **f)
        self.pid = torch.zeros(nenv, modelblob.pid_stride(m) * m["nu"], **f)
        self.qacc_warmstart = torch.zeros(nenv, m["nv"], **f)
        self.body_xpos = torch.zeros(nenv, m["nbody"], 3, **f)
        self.body_xquat = torch.zeros(nenv, m["nbody"], 4, **f)
        self.sensordata = torch.zeros(nenv, m["nsensordata"], **f)
        self.warn = torch.zeros(nenv, dtype=torch.int32)
        self.mocap_pos = self.mocap_quat = None
        if m["nmocap"]:
            ids = sorted((b for b in range(m["nbody"]) if m["body_mocapid"][b] >= 0), key=lambda b: m["body_mocapid"][b])
            self.mocap_pos = torch.tensor(m["body_pos"].reshape(-1, 3)[ids], **f).repeat(nenv, 1, 1).contiguous()
            self.mocap_quat = torch.tensor(m["body_quat"].reshape(-1, 4)[ids], **f).repeat(nenv, 1, 1).contiguous()

    def _run(self, nsub, nforward):
        w = self.pid.shape[1]
        for e, d in enumerate(self.ds):
            d.qpos[:] = self.qpos[e].numpy(); d.qvel[:] = self.qvel[e].numpy(); d.ctrl[:] = self.ctrl[e].numpy()
            d.userdata[:w] = self.pid[e].numpy(); d.qacc_warmstart[:] = self.qacc_warmstart[e].numpy()
            if self.mocap_pos is not None:
                d.mocap_pos[:] = self.mocap_pos[e].numpy().ravel(); d.mocap_quat[:] = self.mocap_quat[e].numpy().ravel()
            for _ in range(nsub):
                d.step()
            for _ in range(int(nforward)):
                d.forward()
            self.qpos[e] = torch.from_numpy(d.qpos.copy()); self.qvel[e] = torch.from_numpy(d.qvel.copy())
            self.pid[e] = torch.from_numpy(d.userdata[:w].copy()); self.qacc_warmstart[e] = torch.from_numpy(d.qacc_warmstart.copy())
            self.body_xpos[e] = torch.from_numpy(d.xpos.reshape(-1, 3).copy()); self.body_xquat[e] = torch.from_numpy(d.xquat.reshape(-1, 4).copy())
            self.sensordata[e] = torch.from_numpy(d.sensordata[:self.sensordata.shape[1]].copy())
            self.warn[e] |= int(d.warning[0])

    def step(self, n_substeps=None, final_forward=True):
        self._run(self.n_substeps if n_substeps is None else n_substeps, final_forward)

    def forward(self):
        self._run(0, 1)
