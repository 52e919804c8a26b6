"""BatchedSim look-alike on the CPU EMULATION of the kernel source (tests/emu: the .inl files compiled with -DRG_EMU, fp32) --
TEST INFRASTRUCTURE ONLY.  Lets the CPU tier run host-side classes written for BatchedSim (robogym_b200.rearrange_arm) on the
kernel's own arithmetic before a GPU is involved.  State tensors are torch views of the emulation batch's numpy arrays."""
import numpy as np
import torch

import pyemu
from robogym_b200 import modelblob


class _Model:
    def __init__(self, blob, batch):
        self.blob = bytes(blob)
        self.host = modelblob.unpack(self.blob)
        self.names = modelblob.unpack_names(self.blob)
        self._batch = batch

    def name2id(self, objtype, name):
        try:
            return self.names[objtype].index(name)
        except ValueError:
            raise ValueError(f'No "{objtype}" with name {name} exists.')

    def set_field(self, name, values):
        arr = self.host[name]
        arr[...] = np.asarray(values, dtype=arr.dtype).reshape(arr.shape)
        self._batch.model_field(name, np.int32 if arr.dtype.kind == "i" else np.float32)[:] = np.asarray(arr).reshape(-1)


class EmuGenericSim:
    def __init__(self, blob, nenv, n_substeps, contact_capacity=0, row_capacity=0):
        self.torch = torch
        m = modelblob.unpack(bytes(blob))
        self.e = pyemu.EmuBatch(blob, m, nenv, contact_capacity=contact_capacity, row_capacity=row_capacity)
        self.model = _Model(blob, self.e)
        self.nenv, self.n_substeps = int(nenv), int(n_substeps)
        self.e.qpos[:] = np.asarray(m["qpos0"], np.float32)
        for n in ("qpos", "qvel", "ctrl", "pid", "body_xpos", "body_xquat", "sensordata", "mocap_pos", "mocap_quat"):
            a = getattr(self.e, n)
            setattr(self, n, torch.from_numpy(a) if a is not None else None)
        self.qacc_warmstart = torch.from_numpy(self.e.warm)
        self.warn = torch.from_numpy(self.e.warn)
        if self.mocap_pos is not None:
            ids = sorted((b for b in range(m["nbody"]) if m["body_mocapid"][b] >= 0), key=lambda b: m["body_mocapid"][b])
            self.mocap_pos[:] = torch.tensor(m["body_pos"].reshape(-1, 3)[ids], dtype=torch.float32)
            self.mocap_quat[:] = torch.tensor(m["body_quat"].reshape(-1, 4)[ids], dtype=torch.float32)

    def step(self, n_substeps=None, final_forward=True):
        self.e.step(self.n_substeps if n_substeps is None else int(n_substeps), int(final_forward))

    def forward(self):
        self.e.step(0, 1)
