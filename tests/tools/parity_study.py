"""(Test tooling: compares the engine with the oracle -- lives under tests/ because only tests may load oracle/.)
Free-running parity study (SURVEY.md 8(d) tiers T2/T3) on a B200: the CUDA engine and the fp64 oracle start
from identical states and receive identical action streams for 200 env-steps (2000 substeps); reports the
divergence curve and statistical invariants.  Writes gpurun_out/parity_study.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import control_matrix, oracle_pair  # noqa: E402
from robogym_b200 import modelblob  # noqa: E402
from robogym_b200 import build, engine  # noqa: E402

build.build()
blob = open(os.path.join(ROOT, "robogym_b200/assets/dactyl_locked.rgm"), "rb").read()
names = json.load(open(os.path.join(ROOT, "robogym_b200/assets/dactyl_locked.names.json")))
N = int(os.environ.get("N", 48))
T = int(os.environ.get("T", 200))
model = engine.DeviceModel(blob, 0)
sim = engine.BatchedSim(model, N, 10, outputs=("site_xpos", "ncon", "warn", "contact"))
oms = [oracle_pair(blob) for _ in range(N)]
cr = oms[0][0].field("actuator_ctrlrange").reshape(-1, 2)
nu = len(cr)
cube_site = names["site"].index("cube:center")
rng = np.random.RandomState(7)
ctrl = np.tile(cr.mean(1), (N, 1))
# settle both (identical, deterministic) then perturb each env's cube pose identically on both sides
for k in range(N):
    oms[k][1].ctrl[:] = ctrl[k]
for _ in range(20):
    for om, d in oms:
        d.env_step(10)
    sim.ctrl.copy_(torch.tensor(ctrl, dtype=torch.float32, device="cuda"))
    sim.step()
for k in range(N):
    d = oms[k][1]
    d.qpos[0:3] += rng.normal(0, 0.005, 3)
    # start both sides from the SAME state: copy the oracle state into the engine
sim.qpos.copy_(torch.tensor(np.stack([d.qpos for _, d in oms]), dtype=torch.float32, device="cuda"))
sim.qvel.copy_(torch.tensor(np.stack([d.qvel for _, d in oms]), dtype=torch.float32, device="cuda"))
sim.pid.copy_(torch.tensor(np.stack([d.userdata[:3 * nu] for _, d in oms]), dtype=torch.float32, device="cuda"))
sim.qacc_warmstart.copy_(torch.tensor(np.stack([d.qacc_warmstart for _, d in oms]), dtype=torch.float32, device="cuda"))
live_q = list(range(0, 7)) + list(range(14, 38))
hand_q = list(range(14, 38))
curve_all, curve_hand, palm_gpu, palm_cpu = [], [], [], []
P = control_matrix(modelblob.unpack(blob))
FULL = os.environ.get("ACTIONS", "contract") == "contract"
same_mode = np.ones(N, bool)        # T3 subset: the engine and the oracle have had the same set of touching geom pairs at every env-step so far
same_curve = []
for t in range(T):
    a = rng.uniform(-1, 1, (N, nu))
    if FULL:   # SURVEY 8(d) cfg 2: relative actions about the ORACLE's current pose (both sides get the same ctrl)
        qo_now = np.stack([d.qpos for _, d in oms])
        ctrl = np.clip(qo_now @ P.T + a * (cr[:, 1] - cr[:, 0]) / 2, cr[:, 0], cr[:, 1])
    else:
        ctrl = np.clip(ctrl + 0.3 * a * (cr[:, 1] - cr[:, 0]) / 2, cr[:, 0], cr[:, 1])
    sim.ctrl.copy_(torch.tensor(ctrl, dtype=torch.float32, device="cuda"))
    sim.step()
    for k, (om, d) in enumerate(oms):
        d.ctrl[:] = ctrl[k]
        d.env_step(10)
    q = sim.qpos.cpu().numpy()
    qo = np.stack([d.qpos for _, d in oms])
    err = np.abs(q - qo)
    curve_all.append(err[:, live_q].max(1))
    curve_hand.append(err[:, hand_q].max(1))
    palm_gpu.append(float((sim.site_xpos[:, cube_site, 2] > 0.04).float().mean().item()))
    palm_cpu.append(float(np.mean([d.site_xpos.reshape(-1, 3)[cube_site, 2] > 0.04 for _, d in oms])))
    con = sim.contact.cpu().numpy()
    nc = sim.ncon.cpu().numpy()
    for k, (om, d) in enumerate(oms):
        oc = d.contact.reshape(-1, 24)[:int(d.ncon[0])]
        so = sorted((int(r[20]), int(r[21])) for r in oc if r[0] < 0)         # penetrating pairs only: margin-only contacts carry ~no force
        sg = sorted((int(r[0]), int(r[1])) for r in con[k][:nc[k]] if r[2] < 0)
        if so != sg:
            same_mode[k] = False
    same_curve.append(same_mode.copy())
curve_all, curve_hand = np.array(curve_all), np.array(curve_hand)
steps = [0, 1, 2, 4, 9, 19, 49, 99, T - 1]
out = dict(
    n_envs=N, env_steps=T, substeps=10 * T,
    median_max_abs_dqpos_live={str(s + 1): float(np.median(curve_all[s])) for s in steps},
    median_max_abs_dqpos_hand={str(s + 1): float(np.median(curve_hand[s])) for s in steps},
    frac_envs_within_1e3_live={str(s + 1): float(np.mean(curve_all[s] < 1e-3)) for s in steps},
    frac_envs_within_1e3_hand={str(s + 1): float(np.mean(curve_hand[s] < 1e-3)) for s in steps},
    t3_no_mode_switch_subset={str(s_ + 1): dict(envs=int(same_curve[s_].sum()),
                                                frac_within_1e3_live=(float(np.mean(curve_all[s_][same_curve[s_]] < 1e-3)) if same_curve[s_].any() else None),
                                                median_live=(float(np.median(curve_all[s_][same_curve[s_]])) if same_curve[s_].any() else None)) for s_ in steps},
    actions="SURVEY 8(d) cfg 2 (full-range relative)" if FULL else "ctrl += 0.3*a*half-range",
    on_palm_rate_gpu_end=palm_gpu[-1], on_palm_rate_oracle_end=palm_cpu[-1],
    on_palm_rate_gpu_mean=float(np.mean(palm_gpu)), on_palm_rate_oracle_mean=float(np.mean(palm_cpu)),
    warn_bits=int(sim.warn.max().item()),
    note="identical initial states and action streams; fp32 CUDA vs fp64 oracle; contact-rich chaotic system: per-env trajectories "
         "decorrelate after the first contact-mode switch, population statistics stay matched",
)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_study.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
