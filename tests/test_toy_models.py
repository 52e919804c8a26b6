"""Engine paths beyond dactyl/locked on a small inline model (free joints, plane contacts, box-box MPR,
geom-derived inertia, affine + motor actuators): oracle sanity, CPU emulation of the kernel vs oracle, and
(gpu) the CUDA engine vs oracle."""
import numpy as np
import pytest

import pyemu
from helpers import oracle_pair
from robogym_b200 import mjcf, modelblob
from toy_models import FREE_BODIES


@pytest.fixture(scope="module")
def toy():
    cm = mjcf.compile_mjcf(FREE_BODIES)
    return cm, cm.blob()


def rollout(blob, n, seed=0):
    om, d = oracle_pair(blob)
    rng = np.random.RandomState(seed)
    states, after = [], []
    for k in range(n):
        d.ctrl[:] = rng.uniform(-1, 1, om.dim("nu"))
        states.append((d.qpos.copy(), d.qvel.copy(), d.ctrl.copy(), np.zeros(3 * om.dim("nu")), d.qacc_warmstart.copy()))
        for _ in range(5):
            d.step()
        d.forward()
        after.append((d.qpos.copy(), d.qvel.copy(), int(d.ncon[0])))
    return states, after, om


def test_oracle_free_bodies_settle_on_the_floor(toy):
    cm, blob = toy
    om, d = oracle_pair(blob)
    assert cm.m["nq"] == 3 * 7 + 2 and cm.m["nv"] == 3 * 6 + 2
    box = cm.name2id("body", "box")
    assert abs(cm.m["body_mass"][box] - 800 * 8 * 0.05 * 0.04 * 0.03) < 1e-12
    for _ in range(700):
        d.step()
    assert d.warning[0] == 0
    z_box, z_ball = d.qpos[2], d.qpos[7 + 2]
    assert 0.029 < z_box < 0.051 and abs(z_ball - 0.04) < 2e-3        # resting on a face / on the sphere radius
    assert np.abs(d.qvel[:12]).max() < 0.05
    assert d.ncon[0] >= 5                                               # 4 box corners + ball (+ brick)


def test_emulated_kernel_matches_oracle_on_free_bodies(toy):
    cm, blob = toy
    states, after, om = rollout(blob, 40)
    dims = {k: cm.m[k] for k in modelblob.DIMS}
    e = pyemu.EmuBatch(blob, dims, len(states))
    for k, st in enumerate(states):
        e.qpos[k], e.qvel[k], e.ctrl[k], e.pid[k], e.warm[k] = st
    e.step(5, 1)
    eq = np.array([np.abs(e.qpos[k] - after[k][0]).max() for k in range(len(states))])
    ev = np.array([np.abs(e.qvel[k] - after[k][1]).max() for k in range(len(states))])
    assert e.warn.max() == 0
    # box-on-box face contact through MPR leaves the contact POINT under-determined (any point of the face overlap):
    # the brick's rocking velocity is the least reproducible quantity here, hence the looser velocity bound
    assert np.median(eq) < 1e-4 and np.median(ev) < 2e-2
    assert np.mean(e.ncon == np.array([a[2] for a in after])) > 0.8


@pytest.mark.gpu
def test_cuda_matches_oracle_on_free_bodies(toy):
    import torch

    from robogym_b200 import build, engine

    build.build()
    cm, blob = toy
    states, after, om = rollout(blob, 40)
    model = engine.DeviceModel(blob, 0)
    sim = engine.BatchedSim(model, len(states), 5, outputs=("ncon", "warn", "body_xpos"))
    f = lambda i: torch.tensor(np.stack([s[i] for s in states]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.qacc_warmstart.copy_(f(4))
    sim.step()
    torch.cuda.synchronize()
    q, v = sim.qpos.cpu().numpy(), sim.qvel.cpu().numpy()
    eq = np.array([np.abs(q[k] - after[k][0]).max() for k in range(len(states))])
    ev = np.array([np.abs(v[k] - after[k][1]).max() for k in range(len(states))])
    assert int(sim.warn.max()) == 0
    # box-on-box face contact through MPR leaves the contact POINT under-determined (any point of the face overlap):
    # the brick's rocking velocity is the least reproducible quantity here, hence the looser velocity bound
    assert np.median(eq) < 1e-4 and np.median(ev) < 2e-2
    # free-joint world positions come back un-shifted
    assert np.abs(sim.body_xpos[:, cm.name2id("body", "box")].cpu().numpy() - q[:, 0:3]).max() < 1e-5


def _equilibrium_penetration(condim, mu=0.9, g=9.81, solref=(0.02, 1.0), solimp=(0.9, 0.95, 0.001, 0.5, 2.0)):
    """Closed form of MuJoCo's soft-contact law for a free sphere at rest on a plane (documentation, 'Computation'):
    reference acceleration aref = -b v - k d(r) r with k = 1 / (dmax^2 timeconst^2 dampratio^2), regulariser
    R = (1 - d) / d * A with A = 1/m for a free body (A = (1 + mu^2)/m for the rows n +- mu t of a friction pyramid, which
    then share R_py = 2 mu^2 R).  At rest the constraint force m g = sum_rows aref / R_row, which gives
    |r| = g (1 - d) / (k d^2) for condim 1 and mu^2 (1 + mu^2) / 2 times that for a 4-row pyramid; d depends on |r|
    through the solimp sigmoid, hence the fixed point."""
    d0, d1, width, mid, power = solimp
    k = 1.0 / (d1 ** 2 * solref[0] ** 2 * solref[1] ** 2)
    scale = 1.0 if condim == 1 else mu * mu * (1.0 + mu * mu) / 2.0
    r = 1e-4
    for _ in range(200):
        x = min(r / width, 1.0)
        y = x ** power / mid ** (power - 1) if x <= mid else 1.0 - (1.0 - x) ** power / (1.0 - mid) ** (power - 1)
        d = d0 + y * (d1 - d0)
        r = scale * g * (1.0 - d) / (k * d * d)
    return r


@pytest.mark.parametrize("condim", [1, 3])
def test_resting_sphere_sits_at_the_closed_form_penetration(condim):
    """Contact impedance, regularisation, pyramid scaling, solver and integrator together reach the analytic fixed
    point: fp64 oracle to 1e-8 m, the fp32 kernel logic (CPU emulation) to 2e-6 m."""
    from toy_models import RESTING_SPHERE

    cm = mjcf.compile_mjcf(RESTING_SPHERE.format(condim=condim))
    blob = cm.blob()
    want = 0.05 - _equilibrium_penetration(condim)
    om, d = oracle_pair(blob)
    for _ in range(4000):
        d.step()
    assert d.ncon[0] == 1 and np.abs(d.qvel).max() < 1e-9
    assert abs(d.qpos[2] - want) < 1e-8, (d.qpos[2], want)
    e = pyemu.EmuBatch(blob, {k: cm.m[k] for k in modelblob.DIMS}, 1)
    e.qpos[0] = cm.m["qpos0"]
    e.step(4000, 1)
    assert int(e.warn[0]) == 0 and int(e.ncon[0]) == 1
    assert abs(float(e.qpos[0, 2]) - want) < 2e-6, (float(e.qpos[0, 2]), want)


def test_dry_friction_stick_and_slip_have_the_closed_form_rates():
    """frictionloss rows (MuJoCo 'Computation': a box constraint |f| <= frictionloss with the same soft regularisation):
    below the threshold the slider creeps at v = u R / b with R = (1 - d0)/d0 / m and b = 2 / (dmax timeconst) -- the
    constraint is soft, not rigid -- and above it accelerates at (u - frictionloss) / m."""
    from toy_models import FRICTION_SLIDER

    cm = mjcf.compile_mjcf(FRICTION_SLIDER)
    blob = cm.blob()
    m_, floss, d0, dmax, tc = 2.0, 1.5, 0.9, 0.95, 0.02
    creep = lambda u: u * ((1 - d0) / d0 / m_) / (2.0 / (dmax * tc))
    dims = {k: cm.m[k] for k in modelblob.DIMS}
    for u, rate_tol in ((0.6, 1e-9), (-1.2, 1e-9)):
        om, d = oracle_pair(blob)
        d.ctrl[0] = u
        for _ in range(2000):
            d.step()
        assert abs(d.qvel[0] - creep(u)) < rate_tol + 1e-6 * abs(creep(u)), (u, d.qvel[0], creep(u))
        e = pyemu.EmuBatch(blob, dims, 1)
        e.ctrl[0, 0] = u
        e.step(2000, 1)
        assert abs(float(e.qvel[0, 0]) - creep(u)) < 1e-5 * abs(creep(u)) + 1e-8
    om, d = oracle_pair(blob)
    d.ctrl[0] = 4.0
    for _ in range(500):
        d.step()
    assert abs(d.qvel[0] - (4.0 - floss) / m_ * 500 * 0.002) < 1e-9
    e = pyemu.EmuBatch(blob, dims, 1)
    e.ctrl[0, 0] = 4.0
    e.step(500, 1)
    assert abs(float(e.qvel[0, 0]) - (4.0 - floss) / m_ * 1.0) < 1e-4


def test_primitive_pairs_have_the_closed_form_penetration():
    """Minkowski Portal Refinement (libccd's algorithm, tolerance 1e-6) on pairs whose answer is known:
    sphere-sphere (r1 + r2 - |c|), capsule-sphere off the axis, sphere pressed into a box face.
    MuJoCo reports dist = -depth along the normal pointing from geom1 to geom2."""
    from toy_models import PRIMITIVE_PAIRS

    cm = mjcf.compile_mjcf(PRIMITIVE_PAIRS)
    blob = cm.blob()
    gid = lambda n: cm.name2id("geom", n)
    want = {
        (gid("s1"), gid("s2")): (-(0.05 + 0.04 - 0.08), (1.0, 0.0, 0.0)),
        (gid("c1"), gid("s3")): (-(0.03 + 0.03 - 0.05), (1.0, 0.0, 0.0)),          # sphere beside the capsule's cylinder part
        (gid("b1"), gid("s4")): (-(0.05 + 0.04 - 0.085), (0.0, 0.0, 1.0)),         # sphere 5 mm into the +z face
    }
    om, d = oracle_pair(blob)
    d.forward()
    # MuJoCo puts the geom of lower type first in a pair (sphere < capsule < box): flip the expected normal accordingly
    canon = lambda recs: {tuple(sorted((int(r[20]), int(r[21])))): (r, 1.0 if int(r[20]) < int(r[21]) else -1.0) for r in recs}
    got = canon(d.contact.reshape(-1, 24)[:int(d.ncon[0])])
    assert set(got) == set(want)
    for key, (dist, n) in want.items():
        r, sgn = got[key]
        assert abs(r[0] - dist) < 2e-6, (key, r[0], dist)
        assert np.abs(r[4:7] - sgn * np.array(n)).max() < 2e-3
    e = pyemu.EmuBatch(blob, {k: cm.m[k] for k in modelblob.DIMS}, 1)
    e.qpos[0] = cm.m["qpos0"]
    e.forward()
    con = e.dbg_view()["con"][:int(e.ncon[0])]
    gote = canon(con)
    assert set(gote) == set(want)
    for key, (dist, n) in want.items():
        r, sgn = gote[key]
        assert abs(r[0] - dist) < 5e-6 and np.abs(r[4:7] - sgn * np.array(n)).max() < 5e-3


def test_cylinder_wrapped_tendon_has_the_closed_form_length_and_moment_arm():
    """Spatial tendon over a cylinder: length = two tangent segments + the arc between the tangent points; its
    derivative w.r.t. the slider (the tendon Jacobian / moment arm) is the cosine between the slide axis and the last
    segment.  Checked for the compile-time numpy geometry, the oracle and the emulated kernel at three slider positions."""
    from toy_models import WRAPPED_TENDON

    cm = mjcf.compile_mjcf(WRAPPED_TENDON)
    blob = cm.blob()
    R, c = 0.03, np.array([0.0, -0.01])

    def analytic(xb):
        p1, p2 = np.array([-0.1, 0.0]) - c, np.array([xb, 0.0]) - c
        d1, d2 = np.linalg.norm(p1), np.linalg.norm(p2)
        phi = np.arccos(np.dot(p1, p2) / (d1 * d2))                    # angle subtended at the axis (over the top)
        arc = phi - np.arccos(R / d1) - np.arccos(R / d2)
        assert arc > 0
        return np.sqrt(d1 * d1 - R * R) + np.sqrt(d2 * d2 - R * R) + R * arc

    dims = {k: cm.m[k] for k in modelblob.DIMS}
    for q in (0.0, 0.05, -0.04):
        xb = 0.1 + q
        want = analytic(xb)
        darm = (analytic(xb + 1e-6) - analytic(xb - 1e-6)) / 2e-6
        L, J = mjcf.tendon_eval(cm.m, np.array([q]))
        assert abs(L[0] - want) < 1e-12 and abs(J[0, 0] - darm) < 1e-6
        om, d = oracle_pair(blob)
        d.qpos[0] = q
        d.forward()
        assert abs(d.ten_length[0] - want) < 1e-12 and abs(d.ten_J[0] - darm) < 1e-6
        e = pyemu.EmuBatch(blob, dims, 1)
        e.qpos[0, 0] = q
        e.forward()
        g = e.dbg_view()
        assert abs(float(g["tlen"][0]) - want) < 1e-6 and abs(float(g["tJ"][0, 0]) - darm) < 1e-5


def test_primitive_inertias_and_torque_free_rotation():
    """Compiler: mass / principal inertia of box, capsule, cylinder, sphere from density (textbook formulas).
    Dynamics: a torque-free asymmetric body keeps its world-frame angular momentum and kinetic energy (gyroscopic bias
    terms + quaternion integration); fp64 oracle and fp32 kernel logic over 0.5 s of tumbling."""
    from toy_models import SPINNING_BRICK

    cm = mjcf.compile_mjcf(SPINNING_BRICK)
    m = cm.m
    bid = lambda n: cm.name2id("body", n)
    I = m["body_inertia"].reshape(-1, 3)
    a, b, c = 0.06, 0.04, 0.02
    mb = 1000 * 8 * a * b * c
    assert abs(m["body_mass"][bid("brick")] - mb) < 1e-12
    assert np.allclose(sorted(I[bid("brick")]), sorted([mb / 3 * (b * b + c * c), mb / 3 * (a * a + c * c), mb / 3 * (a * a + b * b)]), rtol=1e-12)
    r, h = 0.03, 0.08                       # capsule: cylinder of half-length h plus two hemispheres
    mc, ms = 500 * np.pi * r * r * 2 * h, 500 * 4 / 3 * np.pi * r ** 3
    assert abs(m["body_mass"][bid("caps")] - (mc + ms)) < 1e-12
    Izz = 0.5 * mc * r * r + 0.4 * ms * r * r
    Ixx = mc * (r * r / 4 + (2 * h) ** 2 / 12) + ms * (0.4 * r * r + h * h + 0.75 * h * r)
    assert np.allclose(sorted(I[bid("caps")]), sorted([Ixx, Ixx, Izz]), rtol=1e-10)
    mcy = 500 * np.pi * r * r * 2 * h
    assert np.allclose(sorted(I[bid("cyl")]), sorted([mcy * (r * r / 4 + (2 * h) ** 2 / 12)] * 2 + [0.5 * mcy * r * r]), rtol=1e-12)
    msp = 500 * 4 / 3 * np.pi * 0.05 ** 3
    assert np.allclose(I[bid("ball")], 0.4 * msp * 0.05 ** 2, rtol=1e-12)

    blob = cm.blob()
    Ib = np.diag(I[bid("brick")])
    iq = m["body_iquat"].reshape(-1, 4)[bid("brick")]

    def momentum(qpos, qvel):
        R = mjcf.quat2mat(mjcf.quat_mul(qpos[3:7], iq))       # principal axes in the world
        w_world = mjcf.quat2mat(qpos[3:7]) @ qvel[3:6]         # free-joint angular velocity is expressed in the body frame
        return R @ Ib @ R.T @ w_world, 0.5 * w_world @ (R @ Ib @ R.T) @ w_world

    om, d = oracle_pair(blob)
    d.qvel[3:6] = [3.0, 0.2, 5.0]                              # near the unstable middle axis: it tumbles
    q0, v0 = d.qpos.copy(), d.qvel.copy()
    L0, E0 = momentum(d.qpos[:7], d.qvel[:6])
    for _ in range(500):
        d.step()
    L1, E1 = momentum(d.qpos[:7], d.qvel[:6])
    assert np.abs(d.qpos[3:7] - q0[3:7]).max() > 0.3          # it really tumbled
    assert np.abs(L1 - L0).max() < 2e-3 * np.linalg.norm(L0) and abs(E1 - E0) < 2e-3 * E0
    # fp32 kernel logic: the same invariants over the same horizon (tumbling about the middle axis amplifies round-off
    # exponentially, so the trajectories themselves are only compared over the first 50 steps)
    e = pyemu.EmuBatch(blob, {k: m[k] for k in modelblob.DIMS}, 1)
    e.qpos[0], e.qvel[0] = q0, v0
    e.step(50, 1)
    om2, d2 = oracle_pair(blob)
    d2.qpos[:], d2.qvel[:] = q0, v0
    for _ in range(50):
        d2.step()
    assert np.abs(e.qpos[0, :7] - d2.qpos[:7]).max() < 2e-5 and np.abs(e.qvel[0, :6] - d2.qvel[:6]).max() < 2e-4
    e.step(450, 1)
    L2, E2 = momentum(e.qpos[0, :7].astype(float), e.qvel[0, :6].astype(float))
    assert np.abs(L2 - L0).max() < 3e-3 * np.linalg.norm(L0) and abs(E2 - E0) < 3e-3 * E0


@pytest.mark.parametrize("name,qtol,vtol", [("ball_chain", 1e-5, 2e-3), ("contacts", 2e-5, 2e-2)])
def test_emulated_kernel_matches_oracle_on_feature_models(name, qtol, vtol):
    """Teacher-forced 5-substep comparisons along an oracle rollout: fp32 kernel logic vs fp64 oracle on joint / geom /
    contact kinds the dactyl models do not contain (every stage: M, bias, contact counts, state after the step)."""
    from toy_models import MODELS

    cm = mjcf.compile_mjcf(MODELS[name])
    blob, m = cm.blob(), cm.m
    om, d = oracle_pair(blob)
    if name == "ball_chain":
        d.qvel[:] = np.random.RandomState(0).uniform(-3, 3, m["nv"])
    e = pyemu.EmuBatch(blob, {k: m[k] for k in modelblob.DIMS}, 1)
    worst_q = worst_v = 0.0
    same_ncon = 0
    for it in range(20):
        for _ in range(15):
            d.step()
        d.forward()
        e.qpos[0], e.qvel[0], e.warm[0] = d.qpos, d.qvel, d.qacc_warmstart
        e.forward()
        g = e.dbg_view()
        assert np.abs(g["M"].ravel() - d.M).max() < 1e-5 * np.abs(d.M).max()
        assert np.abs(g["bias"] - d.qfrc_bias).max() < 1e-4 * max(1e-6, np.abs(d.qfrc_bias).max())
        e.step(5, 1)
        for _ in range(5):
            d.step()
        d.forward()
        same_ncon += int(e.ncon[0]) == int(d.ncon[0])
        worst_q = max(worst_q, float(np.abs(e.qpos[0] - d.qpos).max()))
        worst_v = max(worst_v, float(np.abs(e.qvel[0] - d.qvel).max()))
        assert int(e.warn[0]) == 0
    assert worst_q < qtol and worst_v < vtol and same_ncon >= 18, (worst_q, worst_v, same_ncon)


def test_emulated_kernel_matches_oracle_on_tendons_and_actuators():
    from toy_models import MODELS

    cm = mjcf.compile_mjcf(MODELS["tendons_actuators"])
    blob, m = cm.blob(), cm.m
    assert (m["ntendon"], m["nu"], m["nwrap"]) == (2, 4, 9)
    om, d = oracle_pair(blob)
    e = pyemu.EmuBatch(blob, {k: m[k] for k in modelblob.DIMS}, 1)
    rng = np.random.RandomState(0)
    for it in range(25):
        d.ctrl[:] = rng.uniform(-1, 1, m["nu"])
        for _ in range(15):
            d.step()
        d.forward()
        e.qpos[0], e.qvel[0], e.ctrl[0], e.warm[0] = d.qpos, d.qvel, d.ctrl, d.qacc_warmstart
        e.forward()
        g = e.dbg_view()
        assert np.abs(g["tlen"] - d.ten_length).max() < 2e-6 and np.abs(g["tJ"].ravel() - d.ten_J).max() < 2e-6
        assert np.abs(g["aforce"] - d.actuator_force).max() < 1e-5 and np.abs(g["smooth"] - d.qfrc_smooth).max() < 2e-4
        e.step(5, 1)
        for _ in range(5):
            d.step()
        assert np.abs(e.qpos[0] - d.qpos).max() < 5e-6 and np.abs(e.qvel[0] - d.qvel).max() < 5e-4
    assert int(e.warn[0]) == 0 and np.abs(d.qpos).max() > 0.3          # the arm really moved


def test_free_floating_chain_conserves_linear_momentum():
    """Internal motion of a floating free-root / ball / hinge chain cannot move its centre of mass: the COM velocity over
    three consecutive windows is the same (oracle, 1e-4 relative), and the fp32 kernel logic follows the oracle."""
    from toy_models import MODELS

    cm = mjcf.compile_mjcf(MODELS["floating_chain"])
    blob, m = cm.blob(), cm.m
    mass = m["body_mass"]
    com = lambda dd: (mass[:, None] * dd.xipos.reshape(-1, 3)).sum(0) / mass.sum()
    om, d = oracle_pair(blob)
    v0 = np.random.RandomState(1).uniform(-2, 2, m["nv"])
    d.qvel[:] = v0
    d.forward()
    pts = [com(d)]
    for n in (100, 100, 300):
        for _ in range(n):
            d.step()
        d.forward()
        pts.append(com(d))
    v = [(pts[1] - pts[0]) / 0.1, (pts[2] - pts[1]) / 0.1, (pts[3] - pts[2]) / 0.3]
    assert np.linalg.norm(v[0]) > 1.0
    assert np.abs(v[1] - v[0]).max() < 1e-4 * np.linalg.norm(v[0]) and np.abs(v[2] - v[0]).max() < 1e-4 * np.linalg.norm(v[0])
    om2, d2 = oracle_pair(blob)
    d2.qvel[:] = v0
    e = pyemu.EmuBatch(blob, {k: m[k] for k in modelblob.DIMS}, 1)
    e.qpos[0], e.qvel[0] = d2.qpos, d2.qvel
    e.step(100, 1)
    for _ in range(100):
        d2.step()
    assert np.abs(e.qpos[0] - d2.qpos).max() < 5e-6 and np.abs(e.qvel[0] - d2.qvel).max() < 1e-4


def test_resting_box_shares_its_weight_between_four_corners():
    """plane-box narrow phase (4 corner contacts) + the same closed form: each of the four condim-1 contacts carries
    m g / 4, so the penetration solves k d(r) r = (g / 4) (1 - d) / d."""
    from toy_models import RESTING_SPHERE

    xml = RESTING_SPHERE.format(condim=1).replace('type="sphere" size="0.05"', 'type="box" size="0.05 0.04 0.05"')
    cm = mjcf.compile_mjcf(xml)
    blob = cm.blob()
    want = 0.05 - _equilibrium_penetration(1, g=9.81 / 4.0)
    om, d = oracle_pair(blob)
    for _ in range(4000):
        d.step()
    assert d.ncon[0] == 4 and np.abs(d.qvel).max() < 1e-9
    assert abs(d.qpos[2] - want) < 1e-8, (d.qpos[2], want)
    e = pyemu.EmuBatch(blob, {k: cm.m[k] for k in modelblob.DIMS}, 1)
    e.qpos[0] = cm.m["qpos0"]
    e.step(4000, 1)
    assert int(e.warn[0]) == 0 and int(e.ncon[0]) == 4
    assert abs(float(e.qpos[0, 2]) - want) < 2e-6


def test_joint_limit_holds_the_weight_at_the_closed_form_violation():
    """Limit rows use the same soft law with A = dof_invweight0 = 1/m: the slider settles |r| below its lower limit."""
    from toy_models import LIMITED_SLIDER

    cm = mjcf.compile_mjcf(LIMITED_SLIDER)
    blob = cm.blob()
    want = -0.1 - _equilibrium_penetration(1)
    om, d = oracle_pair(blob)
    for _ in range(4000):
        d.step()
    assert abs(d.qvel[0]) < 1e-9 and abs(d.qpos[0] - want) < 1e-8, (d.qpos[0], want)
    e = pyemu.EmuBatch(blob, {k: cm.m[k] for k in modelblob.DIMS}, 1)
    e.step(4000, 1)
    assert int(e.warn[0]) == 0 and abs(float(e.qpos[0, 0]) - want) < 2e-6


@pytest.mark.parametrize("seed", range(10))
def test_random_kinematic_trees(seed):
    """Fuzz: random trees (branching, one or two joints per body of every type incl. free roots, random primitive geoms in
    rotated frames, springs, limits, dry friction, armature, position / motor actuators) -- the kernel logic reproduces
    the oracle's mass matrix, bias forces and 5-substep state to fp32 round-off on every one."""
    from toy_models import random_tree_xml

    rng = np.random.RandomState(seed)
    cm = mjcf.compile_mjcf(random_tree_xml(rng, nbody=int(rng.randint(4, 10))))
    blob, m = cm.blob(), cm.m
    om, d = oracle_pair(blob)
    e = pyemu.EmuBatch(blob, {k: m[k] for k in modelblob.DIMS}, 1)
    d.qvel[:] = rng.uniform(-1, 1, m["nv"])
    for it in range(5):
        if m["nu"]:
            d.ctrl[:] = rng.uniform(-1, 1, m["nu"])
        for _ in range(20):
            d.step()
        d.forward()
        e.qpos[0], e.qvel[0], e.warm[0] = d.qpos, d.qvel, d.qacc_warmstart
        if m["nu"]:
            e.ctrl[0] = d.ctrl
        e.forward()
        g = e.dbg_view()
        assert np.abs(g["M"].ravel() - d.M).max() < 2e-6 * np.abs(d.M).max()
        assert np.abs(g["bias"] - d.qfrc_bias).max() < 2e-5 * max(1e-6, np.abs(d.qfrc_bias).max())
        e.step(5, 1)
        for _ in range(5):
            d.step()
        assert np.abs(e.qpos[0] - d.qpos).max() < 5e-6 and np.abs(e.qvel[0] - d.qvel).max() < 2e-4
    assert int(e.warn[0]) == 0


@pytest.mark.parametrize("seed", range(6))
def test_random_piles_of_primitives(seed):
    """Fuzz of the collision + contact solver path: six random primitives (box / capsule / sphere / ellipsoid / cylinder,
    condim 1/3/4/6) dropped on a plane and on each other.  Teacher-forced every 20 substeps along the oracle rollout: same
    contact count, state after 5 substeps equal to round-off once resting; first-touch impacts (penetration ~1e-5 m, where
    the MPR normal is ill-conditioned) may differ by ~1e-3."""
    from toy_models import pile_xml

    rng = np.random.RandomState(seed)
    cm = mjcf.compile_mjcf(pile_xml(rng))
    blob, m = cm.blob(), cm.m
    om, d = oracle_pair(blob)
    e = pyemu.EmuBatch(blob, {k: m[k] for k in modelblob.DIMS}, 1)
    errs, same = [], 0
    for it in range(25):
        for _ in range(20):
            d.step()
        d.forward()
        e.qpos[0], e.qvel[0], e.warm[0] = d.qpos, d.qvel, d.qacc_warmstart
        e.step(5, 1)
        for _ in range(5):
            d.step()
        d.forward()
        same += int(e.ncon[0]) == int(d.ncon[0])
        errs.append(float(np.abs(e.qpos[0] - d.qpos).max()))
    assert int(e.warn[0]) == 0 and same >= 23
    assert np.median(errs) < 2e-6 and max(errs) < 5e-3 and np.median(errs[-8:]) < 1e-6, errs


@pytest.mark.parametrize("seed", range(6))
def test_random_conservative_trees_conserve_energy(seed):
    """Oracle self-consistency on arbitrary trees: with damping, springs, limits, friction and actuators stripped, kinetic
    + potential energy stays put (semi-implicit Euler at dt = 0.2 ms: drift below 0.5 % of the kinetic-energy scale) while
    gravity converts one into the other -- mass matrix, bias forces and gravity terms belong to the same Lagrangian."""
    import re

    from toy_models import random_tree_xml

    rng = np.random.RandomState(seed)
    xml = random_tree_xml(rng, nbody=int(rng.randint(4, 9)))
    for pat in (r' damping="[^"]*"', r' stiffness="[^"]*" springref="[^"]*"', r' limited="true" range="[^"]*"', r' frictionloss="[^"]*"'):
        xml = re.sub(pat, "", xml)
    xml = re.sub(r"<actuator>.*</actuator>", "", xml).replace('timestep="0.002"', 'timestep="0.0002"')
    cm = mjcf.compile_mjcf(xml)
    m = cm.m
    om, d = oracle_pair(cm.blob())
    d.qvel[:] = rng.uniform(-1, 1, m["nv"])

    def energy():
        d.forward()
        M = d.M.reshape(m["nv"], m["nv"])
        ke = 0.5 * d.qvel @ M @ d.qvel
        return ke - m["opt_gravity"][2] * (m["body_mass"] * d.xipos.reshape(-1, 3)[:, 2]).sum(), ke

    e0, k0 = energy()
    for _ in range(1000):
        d.step()
    e1, k1 = energy()
    assert abs(k1 - k0) > 0.05 * max(k0, k1)                    # energy really moved between the two forms
    assert abs(e1 - e0) < 5e-3 * max(k0, k1)


@pytest.mark.parametrize("seed", range(5))
def test_random_convex_hull_piles(seed):
    """Fuzz of the hull path: random lumpy point clouds (12-120 points) compiled to convex hulls, dropped on a box and a
    plane.  The kernel's hill climb on the hull edge graph must find the same support points as the oracle's exhaustive
    search: same contact counts, round-off agreement once resting, compiled mass properties shared by both."""
    from toy_models import mesh_pile

    rng = np.random.RandomState(seed)
    xml, clouds = mesh_pile(rng)
    cm = mjcf.compile_mjcf(xml, asset_loader=lambda p: clouds[p.split("/")[-1]])
    blob, m = cm.blob(), cm.m
    assert m["nmesh"] == 5 and (m["mesh_vertnum"] >= 4).all()
    om, d = oracle_pair(blob)
    e = pyemu.EmuBatch(blob, {k: m[k] for k in modelblob.DIMS}, 1)
    errs, same = [], 0
    for it in range(25):
        for _ in range(20):
            d.step()
        d.forward()
        e.qpos[0], e.qvel[0], e.warm[0] = d.qpos, d.qvel, d.qacc_warmstart
        e.step(5, 1)
        for _ in range(5):
            d.step()
        d.forward()
        same += int(e.ncon[0]) == int(d.ncon[0])
        errs.append(float(np.abs(e.qpos[0] - d.qpos).max()))
    assert int(e.warn[0]) == 0 and same >= 23
    assert np.median(errs) < 2e-6 and max(errs) < 5e-3 and np.median(errs[-8:]) < 1e-6, errs


# ---------------------------------------------------------------------------------------------- elliptic friction cones
# (option cone="elliptic" impratio=10: robogym/assets/xmls/robot/ur16e/base.xml:4, the rearrange scenes)
ELLIPTIC = 'iterations="50" cone="elliptic" impratio="10"'


@pytest.mark.parametrize("condim", [3, 4, 6])
def test_elliptic_cone_resting_sphere_sits_at_the_closed_form_penetration(condim):
    """At rest only the normal row of an elliptic contact carries force, with D = 1/R_normal: the equilibrium is the
    condim-1 closed form whatever the friction dimensions are (a pyramid would sit mu^2 (1 + mu^2) / 2 times deeper)."""
    from toy_models import RESTING_SPHERE

    cm = mjcf.compile_mjcf(RESTING_SPHERE.format(condim=condim).replace('iterations="50"', ELLIPTIC))
    assert cm.m["opt_cone"][0] == 1 and cm.m["opt_impratio"][0] == 10
    blob = cm.blob()
    want = 0.05 - _equilibrium_penetration(1)
    om, d = oracle_pair(blob)
    for _ in range(4000):
        d.step()
    assert d.ncon[0] == 1 and np.abs(d.qvel).max() < 1e-9
    assert abs(d.qpos[2] - want) < 1e-8, (d.qpos[2], want)
    e = pyemu.EmuBatch(blob, {k: cm.m[k] for k in modelblob.DIMS}, 1)
    e.qpos[0] = cm.m["qpos0"]
    e.step(4000, 1)
    assert int(e.warn[0]) == 0 and int(e.ncon[0]) == 1
    assert abs(float(e.qpos[0, 2]) - want) < 2e-6, (float(e.qpos[0, 2]), want)


def test_elliptic_cone_sliding_force_lies_on_the_cone():
    """A box sliding on a plane: every contact is in the middle zone of its cone, where the force is the projection onto the
    cone's boundary -- tangential / normal force = the friction coefficient exactly, opposing the sliding direction."""
    from toy_models import RESTING_SPHERE

    xml = RESTING_SPHERE.format(condim=3).replace('iterations="50"', ELLIPTIC).replace('type="sphere" size="0.05"', 'type="box" size="0.05 0.05 0.05"')
    cm = mjcf.compile_mjcf(xml)
    om, d = oracle_pair(cm.blob())
    for _ in range(2000):
        d.step()
    d.qvel[0] = 1.0
    for _ in range(20):
        d.step()
    nefc = int(d.nefc[0])
    typ, f = d.efc_type[:nefc], d.efc_force[:nefc]
    rows = [i for i in range(nefc) if typ[i] == 4]
    assert len(rows) >= 2 and d.qvel[0] > 0.5      # the sliding box leans on its leading edge
    for i in rows:
        assert f[i] > 0
        assert abs(np.hypot(f[i + 1], f[i + 2]) / f[i] - 0.9) < 1e-9          # on the cone: |f_t| = mu f_n
    # the net friction force opposes the motion (contact frames differ per corner, so sum in world coordinates)
    assert d.qacc[0] < -5.0                        # ~ -mu g (soft cone: a little less)


def _elliptic_pile():
    cm = mjcf.compile_mjcf(FREE_BODIES.replace('iterations="20"', 'iterations="20" cone="elliptic" impratio="3"'))
    return cm, cm.blob()


def test_emulated_kernel_matches_oracle_on_an_elliptic_pile():
    """FREE_BODIES (box, ball, brick with condim 3/4, plane and box-box contacts) with elliptic cones: the kernel logic in CPU
    emulation against the oracle, teacher-forced over a rollout that goes through the impacts and into resting contact
    (all three cone zones, dense middle-zone Hessian blocks)."""
    cm, blob = _elliptic_pile()
    states, after, om = rollout(blob, 60)
    e = pyemu.EmuBatch(blob, {k: cm.m[k] for k in modelblob.DIMS}, len(states))
    for k, st in enumerate(states):
        e.qpos[k], e.qvel[k], e.ctrl[k], e.pid[k], e.warm[k] = st
    e.step(5, 1)
    eq = np.array([np.abs(e.qpos[k] - after[k][0]).max() for k in range(len(states))])
    ev = np.array([np.abs(e.qvel[k] - after[k][1]).max() for k in range(len(states))])
    assert e.warn.max() == 0
    assert np.median(eq) < 1e-5 and np.mean(eq < 1e-3) > 0.85 and np.median(ev) < 2e-2
    assert np.mean(e.ncon == np.array([a[2] for a in after])) > 0.8
    assert max(a[2] for a in after) >= 5


@pytest.mark.gpu
def test_cuda_matches_oracle_on_an_elliptic_pile():
    import torch

    from robogym_b200 import build, engine

    build.build()
    cm, blob = _elliptic_pile()
    states, after, om = rollout(blob, 60)
    model = engine.DeviceModel(blob, 0)
    sim = engine.BatchedSim(model, len(states), 5, outputs=("ncon", "warn"))
    f = lambda i: torch.tensor(np.stack([s[i] for s in states]), dtype=torch.float32, device=sim.device)
    sim.qpos.copy_(f(0)); sim.qvel.copy_(f(1)); sim.ctrl.copy_(f(2)); sim.qacc_warmstart.copy_(f(4))
    sim.step()
    torch.cuda.synchronize()
    q, v = sim.qpos.cpu().numpy(), sim.qvel.cpu().numpy()
    eq = np.array([np.abs(q[k] - after[k][0]).max() for k in range(len(states))])
    ev = np.array([np.abs(v[k] - after[k][1]).max() for k in range(len(states))])
    assert int(sim.warn.max()) == 0
    assert np.median(eq) < 1e-5 and np.mean(eq < 1e-3) > 0.85 and np.median(ev) < 2e-2
    assert np.mean(sim.ncon.cpu().numpy() == np.array([a[2] for a in after])) > 0.8
