"""Observation noise of the dactyl environments, batched (robogym/wrappers/randomizations.py:310-389, `RandomizeObservationWrapper`;
levels: robogym/envs/dactyl/locked.py:233-238 -- fingertip_pos, hand_angle, cube_pos, cube_quat).

Per episode and key: an additive bias N(0,1) * additive and a multiplicative bias 1 + N(0,1) * multiplicative (both times the
correlated multiplier).  Per step and key: an uncorrelated term N(0,1) * uncorrelated.  Vector keys: noisy = value * mult + (bias +
uncorrelated).  Quaternion keys (one noise value per key): noisy = quat_normalize(quat_mul(value, quat_from_angle_and_axis((bias +
uncorrelated) * 1.96, axis))) with axis ~ U(-1,1)^3.  Output keys are `noisy_<key>`, next to the clean ones, as in the reference.
Draw order per call follows the reference (sorted keys; the quaternion's axis after its uncorrelated term), so a replayed stream of
draws gives the same numbers (tests/test_obs_noise.py)."""

QUAT_NOISE_CORRECTION = 1.96          # randomizations.py:311

LOCKED_LEVELS = {                     # locked.py:233-238
    "fingertip_pos": {"uncorrelated": 0.002, "additive": 0.001},
    "hand_angle": {"additive": 0.1, "uncorrelated": 0.1},
    "cube_pos": {"additive": 0.005, "uncorrelated": 0.001},
    "cube_quat": {"additive": 0.1, "uncorrelated": 0.09},
}


class BatchedObservationNoise:
    def __init__(self, torch, rand, nenv, widths, levels=None, correlated_multiplier=1.0, uncorrelated_multiplier=1.0):
        """`widths`: key -> vector length of the clean observation (quaternion keys count 1); `rand`: object with randn(n, k) /
        uniform(lo, hi, n, k) like locked_env.TorchRand."""
        self.torch, self.rand, self.nenv = torch, rand, nenv
        self.levels = dict(LOCKED_LEVELS if levels is None else levels)
        self.widths = {k: (1 if k.endswith("_quat") else int(widths[k])) for k in self.levels}
        self.cm, self.um = float(correlated_multiplier), float(uncorrelated_multiplier)
        self.additive, self.multiplicative = {}, {}
        self.reset()

    def reset(self, idx=None):
        """new per-episode biases for all environments or for the rows `idx`"""
        n = self.nenv if idx is None else len(idx)
        for key in sorted(self.levels):
            w, lv = self.widths[key], self.levels[key]
            add = self.rand.randn(n, w) * lv.get("additive", 0.0) * self.cm
            mul = 1.0 + self.rand.randn(n, w) * lv.get("multiplicative", 0.0) * self.cm
            if idx is None or key not in self.additive:
                if idx is None:
                    self.additive[key], self.multiplicative[key] = add, mul
                else:
                    raise ValueError("reset(idx) before a full reset")
            else:
                self.additive[key][idx] = add
                self.multiplicative[key][idx] = mul

    def __call__(self, obs):
        """obs: dict of [nenv, ...] tensors -> the same dict plus noisy_<key>"""
        t = self.torch
        out = dict(obs)
        for key in sorted(self.levels):
            w, lv = self.widths[key], self.levels[key]
            bias = self.additive[key] + self.rand.randn(self.nenv, w) * lv.get("uncorrelated", 0.0) * self.um
            src = obs.get("noisy_" + key, obs[key])
            if not key.endswith("_quat"):
                out["noisy_" + key] = src * self.multiplicative[key].to(src.dtype) + bias.to(src.dtype)
            else:
                axis = self.rand.uniform(-1.0, 1.0, self.nenv, 3).to(src.dtype)
                axis = axis / axis.norm(dim=1, keepdim=True)
                ang = (bias * QUAT_NOISE_CORRECTION).to(src.dtype)              # [nenv, 1]
                nq = t.cat([t.cos(ang / 2.0), t.sin(ang / 2.0) * axis], dim=1)
                nq = nq / nq.norm(dim=1, keepdim=True)
                w0, x0, y0, z0 = src.unbind(-1)
                w1, x1, y1, z1 = nq.unbind(-1)
                q = t.stack([w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                             w0 * y1 + y0 * w1 + z0 * x1 - x0 * z1, w0 * z1 + z0 * w1 + x0 * y1 - y0 * x1], dim=-1)
                out["noisy_" + key] = q * t.where(q[:, :1] < 0, -t.ones_like(q[:, :1]), t.ones_like(q[:, :1]))   # quat_normalize: w >= 0
        return out
