"""Contact-based observations and penalties of the rearrange environments, batched.

The reference scans `data.contact[0:ncon]` of one simulation in Python:
* `ArmSimulationInterface.get_gripper_table_contact` (robogym/robot/ur16e/mujoco/simulation/base.py:142-167): does any geom of
  the gripper bodies touch the `table_collision_plane`?  (feeds the table-collision penalty / safety logic)
* `RearrangeSimulationInterface.get_wrist_cam_collisions` (robogym/envs/rearrange/simulation/base.py:562-592): what does the wrist
  camera's collision sphere touch -- the table plane, the robot itself, or an object?
* `RearrangeSimulationInterface.get_object_gripper_contact` (:594-636): per object, is it in contact (dist < cutoff) with the left /
  right finger pad?

Here the same questions are answered for every environment at once from the engine's contact output
(`RG_FIELD_CONTACT`: [nenv, K, 4] rows of geom1, geom2, dist, dim, `RG_FIELD_NCON`), with membership tables built once from the
model's names -- a few tensor comparisons, no host round trip.
"""
import numpy as np

GRIPPER_BODIES = ("robot0:gripper_base", "left_gripper", "left_inner_follower", "left_outer_driver", "right_gripper", "right_inner_follower",
                  "right_outer_driver")      # ArmSimulationInterface.gripper_bodies (simulation/base.py:44-52)


class BatchedRearrangeContacts:
    def __init__(self, sim, num_objects, prefix="robot0:"):
        """`sim`: BatchedSim-like with outputs contact / ncon and a model with name tables."""
        self.sim, self.t = sim, sim.torch
        m, model = sim.model.host, sim.model
        t, dev = self.t, sim.qpos.device
        ng = int(m["ngeom"])
        geom_body = np.asarray(m["geom_bodyid"])
        names = [model.id2name("geom", g) for g in range(ng)]
        table = lambda ids: t.tensor([g in ids for g in range(ng)], dtype=t.bool, device=dev)
        gb = {model.name2id("body", b) for b in GRIPPER_BODIES}
        self.is_gripper = table({g for g in range(ng) if int(geom_body[g]) in gb})
        self.table_plane = model.name2id("geom", "table_collision_plane")
        self.wrist_sphere = model.name2id("geom", prefix + "wrist_cam_collision_sphere")
        self.is_robot = table({g for g in range(ng) if names[g] and names[g].startswith(prefix)})
        self.pads = t.tensor([model.name2id("geom", prefix + "left_contact_v"), model.name2id("geom", prefix + "right_contact_v")], dtype=t.long, device=dev)
        # object id of every geom (-1: not an object)
        obj = np.full(ng, -1)
        for k in range(num_objects):
            b = model.name2id("body", f"object{k}")
            obj[geom_body == b] = k
        self.geom_object = t.tensor(obj, dtype=t.long, device=dev)
        self.num_objects = num_objects

    def _rows(self):
        c, ncon = self.sim.contact, self.sim.ncon
        K = c.shape[1]
        valid = self.t.arange(K, device=c.device).unsqueeze(0) < ncon.unsqueeze(1)
        g1, g2 = c[:, :, 0].long().clamp(min=0), c[:, :, 1].long().clamp(min=0)
        return valid, g1, g2, c[:, :, 2]

    def gripper_table_contact(self):
        """[nenv] bool"""
        valid, g1, g2, _ = self._rows()
        hit = (self.is_gripper[g1] & (g2 == self.table_plane)) | (self.is_gripper[g2] & (g1 == self.table_plane))
        return (hit & valid).any(dim=1)

    def wrist_cam_collisions(self):
        """dict of [nenv] bools: table_collision_plane, robot, object, any"""
        valid, g1, g2, _ = self._rows()
        mine1, mine2 = g1 == self.wrist_sphere, g2 == self.wrist_sphere
        other = self.t.where(mine1, g2, g1)
        touch = (mine1 | mine2) & valid
        tab = touch & (other == self.table_plane)
        rob = touch & ~(other == self.table_plane) & self.is_robot[other]
        objc = touch & ~(other == self.table_plane) & ~self.is_robot[other]
        out = {"table_collision_plane": tab.any(dim=1), "robot": rob.any(dim=1), "object": objc.any(dim=1)}
        out["any"] = out["table_collision_plane"] | out["robot"] | out["object"]
        return out

    def object_gripper_contact(self, dist_cutoff=1.0e-5):
        """[nenv, num_objects, 2] float: 1 where the object touches the left / right finger pad"""
        t = self.t
        valid, g1, g2, dist = self._rows()
        ok = valid & (dist < dist_cutoff)
        out = t.zeros(g1.shape[0], self.num_objects, 2, dtype=self.sim.qpos.dtype, device=g1.device)
        for side in range(2):
            pad = self.pads[side]
            other = t.where(g1 == pad, g2, g1)
            touch = ok & ((g1 == pad) | (g2 == pad))
            o = self.geom_object[other]
            for k in range(self.num_objects):
                out[:, k, side] = (touch & (o == k)).any(dim=1).to(out.dtype)
        return out
