"""Compile the reference's MJCF models into committed model blobs (robogym_b200/assets/*.rgm).

Runs in the build container only: it needs /root/reference for the XML + STL assets and composes the
documents with the reference's own MujocoXML (tools/compose_reference_xml.py).  The GPU box has no
/root/reference, so bench.py / the gpu tests load these blobs.  A blob holds derived data only
(frames, inertias, convex-hull vertices and adjacency, collision pair list, constants of
mj_setConst); no reference source text.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import compose_reference_xml as ref  # noqa: E402
from robogym_b200 import mjcf  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "robogym_b200", "assets")


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("dactyl_locked", ref.locked_xml), ("dactyl_reach", ref.reach_xml), ("rearrange_blocks5", ref.rearrange_blocks_xml), ("rearrange_ycb8", ref.rearrange_ycb_xml)):
        try:
            xml = fn()
        except Exception as e:  # reach needs extra assets
            print(f"skip {name}: {e}")
            continue
        cm = mjcf.compile_mjcf(xml)
        cm.m["opt_pid"][0] = 1  # the sims call enable_pid() right after build (cube_env.py:157-160, ur16e/mujoco/simulation/base.py:29)
        with open(os.path.join(OUT, name + ".rgm"), "wb") as f:
            f.write(cm.blob())
        with open(os.path.join(OUT, name + ".names.json"), "w") as f:
            json.dump(cm.names, f)
        print(name, {k: cm.m[k] for k in ("nq", "nv", "nu", "nbody", "ngeom", "npair", "nmeshvert")}, len(cm.blob()), "bytes")


def full_perpendicular():
    """BASELINE.json configs[2]: the reference itself builds dactyl/full_perpendicular (Rubik's cube, nq170/nv168) on the
    mujoco_py shim (robogym/envs/dactyl/full_perpendicular.py:92-154); the compiled model is taken from that MjSim."""
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    for p in (os.path.join(root, "tests", "stubs"), ref.REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import robogym_b200.mujoco_py_shim as shim

    shim.install()
    from oracle_engine import OracleEngine

    shim.set_engine_factory(OracleEngine)
    try:
        from robogym.envs.dactyl.full_perpendicular import make_simple_env

        env = make_simple_env(starting_seed=0)
        cm = env.mujoco_simulation.mj_sim.model._cm
        name = "dactyl_full_perpendicular"
        with open(os.path.join(OUT, name + ".rgm"), "wb") as f:
            f.write(cm.blob())
        with open(os.path.join(OUT, name + ".names.json"), "w") as f:
            json.dump(cm.names, f)
        print(name, {k: cm.m[k] for k in ("nq", "nv", "nu", "nbody", "ngeom", "npair", "nmeshvert")}, len(cm.blob()), "bytes")
    finally:
        shim.set_engine_factory(None)


if __name__ == "__main__":
    if "--full-only" in sys.argv:
        full_perpendicular()
    else:
        main()
        import subprocess

        # fresh interpreter: the composer above imported robogym under a mujoco_py stub, the env needs the shim
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--full-only"])
