"""Compose the reference's MJCF documents with the reference's OWN composer.

Runs only in the build container (needs /root/reference).  It imports
robogym.mujoco.mujoco_xml.MujocoXML under a 10-line `mujoco_py` stub (the real
mujoco-py is not installable here) and repeats, call for call, what
robogym/envs/dactyl/common/cube_env.py:172-218 + robogym/envs/dactyl/locked.py:79-96 (locked) and
robogym/envs/dactyl/reach.py:79-143 (reach) do before `xml.build()`.
The resulting XML strings are then compiled by robogym_b200.mjcf (tools/compile_models.py).
"""
import os
import sys
import tempfile
import types

REF = os.environ.get("ROBOGYM_REFERENCE", "/root/reference")


def _install_stub():
    if "mujoco_py" in sys.modules:
        return
    mp = types.ModuleType("mujoco_py")

    class MjSim:  # noqa
        pass

    class MjSimState:  # noqa
        pass

    mp.MjSim, mp.MjSimState, mp.cymj = MjSim, MjSimState, None
    gen = types.ModuleType("mujoco_py.generated")
    const = types.ModuleType("mujoco_py.generated.const")
    const.JNT_FREE, const.JNT_BALL, const.JNT_SLIDE, const.JNT_HINGE = 0, 1, 2, 3
    gen.const = const
    mp.generated, mp.const = gen, const
    sys.modules.update({"mujoco_py": mp, "mujoco_py.generated": gen, "mujoco_py.generated.const": const})


def mujoco_xml_cls():
    import numpy as np

    if not hasattr(np, "float"):
        np.float = float
    _install_stub()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from robogym.mujoco.mujoco_xml import MujocoXML

    return MujocoXML


def locked_xml():
    import numpy as np

    X = mujoco_xml_cls()
    xml = X()
    xml.add_default_compiler_directive()
    p = "rubik/rubik_locked.xml"
    xml.append(X.parse(p).remove_objects_by_name("annotation:outer_bound").add_name_prefix("cube:")
               .set_named_objects_attr("cube:middle", tag="body", pos=[1.0, 0.87, 0.2])
               .set_named_objects_attr("cube:middle", tag="geom", density=421.0))
    xml.append(X.parse(p).remove_objects_by_name("annotation:outer_bound").add_name_prefix("target:")
               .set_named_objects_attr("target:middle", tag="body", pos=[1.0, 0.87, 0.2])
               .set_objects_attr(tag="geom", group="2", conaffinity="0", contype="0"))
    xml.append(X.parse("floor/basic_floor.xml").set_named_objects_attr("floor", tag="body", pos=[1, 1, 0]))
    xml.append(X.parse("robot/shadowhand/main.xml").add_name_prefix("robot0:")
               .set_objects_attr(tag="size")
               .set_named_objects_attr("robot0:hand_mount", tag="body", pos=[1.0, 1.25, 0.15],
                                       euler=[np.pi / 2, 0, np.pi])
               .remove_objects_by_name("robot0:annotation:outer_bound")
               .remove_objects_by_name("robot0:hand_base"))
    xml.append(X.parse("light/default.xml"))
    return xml.xml_string()


def reach_xml():
    import numpy as np

    X = mujoco_xml_cls()
    xml = X()
    xml.add_default_compiler_directive()
    xml.append(X.parse("floor/basic_floor.xml").set_named_objects_attr("floor", tag="body", pos=[1, 1, 0]))
    xml.append(X.parse("robot/shadowhand/main.xml").add_name_prefix("robot0:")
               .set_named_objects_attr("robot0:hand_mount", tag="body", pos=[1.0, 1.25, 0.15],
                                       euler=[np.pi / 2, 0, np.pi])
               .remove_objects_by_name("robot0:annotation:outer_bound")
               .remove_objects_by_name("robot0:hand_base"))
    xml.append(X.parse("light/default.xml"))
    return xml.xml_string()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "locked"
    sys.stdout.write({"locked": locked_xml, "reach": reach_xml}[which]())


def rearrange_blocks_xml(num_objects=5, object_size=0.0254, mujoco_timestep=0.002):
    """BASELINE.json configs[3] (rearrange/blocks): repeats, call for call, RearrangeSimulationInterface.make_xml +
    ArmSimulationInterface.make_robot_xml for TCP control through the mocap weld (robogym/envs/rearrange/simulation/
    base.py:279-322, robogym/robot/ur16e/mujoco/simulation/base.py:73-110) with make_blocks_and_targets' block / target
    documents (robogym/envs/rearrange/common/utils.py:195-241, 284-291) and the default material
    (robogym/envs/rearrange/materials/default.jsonnet: geom condim 6, margin 5e-5; joint damping 0.01, armature 0.001)."""
    import copy

    X = mujoco_xml_cls()
    xml = (X.parse("robot/ur16e/base.xml")
           .set_objects_attr(tag="option", timestep=mujoco_timestep)
           .set_objects_attr(tag="size", njmax=2000, nconmax=500, nuserdata=2000, nuser_actuator=16)
           .add_default_compiler_directive())
    material = dict(geom=dict(condim="6", margin=0.00005), joint=dict(damping="0.01", armature="0.001"))
    for i in range(num_objects):
        name = f"object{i}"
        obj = X.from_string(f"""
        <mujoco>
          <worldbody>
            <body name="{name}" pos="0.0 0.0 0.0">
              <geom type="box" rgba="0.0 0.0 0.0 0.0" material="block_mat"/>
              <joint name="{name}:joint" type="free"/>
            </body>
          </worldbody>
        </mujoco>
        """).set_objects_attr(tag="geom", size=[object_size] * 3)
        target = (copy.deepcopy(obj).remove_objects_by_tag("joint")
                  .add_name_prefix("target:", exclude_attribs=["material", "mesh", "class"])
                  .set_objects_attr(tag="geom", contype=0, conaffinity=0))
        obj.set_objects_attrs(material)
        xml.append(obj)
        xml.append(target)
    xml.append(X.parse("robot/ur16e/jointspec/ur16e_mocap_class.xml"))
    xml.append(X.parse("robot/ur16e/gripper_actuators.xml"))
    return xml.xml_string()


YCB_SCENE = ("003_cracker_box", "011_banana", "025_mug", "035_power_drill", "048_hammer", "005_tomato_soup_can", "037_scissors", "013_apple")


def rearrange_ycb_xml(mesh_dirs=YCB_SCENE, mujoco_timestep=0.002):
    """BASELINE.json configs[4] (rearrange/ycb, num_objects=8): like rearrange_blocks_xml, with make_mesh_object's documents
    (robogym/envs/rearrange/common/utils.py:250-281) -- one free body per object holding one mesh geom per STL part of the
    object's YCB directory (find_meshes_by_dirname, utils.py:997-1021), shifted so that the combined centre of mass sits at the
    body origin -- plus make_target's copy and the default material.  The reference takes the centre of mass from
    trimesh.util.concatenate(...).center_mass; here it is the volume-weighted mean of the parts' centres of mass
    (mjcf.polyhedron_mass_props on the raw triangles), the same integral for closed, consistently oriented parts.
    One fixed choice of eight objects (the env draws them at random per reset, envs/rearrange/ycb.py:75-95)."""
    import copy
    import glob

    import numpy as np

    from robogym_b200 import mjcf

    X = mujoco_xml_cls()
    xml = (X.parse("robot/ur16e/base.xml")
           .set_objects_attr(tag="option", timestep=mujoco_timestep)
           .set_objects_attr(tag="size", njmax=2000, nconmax=500, nuserdata=2000, nuser_actuator=16)
           .add_default_compiler_directive())
    material = dict(geom=dict(condim="6", margin=0.00005), joint=dict(damping="0.01", armature="0.001"))
    stl_root = os.path.join(REF, "robogym", "assets", "stls")
    for i, d in enumerate(mesh_dirs):
        name = f"object{i}"
        files = sorted(glob.glob(os.path.join(stl_root, "ycb", d, "*.stl")))
        vol, mom = 0.0, np.zeros(3)
        for f in files:
            verts = mjcf.load_stl(f)                              # three vertices per triangle, in order
            v, com, _ = mjcf.polyhedron_mass_props(verts, np.arange(len(verts)).reshape(-1, 3))
            vol += v
            mom += v * np.asarray(com)
        pos = " ".join(map(str, (-mom / vol).tolist()))
        rel = [os.path.relpath(f, stl_root) for f in files]
        assets = "\n".join(f'<mesh file="{f}" name="{name}-{k}" scale="1.0 1.0 1.0" />' for k, f in enumerate(rel))
        geoms = "\n".join(f'<geom type="mesh" mesh="{name}-{k}" pos="{pos}"/>' for k in range(len(rel)))
        obj = X.from_string(f"""
        <mujoco>
          <asset>
            {assets}
          </asset>
          <worldbody>
            <body name="{name}" pos="0.0 0.0 0.0">
              {geoms}
              <joint name="{name}:joint" type="free"/>
            </body>
          </worldbody>
        </mujoco>
        """)
        target = (copy.deepcopy(obj).remove_objects_by_tag("joint")
                  .add_name_prefix("target:", exclude_attribs=["material", "mesh", "class"])
                  .set_objects_attr(tag="geom", contype=0, conaffinity=0))
        obj.set_objects_attrs(material)
        xml.append(obj)
        xml.append(target)
    xml.append(X.parse("robot/ur16e/jointspec/ur16e_mocap_class.xml"))
    xml.append(X.parse("robot/ur16e/gripper_actuators.xml"))
    return xml.xml_string()
