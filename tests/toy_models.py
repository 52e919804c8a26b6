"""Small inline MJCF models (no reference assets) that exercise engine paths dactyl/locked does not:
free joints, plane-box / plane-sphere contacts, box-box through MPR, capsule inertia from geoms,
affine (position) actuators, condim 1/3/6, joint springs."""

FREE_BODIES = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.004" iterations="20"/>
  <size nuserdata="0" njmax="200" nconmax="50"/>
  <worldbody>
    <body name="floor" pos="0 0 0"><geom name="floor" type="plane" size="2 2 1" condim="3" friction="0.9 0.005 0.0001"/></body>
    <body name="box" pos="0 0 0.2" euler="0.3 0.2 0.1">
      <joint name="box_free" type="free"/>
      <geom name="box" type="box" size="0.05 0.04 0.03" density="800" condim="3"/>
    </body>
    <body name="ball" pos="0.3 0.1 0.15">
      <joint name="ball_free" type="free"/>
      <geom name="ball" type="sphere" size="0.04" density="600" condim="3"/>
    </body>
    <body name="brick" pos="0.02 0.01 0.45" euler="0.1 0.4 0.7">
      <joint name="brick_free" type="free"/>
      <geom name="brick" type="box" size="0.04 0.03 0.02" density="900" condim="4"/>
    </body>
    <body name="arm" pos="-0.4 0 0.5">
      <joint name="shoulder" type="hinge" axis="0 1 0" damping="0.05" armature="0.001" limited="true" range="-1.5 1.5" stiffness="0.2"/>
      <geom name="upper" type="capsule" fromto="0 0 0 0.2 0 0" size="0.02" density="500" contype="0" conaffinity="0"/>
      <body name="fore" pos="0.2 0 0">
        <joint name="elbow" type="hinge" axis="0 1 0" damping="0.02" frictionloss="0.01"/>
        <geom name="lower" type="capsule" fromto="0 0 0 0.15 0 0" size="0.015" density="500" contype="0" conaffinity="0"/>
        <site name="tip" pos="0.15 0 0"/>
      </body>
    </body>
  </worldbody>
  <actuator>
    <position name="a_shoulder" joint="shoulder" kp="2.0" ctrllimited="true" ctrlrange="-1 1"/>
    <motor name="a_elbow" joint="elbow" gear="0.5" ctrllimited="true" ctrlrange="-1 1"/>
  </actuator>
</mujoco>
"""

# a sphere resting on a plane: the equilibrium penetration of MuJoCo's soft contact model has a closed form
RESTING_SPHERE = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.002" iterations="50" tolerance="1e-12"/>
  <size nuserdata="0" njmax="50" nconmax="10"/>
  <worldbody>
    <body name="floor" pos="0 0 0"><geom name="floor" type="plane" size="2 2 1" condim="{condim}" friction="0.9 0.005 0.0001"/></body>
    <body name="ball" pos="0 0 0.0499">
      <joint name="ball_free" type="free"/>
      <geom name="ball" type="sphere" size="0.05" density="700" condim="{condim}" friction="0.9 0.005 0.0001"/>
    </body>
  </worldbody>
</mujoco>
"""

# a 2 kg slider with dry friction (frictionloss) pushed by a motor: stick / slip have closed forms
FRICTION_SLIDER = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.002" iterations="50" tolerance="1e-12" gravity="0 0 0"/>
  <size nuserdata="0" njmax="50" nconmax="10"/>
  <worldbody>
    <body name="cart" pos="0 0 0">
      <joint name="x" type="slide" axis="1 0 0" frictionloss="1.5"/>
      <geom name="cart" type="box" size="0.1 0.1 0.1" mass="2.0" contype="0" conaffinity="0"/>
    </body>
  </worldbody>
  <actuator>
    <motor name="push" joint="x" gear="1" ctrllimited="true" ctrlrange="-10 10"/>
  </actuator>
</mujoco>
"""

# pairs of primitives held in known relative poses (no joints: everything is welded to the world, contacts are still
# generated between geoms of different bodies) -- penetration depth and normal have closed forms
PRIMITIVE_PAIRS = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.002" gravity="0 0 0"/>
  <size nuserdata="0" njmax="100" nconmax="20"/>
  <worldbody>
    <body name="anchor" pos="0 0 0"><joint name="dummy" type="slide" axis="0 0 1"/><geom name="dummy" type="sphere" size="0.001" pos="5 5 5" mass="1" contype="0" conaffinity="0"/></body>
    <body name="s1" pos="0 0 0"><geom name="s1" type="sphere" size="0.05" condim="1"/></body>
    <body name="s2" pos="0.08 0 0"><joint name="j_s2" type="slide" axis="1 0 0"/><geom name="s2" type="sphere" size="0.04" mass="1" condim="1"/></body>
    <body name="c1" pos="0 1 0"><geom name="c1" type="capsule" size="0.03 0.1" condim="1"/></body>
    <body name="s3" pos="0.05 1 0.02"><joint name="j_s3" type="slide" axis="1 0 0"/><geom name="s3" type="sphere" size="0.03" mass="1" condim="1"/></body>
    <body name="b1" pos="0 2 0"><geom name="b1" type="box" size="0.1 0.08 0.05" condim="1"/></body>
    <body name="s4" pos="0.02 2.01 0.085"><joint name="j_s4" type="slide" axis="0 0 1"/><geom name="s4" type="sphere" size="0.04" mass="1" condim="1"/></body>
  </worldbody>
</mujoco>
"""

# a string from site a over a cylinder (axis z) to site b on a slider: the wrapped length is two tangents plus an arc
WRAPPED_TENDON = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.002" gravity="0 0 0"/>
  <size nuserdata="0" njmax="50" nconmax="10"/>
  <worldbody>
    <site name="a" pos="-0.1 0 0"/>
    <body name="post" pos="0 -0.01 0">
      <geom name="cyl" type="cylinder" size="0.03 0.1" contype="0" conaffinity="0"/>
      <site name="side" pos="0 0.05 0"/>
    </body>
    <body name="cart" pos="0.1 0 0">
      <joint name="x" type="slide" axis="1 0 0"/>
      <geom name="cart" type="sphere" size="0.01" mass="1" contype="0" conaffinity="0"/>
      <site name="b" pos="0 0 0"/>
    </body>
  </worldbody>
  <tendon>
    <spatial name="string"><site site="a"/><geom geom="cyl" sidesite="side"/><site site="b"/></spatial>
  </tendon>
</mujoco>
"""

# torque-free spinning brick (no gravity, no contacts) and primitives whose mass / inertia have textbook formulas
SPINNING_BRICK = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.001" gravity="0 0 0"/>
  <size nuserdata="0" njmax="20" nconmax="5"/>
  <worldbody>
    <body name="brick" pos="0 0 1">
      <joint name="free" type="free"/>
      <geom name="brick" type="box" size="0.06 0.04 0.02" density="1000" contype="0" conaffinity="0"/>
    </body>
    <body name="caps" pos="1 0 1"><joint name="f2" type="free"/><geom name="caps" type="capsule" size="0.03 0.08" density="500" contype="0" conaffinity="0"/></body>
    <body name="cyl" pos="2 0 1"><joint name="f3" type="free"/><geom name="cyl" type="cylinder" size="0.03 0.08" density="500" contype="0" conaffinity="0"/></body>
    <body name="ball" pos="3 0 1"><joint name="f4" type="free"/><geom name="ball" type="sphere" size="0.05" density="500" contype="0" conaffinity="0"/></body>
  </worldbody>
</mujoco>
"""

# more joint / geom / contact kinds than any dactyl model has: ball joints inside a chain, joint springs, a limited slide,
# ellipsoid inertia; capsule / cylinder / ellipsoid / box-on-box / sphere-on-sphere contacts with condim 3, 4 and 6
MODELS = {}
MODELS['ball_chain'] = """
<mujoco><compiler angle="radian" coordinate="local"/><option timestep="0.002"/>
<size nuserdata="0" njmax="100" nconmax="20"/>
<worldbody>
 <body name="a" pos="0 0 1"><joint name="b1" type="ball" damping="0.01"/><geom type="capsule" fromto="0 0 0 0.2 0 0" size="0.02" density="800" contype="0" conaffinity="0"/>
  <body name="b" pos="0.2 0 0"><joint name="h1" type="hinge" axis="0 1 0" damping="0.01" stiffness="0.3" springref="0.2"/><geom type="box" size="0.08 0.02 0.03" pos="0.08 0 0" density="600" contype="0" conaffinity="0"/>
   <body name="c" pos="0.16 0 0"><joint name="b2" type="ball"/><joint name="s1" type="slide" axis="1 0 0" damping="0.5" limited="true" range="-0.02 0.05"/><geom type="ellipsoid" size="0.04 0.02 0.03" density="700" contype="0" conaffinity="0"/></body>
  </body>
 </body>
</worldbody></mujoco>"""
MODELS['contacts'] = """
<mujoco><compiler angle="radian" coordinate="local"/><option timestep="0.002"/>
<size nuserdata="0" njmax="200" nconmax="40"/>
<worldbody>
 <body name="floor" pos="0 0 0"><geom name="floor" type="plane" size="2 2 1" condim="3"/></body>
 <body name="cap" pos="0 0 0.06" euler="0 1.4 0.3"><joint type="free"/><geom type="capsule" size="0.03 0.08" density="600" condim="4"/></body>
 <body name="cyl" pos="0.3 0 0.05" euler="1.5 0.1 0"><joint type="free"/><geom type="cylinder" size="0.04 0.06" density="600" condim="3"/></body>
 <body name="ell" pos="-0.3 0 0.05"><joint type="free"/><geom type="ellipsoid" size="0.05 0.03 0.04" density="600" condim="6"/></body>
 <body name="bx" pos="0 0.3 0.04"><joint type="free"/><geom type="box" size="0.05 0.04 0.03" density="600" condim="3"/></body>
 <body name="bx2" pos="0.02 0.31 0.105" euler="0 0 0.4"><joint type="free"/><geom type="box" size="0.03 0.03 0.03" density="600" condim="3"/></body>
 <body name="sp" pos="0.01 -0.3 0.04"><joint type="free"/><geom type="sphere" size="0.04" density="600" condim="3"/></body>
 <body name="sp2" pos="0.02 -0.29 0.115"><joint type="free"/><geom type="sphere" size="0.035" density="600" condim="3"/></body>
</worldbody></mujoco>"""

# tendons and actuators the dactyl models do not have: a spatial tendon over a wrapping sphere with a pulley, tendon spring /
# damper / limits, a fixed tendon with limits, position / velocity / general actuators, a force-limited motor on a tendon
MODELS["tendons_actuators"] = """
<mujoco><compiler angle="radian" coordinate="local"/><option timestep="0.002"/>
<size nuserdata="0" njmax="100" nconmax="20"/>
<worldbody>
 <site name="anchor" pos="0 0 1.2"/>
 <body name="l1" pos="0 0 1"><joint name="j1" type="hinge" axis="0 1 0" damping="0.05" armature="0.002" frictionloss="0.02" limited="true" range="-1.2 1.2"/>
   <geom type="capsule" fromto="0 0 0 0.25 0 0" size="0.02" density="700" contype="0" conaffinity="0"/><site name="s1" pos="0.1 0 0.03"/>
   <geom name="wrapsph" type="sphere" size="0.035" pos="0.25 0 0" contype="0" conaffinity="0" density="1"/><site name="side1" pos="0.25 0 0.08"/>
  <body name="l2" pos="0.25 0 0"><joint name="j2" type="hinge" axis="0 1 0" damping="0.03" limited="true" range="-0.3 1.9"/>
    <geom type="capsule" fromto="0 0 0 0.2 0 0" size="0.018" density="700" contype="0" conaffinity="0"/><site name="s2" pos="0.12 0 0.025"/>
   <body name="l3" pos="0.2 0 0"><joint name="j3" type="hinge" axis="0 1 0" damping="0.02"/><joint name="j3s" type="slide" axis="1 0 0" damping="1.0" stiffness="50"/>
     <geom type="box" size="0.05 0.02 0.02" pos="0.05 0 0" density="700" contype="0" conaffinity="0"/><site name="s3" pos="0.05 0 0.02"/></body>
  </body>
 </body>
</worldbody>
<tendon>
 <fixed name="couple" limited="true" range="-0.5 0.8"><joint joint="j2" coef="1"/><joint joint="j3" coef="0.7"/></fixed>
 <spatial name="flexor" stiffness="20" damping="0.5" limited="true" range="0 0.7"><site site="anchor"/><site site="s1"/><geom geom="wrapsph" sidesite="side1"/><site site="s2"/><pulley divisor="2"/><site site="s2"/><site site="s3"/></spatial>
</tendon>
<actuator>
 <position name="p1" joint="j1" kp="8" ctrllimited="true" ctrlrange="-1 1"/>
 <velocity name="v2" joint="j2" kv="0.5" ctrllimited="true" ctrlrange="-2 2"/>
 <motor name="tm" tendon="flexor" gear="3" ctrllimited="true" ctrlrange="-1 1" forcelimited="true" forcerange="-2 2"/>
 <general name="g3" joint="j3" gainprm="2 0 0" biasprm="0.1 -1.5 -0.05" ctrllimited="true" ctrlrange="-1 1"/>
</actuator>
</mujoco>"""

# a free-floating articulated body (free root, ball shoulder, hinge elbow) in zero gravity: its centre of mass must move uniformly
MODELS["floating_chain"] = """
<mujoco><compiler angle="radian" coordinate="local"/><option timestep="0.001" gravity="0 0 0"/>
<size nuserdata="0" njmax="50" nconmax="10"/>
<worldbody>
 <body name="base" pos="0 0 1"><joint name="root" type="free"/>
   <geom type="box" size="0.08 0.05 0.03" density="900" contype="0" conaffinity="0"/>
   <body name="arm" pos="0.08 0 0"><joint name="sh" type="ball"/>
     <geom type="capsule" fromto="0 0 0 0.15 0 0" size="0.02" density="900" contype="0" conaffinity="0"/>
     <body name="tip" pos="0.15 0 0"><joint name="el" type="hinge" axis="0 0 1"/>
       <geom type="box" size="0.05 0.015 0.02" pos="0.05 0 0" density="900" contype="0" conaffinity="0"/></body>
   </body>
 </body>
</worldbody></mujoco>"""

# a 1.5 kg slider hanging on its lower joint limit under gravity
LIMITED_SLIDER = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.002" iterations="50" tolerance="1e-12"/>
  <size nuserdata="0" njmax="20" nconmax="5"/>
  <worldbody>
    <body name="w" pos="0 0 1">
      <joint name="z" type="slide" axis="0 0 1" limited="true" range="-0.1 0.3"/>
      <geom type="sphere" size="0.03" mass="1.5" contype="0" conaffinity="0"/>
    </body>
  </worldbody>
</mujoco>
"""


import numpy as np  # noqa: E402


def random_tree_xml(rng, nbody=8):
    """random kinematic tree: every body hangs on 1-2 joints of random type, random primitive, random spring/damping/limits"""
    bodies = {0: []}
    parent = {}
    for b in range(1, nbody + 1):
        p = int(rng.randint(0, b)); parent[b] = p; bodies.setdefault(p, []).append(b); bodies.setdefault(b, [])
    jn = [0]
    acts = []
    def body_xml(b, depth):
        pos = rng.uniform(-0.15, 0.15, 3); pos[2] = abs(pos[2]) + (1.0 if parent[b] == 0 else 0.05)
        s = '%s<body name="b%d" pos="%.4f %.4f %.4f">\n' % ('  ' * depth, b, *pos)
        kinds = rng.choice(['hinge', 'slide', 'ball', 'hinge+slide', 'hinge+hinge', 'free'] if parent[b] == 0 else ['hinge', 'slide', 'ball', 'hinge+slide', 'hinge+hinge'])
        for k in kinds.split('+'):
            jn[0] += 1
            name = 'j%d' % jn[0]
            ax = rng.randn(3); ax /= np.linalg.norm(ax)
            extra = ''
            if k in ('hinge', 'slide'):
                extra = ' axis="%.4f %.4f %.4f" damping="%.3f" armature="%.4f"' % (*ax, rng.uniform(0.01, 0.2), rng.uniform(0, 0.01))
                if rng.rand() < 0.4: extra += ' stiffness="%.3f" springref="%.3f"' % (rng.uniform(0.5, 5), rng.uniform(-0.2, 0.2))
                if rng.rand() < 0.4: extra += ' limited="true" range="%.3f %.3f"' % ((-0.3, 0.4) if k == 'hinge' else (-0.03, 0.05))
                if rng.rand() < 0.3: extra += ' frictionloss="%.3f"' % rng.uniform(0.005, 0.05)
                if rng.rand() < 0.5: acts.append('<%s joint="%s" %s ctrllimited="true" ctrlrange="-1 1"/>' % (('position', name, 'kp="%.2f"' % rng.uniform(1, 5)) if rng.rand() < 0.5 else ('motor', name, 'gear="%.2f"' % rng.uniform(0.2, 1))))
            elif k == 'ball':
                extra = ' damping="%.3f"' % rng.uniform(0.01, 0.1)
            s += '%s  <joint name="%s" type="%s" pos="%.3f %.3f %.3f"%s/>\n' % ('  ' * depth, name, k, *(rng.uniform(-0.02, 0.02, 3) if k != 'free' else (0, 0, 0)), extra)
        g = rng.choice(['box', 'capsule', 'sphere', 'ellipsoid', 'cylinder'])
        size = {'box': '%.3f %.3f %.3f' % tuple(rng.uniform(0.02, 0.06, 3)), 'capsule': '%.3f %.3f' % tuple(rng.uniform(0.015, 0.05, 2)),
                'sphere': '%.3f' % rng.uniform(0.02, 0.05), 'ellipsoid': '%.3f %.3f %.3f' % tuple(rng.uniform(0.02, 0.06, 3)), 'cylinder': '%.3f %.3f' % tuple(rng.uniform(0.015, 0.05, 2))}[g]
        s += '%s  <geom type="%s" size="%s" pos="%.3f %.3f %.3f" euler="%.2f %.2f %.2f" density="%.0f" contype="0" conaffinity="0"/>\n' % ('  ' * depth, g, size, *rng.uniform(-0.03, 0.03, 3), *rng.uniform(-1, 1, 3), rng.uniform(300, 1500))
        for c in bodies[b]: s += body_xml(c, depth + 1)
        return s + '%s</body>\n' % ('  ' * depth)
    wb = ''.join(body_xml(c, 2) for c in bodies[0])
    return '<mujoco><compiler angle="radian" coordinate="local"/><option timestep="0.002"/><size nuserdata="0" njmax="200" nconmax="10"/>\n<worldbody>\n%s</worldbody>\n<actuator>%s</actuator></mujoco>' % (wb, ''.join(acts))


def pile_xml(rng, n=6):
    s=''
    for b in range(n):
        g = rng.choice(['box','capsule','sphere','ellipsoid','cylinder'])
        size = {'box': '%.3f %.3f %.3f' % tuple(rng.uniform(0.03, 0.06, 3)), 'capsule': '%.3f %.3f' % tuple(rng.uniform(0.02, 0.05, 2)),
                'sphere': '%.3f' % rng.uniform(0.03, 0.05), 'ellipsoid': '%.3f %.3f %.3f' % tuple(rng.uniform(0.03, 0.06, 3)), 'cylinder': '%.3f %.3f' % tuple(rng.uniform(0.02, 0.05, 2))}[g]
        pos = (rng.uniform(-0.06,0.06), rng.uniform(-0.06,0.06), 0.08+0.11*b)
        s += '<body name="o%d" pos="%.3f %.3f %.3f" euler="%.2f %.2f %.2f"><joint type="free"/><geom type="%s" size="%s" density="%.0f" condim="%d" friction="%.2f 0.005 0.0001"/></body>\n' % (b, *pos, *rng.uniform(-1.5,1.5,3), g, size, rng.uniform(400,1200), int(rng.choice([1,3,4,6])), rng.uniform(0.3,1.0))
    return '<mujoco><compiler angle="radian" coordinate="local"/><option timestep="0.002"/><size nuserdata="0" njmax="400" nconmax="60"/><worldbody><body name="floor" pos="0 0 0"><geom name="floor" type="plane" size="2 2 1" condim="3"/></body>\n%s</worldbody></mujoco>' % s


def mesh_pile(rng, n=5):
    clouds={}
    assets=''; bodies=''
    for b in range(n):
        npts=int(rng.randint(12,120))
        pts=rng.randn(npts,3); pts/=np.linalg.norm(pts,axis=1,keepdims=True)
        pts*=rng.uniform(0.03,0.06,3)*rng.uniform(0.7,1.0,(npts,1))     # lumpy ellipsoid-ish hull
        clouds['m%d.stl'%b]=pts
        assets+='<mesh name="m%d" file="m%d.stl"/>'%(b,b)
        pos=(rng.uniform(-0.05,0.05), rng.uniform(-0.05,0.05), 0.08+0.12*b)
        bodies+='<body name="o%d" pos="%.3f %.3f %.3f" euler="%.2f %.2f %.2f"><joint type="free"/><geom type="mesh" mesh="m%d" density="%.0f" condim="%d"/></body>\n'%(b,*pos,*rng.uniform(-1.5,1.5,3),b,rng.uniform(400,1200),int(rng.choice([1,3,4])))
    xml='<mujoco><compiler angle="radian" coordinate="local"/><option timestep="0.002"/><size nuserdata="0" njmax="400" nconmax="60"/><asset>%s</asset><worldbody><body name="floor" pos="0 0 0"><geom name="floor" type="plane" size="2 2 1" condim="3"/></body>\n%s<body name="blk" pos="0 0 0.03"><geom type="box" size="0.08 0.08 0.03" condim="3"/></body></worldbody></mujoco>'%(assets,bodies)
    return xml, clouds

# A 3-link arm whose tip is dragged by a mocap body through a weld (the UR16e tool-centre-point scheme of
# robogym/assets/xmls/robot/ur16e/base.xml:52-54 + tcp_mocap.xml), a two-finger gripper whose fingers are coupled by a joint
# equality with a direct (negative) solref (gripper_actuators.xml:2-4), and a free brick hanging from a second mocap body.
MOCAP_ARM = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.002" iterations="30" tolerance="1e-10"/>
  <size nuserdata="0" njmax="200" nconmax="20"/>
  <worldbody>
    <body name="mocap" mocap="true" pos="0.45 0.0 0.55">
      <geom name="mocap_viz" type="box" size="0.005 0.005 0.005" contype="0" conaffinity="0"/>
    </body>
    <body name="hook" mocap="true" pos="-0.3 0.2 0.6"/>
    <body name="floor" pos="0 0 0"><geom name="floor" type="plane" size="2 2 1" condim="3"/></body>
    <body name="base" pos="0 0 0.5">
      <joint name="yaw" type="hinge" axis="0 0 1" damping="0.5" armature="0.01"/>
      <geom name="g_base" type="capsule" fromto="0 0 0 0.2 0 0" size="0.025" density="600" contype="0" conaffinity="0"/>
      <body name="upper" pos="0.2 0 0">
        <joint name="shoulder" type="hinge" axis="0 1 0" damping="0.5" armature="0.01" limited="true" range="-2 2"/>
        <geom name="g_upper" type="capsule" fromto="0 0 0 0.15 0 0" size="0.02" density="600" contype="0" conaffinity="0"/>
        <body name="fore" pos="0.15 0 0">
          <joint name="elbow" type="hinge" axis="0 1 0" damping="0.3" armature="0.01"/>
          <joint name="roll" type="hinge" axis="1 0 0" damping="0.1" armature="0.005"/>
          <geom name="g_fore" type="capsule" fromto="0 0 0 0.1 0 0" size="0.015" density="600" contype="0" conaffinity="0"/>
          <body name="tcp" pos="0.1 0 0">
            <geom name="g_palm" type="box" size="0.01 0.03 0.01" density="600" contype="0" conaffinity="0"/>
            <body name="finger_r" pos="0.02 -0.02 0">
              <joint name="r_slide" type="slide" axis="0 1 0" damping="2" armature="0.001" limited="true" range="-0.005 0.02"/>
              <geom name="g_fr" type="box" size="0.015 0.003 0.008" density="600" contype="0" conaffinity="0"/>
            </body>
            <body name="finger_l" pos="0.02 0.02 0">
              <joint name="l_slide" type="slide" axis="0 -1 0" damping="2" armature="0.001"/>
              <geom name="g_fl" type="box" size="0.015 0.003 0.008" density="600" contype="0" conaffinity="0"/>
            </body>
          </body>
        </body>
      </body>
    </body>
    <body name="brick" pos="-0.3 0.2 0.6">
      <joint name="brick_free" type="free"/>
      <geom name="brick" type="box" size="0.03 0.02 0.01" density="900" condim="3"/>
    </body>
  </worldbody>
  <equality>
    <weld name="mocap_weld" body1="mocap" body2="tcp" solimp="0.9 0.95 0.001" solref="0.02 1"/>
    <weld name="hook_weld" body1="hook" body2="brick"/>
    <joint name="coupling" joint1="r_slide" joint2="l_slide" polycoef="0 1 0 0 0" solref="-50000 -100"/>
  </equality>
  <actuator>
    <position name="a_grip" joint="r_slide" kp="200" ctrllimited="true" ctrlrange="0 0.02"/>
  </actuator>
</mujoco>
"""


# A 3-joint arm driven the way the UR16e's default calibration drives its joints: mujoco-py's cascaded-PI user controller
# (user="1", 10 gain parameters; robogym/assets/xmls/robot/ur16e/jointspec/calibrations/cascaded_pi/joint_actuations.xml:4-10)
# on light joints (cascaded_pi/ur16e_ik_class.xml: damping / armature / frictionloss 0.01), plus a plain-PID gripper actuator
# (gripper_actuators.xml:7) in the same model, so both user controllers share one userdata block.
CASCADED_ARM = """
<mujoco>
  <compiler angle="radian" coordinate="local"/>
  <option timestep="0.001" iterations="30" tolerance="1e-10"/>
  <size nuserdata="100" njmax="200" nconmax="20" nuser_actuator="16"/>
  <worldbody>
    <body name="base" pos="0 0 0.5">
      <joint name="J1" type="hinge" axis="0 0 1" damping="0.01" armature="0.01" frictionloss="0.01"/>
      <geom name="g_base" type="capsule" fromto="0 0 0 0.3 0 0" size="0.04" density="800" contype="0" conaffinity="0"/>
      <body name="upper" pos="0.3 0 0">
        <joint name="J2" type="hinge" axis="0 1 0" damping="0.01" armature="0.01" frictionloss="0.01" limited="true" range="-3 3"/>
        <geom name="g_upper" type="capsule" fromto="0 0 0 0.25 0 0" size="0.03" density="800" contype="0" conaffinity="0"/>
        <body name="fore" pos="0.25 0 0">
          <joint name="J3" type="hinge" axis="0 1 0" damping="0.01" armature="0.01" frictionloss="0.01"/>
          <geom name="g_fore" type="capsule" fromto="0 0 0 0.2 0 0" size="0.02" density="800" contype="0" conaffinity="0"/>
          <body name="finger" pos="0.2 0 0">
            <joint name="grip" type="slide" axis="0 1 0" damping="2" armature="0.001" limited="true" range="-0.045 0.001"/>
            <geom name="g_finger" type="box" size="0.03 0.02 0.03" density="2000" contype="0" conaffinity="0"/>
          </body>
        </body>
      </body>
    </body>
  </worldbody>
  <actuator>
    <general gaintype="user" biastype="user" name="A1" joint="J1" ctrlrange="-6.1959 6.1959" forcerange="-330 330" gainprm="12 0 0 0 0 70 .05 1.0 .97 2.094" user="1"/>
    <general gaintype="user" biastype="user" name="A2" joint="J2" ctrlrange="-2.8 2.8" forcerange="-150 150" gainprm="12 0.5 0.05 0.01 0.3 10 .10 1.5 .97 3.142" user="1"/>
    <general gaintype="user" biastype="user" name="A3" joint="J3" ctrlrange="-6.1959 6.1959" forcerange="-56 56" gainprm="12 0 0 0 0 20 0 0 .97 3.142" user="1"/>
    <general gaintype="user" biastype="user" name="AG" joint="grip" ctrllimited="true" ctrlrange="-.04473 0" forcelimited="true" forcerange="-230 230" gainprm="2000 1000.0 .2 0.005 0.1 0.00"/>
  </actuator>
</mujoco>
"""
