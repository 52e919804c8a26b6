/* rg_dyn.inl -- smooth dynamics stages of the fused step (S1-S6, S10-S12 of SURVEY.md 8(a')):
 * kinematics, motion axes, spatial inertias, composite-rigid-body mass matrix, velocity-product
 * and gravity bias (recursive Newton-Euler in world coordinates), tendons with pulley wrapping,
 * transmission, passive forces, PID actuation.  Replaces the smooth half of mj_forward that
 * robogym reaches through sim.step()/sim.forward() (robogym/mujoco/simulation_interface.py:184-185).
 *
 * Warp-per-environment phase code, see rg_defs.h.  `s` is the warp's shared-memory scratch.
 */
#pragma once
#include "rg_defs.h"

/* Per-environment context.  It is passed BY VALUE (a handful of registers) to the non-inlined stage functions:
 * behind a reference it would live in local memory and every scratch / layout access would start with an LDL. */
#ifdef RG_EMU
struct RgCtx {
  RgMRef mref;           /* the model view (see RG_MDEREF) */
  const RgLayout* L;
  float* s;              /* per-warp scratch */
  const float* xfrc;     /* this env's xfrc_applied [nbody*6] or nullptr */
  float timestep;        /* per-env timestep (opt.timestep unless overridden) */
};
#define RG_CL(c) (*(c).L)
#else
struct RgCtx {
  RgMRef mref;           /* byte offset of the model view in the CTA's dynamic shared memory */
  int soff;              /* float offset of this warp's scratch in it */
  const float* xfrc;     /* global: this env's xfrc_applied [nbody*6] or nullptr */
  float timestep;        /* per-env timestep (opt.timestep unless overridden) */
};
#define RG_CL(c) (RG_MDEREF((c).mref).L)
#endif

#define RG_SI(c, k) (((int*)(RG_SCRATCH(c) + RG_CL(c).scal))[k])
enum { RG_S_NCON = 0, RG_S_NEL = 1, RG_S_WARN = 2, RG_S_NITER = 3, RG_S_TL0 = 4, RG_S_SIG = 5 /* signature of the solver's active set, see rg_solver_update */,
       RG_S_WORK = 6 /* work estimate of this launch (Newton iterations, narrow-phase pairs): orders the next launch */,
       RG_S_CONE = 7 /* some elliptic contact sits in the middle zone of its cone: its Hessian block moves with every step */ };

/* optional per-stage cycle counters (lane 0 of each warp), dumped at the end of RG_DBG */
#if !defined(RG_EMU) && defined(RG_PROFILE)
#define RG_PROF_BEGIN long long prof_t_ = clock64();
#define RG_PROF(c, k) { const long long t2_ = clock64(); if ((threadIdx.x & 31) == 0) RG_SCRATCH(c)[RG_CL(c).scal + 8 + (k)] += (float)(t2_ - prof_t_); prof_t_ = clock64(); }
#else
#define RG_PROF_BEGIN
#define RG_PROF(c, k)
#endif
/* -DRG_PROFILE=2: slots 9..15 break the Newton solve down instead of the collision stage */
#if !defined(RG_EMU) && defined(RG_PROFILE) && RG_PROFILE == 2
#define RG_PROFS_BEGIN RG_PROF_BEGIN
#define RG_PROFS(c, k) RG_PROF(c, k)
#define RG_PROFC(c, k)
#else
#define RG_PROFS_BEGIN
#define RG_PROFS(c, k)
#define RG_PROFC(c, k) RG_PROF(c, k)
#endif

/* Spatial vectors of a kinematic tree are expressed about that tree's own reference point (the
 * world position of its root body), not the world origin: in fp32 the parallel-axis terms m*c^2
 * would otherwise swamp the link inertias of anything that drifts far away (a dropped cube). */
RG_DEV const float* rg_body_ref(const RgCtx c, int body) { return RG_SCRATCH(c) + RG_CL(c).xpos + 3 * RG_MDEREF(c.mref).body_rootid[body]; }
RG_DEV const float* rg_dof_ref(const RgCtx c, int dof) { return rg_body_ref(c, RG_MDEREF(c.mref).dof_bodyid[dof]); }
/* world position of a body's centre of mass (recomputed where needed: two uses per step do not earn it a scratch array) */
RG_DEV void rg_body_xipos(const RgCtx c, int b, float* out) {
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  float t[3];
  rg_rot(t, RG_SCRATCH(c) + RG_CL(c).xquat + 4 * b, m.body_ipos + 3 * b);
  rg_add3(out, RG_SCRATCH(c) + RG_CL(c).xpos + 3 * b, t);
}
/* translational Jacobian column of dof d at world point p */
RG_DEV void rg_jacp_world(const RgCtx c, int d, const float* p, float* jp) {
  float rel[3];
  rg_sub3(rel, p, rg_dof_ref(c, d));
  rg_jacp(jp, RG_SCRATCH(c) + RG_CL(c).S + 6 * d, rel);
}

RG_DEV int rg_ctz(unsigned x) {
#ifdef RG_EMU
  return __builtin_ctz(x);
#else
  return __ffs((int)x) - 1;
#endif
}

/* apply joint j (of body b) to the running frame (pos, quat) */
RG_DEV_NOINLINE void rg_apply_joint(RgMRef mr, const float* qpos, int j, float* pos, float* quat) {
  const RG_MODEL_T& m = RG_MDEREF(mr);
  const int type = m.jnt_type[j], qa = m.jnt_qposadr[j];
  if (type == RG_JNT_FREE) {
    pos[0] = qpos[qa]; pos[1] = qpos[qa + 1]; pos[2] = qpos[qa + 2];
    quat[0] = qpos[qa + 3]; quat[1] = qpos[qa + 4]; quat[2] = qpos[qa + 5]; quat[3] = qpos[qa + 6];
    rg_quat_norm(quat);
  } else if (type == RG_JNT_SLIDE) {
    float ax[3];
    rg_rot(ax, quat, m.jnt_axis + 3 * j);
    rg_addscl3(pos, ax, qpos[qa] - m.qpos0[qa]);
  } else {
    float t[3], anchor[3], ql[4], qn[4];
    rg_rot(t, quat, m.jnt_pos + 3 * j);
    rg_add3(anchor, pos, t);
    if (type == RG_JNT_HINGE) {
      float half = 0.5f * (qpos[qa] - m.qpos0[qa]);
      float sn, cs;
      RG_SINCOS(half, &sn, &cs);
      ql[0] = cs; ql[1] = m.jnt_axis[3 * j] * sn; ql[2] = m.jnt_axis[3 * j + 1] * sn; ql[3] = m.jnt_axis[3 * j + 2] * sn;
    } else {
      ql[0] = qpos[qa]; ql[1] = qpos[qa + 1]; ql[2] = qpos[qa + 2]; ql[3] = qpos[qa + 3];
      rg_quat_norm(ql);
    }
    rg_quat_mul(qn, quat, ql);
    quat[0] = qn[0]; quat[1] = qn[1]; quat[2] = qn[2]; quat[3] = qn[3];
    rg_rot(t, quat, m.jnt_pos + 3 * j);
    rg_sub3(pos, anchor, t);
  }
}

/* ---------------------------------------------------------------- S1 kinematics + axes */
RG_DEV_NOINLINE void rg_kinematics(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  /* local frames of every body relative to its parent */
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int b = lane; b < m.nbody; b += 32) {
    float pos[3] = {m.body_pos[3 * b], m.body_pos[3 * b + 1], m.body_pos[3 * b + 2]};
    float quat[4] = {m.body_quat[4 * b], m.body_quat[4 * b + 1], m.body_quat[4 * b + 2], m.body_quat[4 * b + 3]};
    const int ja = m.body_jntadr[b], jn = m.body_jntnum[b];
    for (int k = 0; k < jn; k++) rg_apply_joint(RG_MREF(m), s + L.qpos, ja + k, pos, quat);
    if (m.body_mocapid[b] >= 0) {   /* mocap body (child of the world, no joints): its pose is data, not model */
      const float* mp = s + L.mocap + 7 * m.body_mocapid[b];
      rg_copy3(pos, mp);
      quat[0] = mp[3]; quat[1] = mp[4]; quat[2] = mp[5]; quat[3] = mp[6];
    }
    rg_copy3(s + L.lpos + 3 * b, pos);
    float* lq = s + L.lquat + 4 * b;
    lq[0] = quat[0]; lq[1] = quat[1]; lq[2] = quat[2]; lq[3] = quat[3];
  }
  RG_PHASE_END
  /* world frames: compose along the ancestor chain (no level barriers needed) */
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int b = lane; b < m.nbody; b += 32) {
    float p[3], q[4];
    rg_copy3(p, s + L.lpos + 3 * b);
    const float* lq = s + L.lquat + 4 * b;
    q[0] = lq[0]; q[1] = lq[1]; q[2] = lq[2]; q[3] = lq[3];
    int a = b == 0 ? 0 : m.body_parentid[b];
    while (a != 0) {
      const float* aq = s + L.lquat + 4 * a;
      float t[3], qn[4];
      rg_rot(t, aq, p);
      rg_add3(p, s + L.lpos + 3 * a, t);
      rg_quat_mul(qn, aq, q);
      q[0] = qn[0]; q[1] = qn[1]; q[2] = qn[2]; q[3] = qn[3];
      a = m.body_parentid[a];
    }
    rg_quat_norm(q);
    rg_copy3(s + L.xpos + 3 * b, p);
    float* xq = s + L.xquat + 4 * b;
    xq[0] = q[0]; xq[1] = q[1]; xq[2] = q[2]; xq[3] = q[3];
  }
  RG_PHASE_END
  /* geom / site positions, motion axes */
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int g = lane; g < m.ngeom; g += 32) {
    const int b = m.geom_bodyid[g];
    float t[3];
    rg_rot(t, s + L.xquat + 4 * b, m.geom_pos + 3 * g);
    rg_add3(s + L.gxpos + 3 * g, s + L.xpos + 3 * b, t);
  }
  RG_NOUNROLL for (int k = lane; k < m.nsite; k += 32) {
    const int b = m.site_bodyid[k];
    float t[3];
    rg_rot(t, s + L.xquat + 4 * b, m.site_pos + 3 * k);
    rg_add3(s + L.sxpos + 3 * k, s + L.xpos + 3 * b, t);
  }
  RG_NOUNROLL for (int d = lane; d < m.nv; d += 32) {
    const int b = m.dof_bodyid[d], j = m.dof_jntid[d], par = m.body_parentid[b];
    const int type = m.jnt_type[j];
    float* S = s + L.S + 6 * d;
    float pos[3], quat[4], t[3];
    rg_rot(t, s + L.xquat + 4 * par, m.body_pos + 3 * b);
    rg_add3(pos, s + L.xpos + 3 * par, t);
    rg_quat_mul(quat, s + L.xquat + 4 * par, m.body_quat + 4 * b);
    for (int jj = m.body_jntadr[b]; jj < j; jj++) rg_apply_joint(RG_MREF(m), s + L.qpos, jj, pos, quat);
    if (type == RG_JNT_SLIDE) {
      S[0] = S[1] = S[2] = 0;
      rg_rot(S + 3, quat, m.jnt_axis + 3 * j);
    } else if (type == RG_JNT_HINGE) {
      float anchor[3];
      rg_rot(t, quat, m.jnt_pos + 3 * j);
      rg_add3(anchor, pos, t);
      rg_sub3(anchor, anchor, rg_body_ref(c, b));
      rg_rot(S, quat, m.jnt_axis + 3 * j);
      rg_cross(S + 3, anchor, S);
    } else {
      const int k = d - m.jnt_dofadr[j];
      if (type == RG_JNT_FREE && k < 3) {
        S[0] = S[1] = S[2] = 0; S[3] = S[4] = S[5] = 0; S[3 + k] = 1.0f;
      } else {
        /* body-frame axis of the final body orientation, through the anchor (ball) or origin (free) */
        float anchor[3], R[9];
        if (type == RG_JNT_BALL) { rg_rot(t, quat, m.jnt_pos + 3 * j); rg_add3(anchor, pos, t); }
        else rg_copy3(anchor, s + L.xpos + 3 * b);
        rg_sub3(anchor, anchor, rg_body_ref(c, b));
        rg_quat2mat(R, s + L.xquat + 4 * b);
        const int a = type == RG_JNT_BALL ? k : k - 3;
        S[0] = R[a]; S[1] = R[3 + a]; S[2] = R[6 + a];
        rg_cross(S + 3, anchor, S);
      }
    }
  }
  RG_PHASE_END
}

/* ---------------------------------------------------------------- S2/S5 inertias + mass matrix */
/* spatial inertia of body b about its tree's reference point, world axes: mass, m c, then the 6 second moments (xx yy zz xy xz yz) */
RG_DEV void rg_body_inertia10(const RgCtx c, int b, float* I) {
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  float q[4], R[9];
  rg_quat_mul(q, s + L.xquat + 4 * b, m.body_iquat + 4 * b);
  rg_quat2mat(R, q);
  const float* di = m.body_inertia + 3 * b;
  const float mass = m.body_mass[b];
  float cm[3];
  rg_body_xipos(c, b, cm);
  rg_sub3(cm, cm, rg_body_ref(c, b));
  float Ic[6]; /* xx yy zz xy xz yz */
  Ic[0] = R[0] * R[0] * di[0] + R[1] * R[1] * di[1] + R[2] * R[2] * di[2];
  Ic[1] = R[3] * R[3] * di[0] + R[4] * R[4] * di[1] + R[5] * R[5] * di[2];
  Ic[2] = R[6] * R[6] * di[0] + R[7] * R[7] * di[1] + R[8] * R[8] * di[2];
  Ic[3] = R[0] * R[3] * di[0] + R[1] * R[4] * di[1] + R[2] * R[5] * di[2];
  Ic[4] = R[0] * R[6] * di[0] + R[1] * R[7] * di[1] + R[2] * R[8] * di[2];
  Ic[5] = R[3] * R[6] * di[0] + R[4] * R[7] * di[1] + R[5] * R[8] * di[2];
  const float cc = rg_dot3(cm, cm);
  I[0] = mass;
  I[1] = mass * cm[0]; I[2] = mass * cm[1]; I[3] = mass * cm[2];
  I[4] = Ic[0] + mass * (cc - cm[0] * cm[0]);
  I[5] = Ic[1] + mass * (cc - cm[1] * cm[1]);
  I[6] = Ic[2] + mass * (cc - cm[2] * cm[2]);
  I[7] = Ic[3] - mass * cm[0] * cm[1];
  I[8] = Ic[4] - mass * cm[0] * cm[2];
  I[9] = Ic[5] - mass * cm[1] * cm[2];
}

RG_DEV_NOINLINE void rg_massmatrix(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int nv = m.nv;
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int b = lane; b < m.nbody; b += 32) rg_body_inertia10(c, b, s + L.I10 + 10 * b);
  RG_PHASE_END
  /* composite inertias: subtree(b) is the contiguous id range [b, b+size) */
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < m.nbody * 10; i += 32) {
    const int b = i / 10, k = i - 10 * b;
    float acc = 0.0f;
    const int e = b + m.body_subtreesize[b];
    RG_NOUNROLL for (int bb = b; bb < e; bb++) acc += s[L.I10 + 10 * bb + k];
    s[L.crb + i] = acc;
  }
  RG_PHASE_END
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < nv; i += 32) {
    float F[6];
    rg_inertia_mul(F, s + L.crb + 10 * m.dof_bodyid[i], s + L.S + 6 * i);
    int j = i;
    float* row = s + L.M + m.dof_mrow[3 * i];   /* tree-sparse row: M(i,i), M(i,parent), M(i,grandparent), ... */
    for (int k = 0; j >= 0; k++) {
      float v = rg_dot6(s + L.S + 6 * j, F);
      if (j == i) v += m.dof_armature[i];
      row[k] = v;
      j = m.dof_parentid[j];
    }
  }
  RG_PHASE_END
}

/* ---------------------------------------------------------------- velocities + bias force */
RG_DEV_NOINLINE void rg_bias(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const float* qvel = s + L.qvel;
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int d = lane; d < m.nv; d += 32) {
    /* velocity the axis of dof d rides on: every ancestor dof EXCEPT the other rotational dofs of its own ball / free
       joint -- their axes are fixed in the same body and turn together (MuJoCo's mj_comVel computes the three
       cdof_dot of a ball joint from the velocity before any of them is added); a free joint's translations do count */
    float V[6] = {0, 0, 0, 0, 0, 0};
    const int jd = m.dof_jntid[d], jt = m.jnt_type[jd], j0 = m.jnt_dofadr[jd];
    int a = m.dof_parentid[d];
    while (a >= 0) {
      const int own_rot = a >= j0 && ((jt == RG_JNT_BALL) || (jt == RG_JNT_FREE && a >= j0 + 3));
      if (!own_rot) {
        const float* Sa = s + L.S + 6 * a;
        const float qa = qvel[a];
        for (int k = 0; k < 6; k++) V[k] += Sa[k] * qa;
      }
      a = m.dof_parentid[a];
    }
    rg_cross_motion(s + L.Sdot + 6 * d, V, s + L.S + 6 * d);
  }
  RG_PHASE_END
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int b = lane; b < m.nbody; b += 32) {
    float V[6] = {0, 0, 0, 0, 0, 0}, A[6] = {0, 0, 0, 0, 0, 0};
    if (!(m.opt_disableflags[0] & RG_DSBL_GRAVITY)) { A[3] = -m.opt_gravity[0]; A[4] = -m.opt_gravity[1]; A[5] = -m.opt_gravity[2]; }
    RG_NOUNROLL for (int w = 0; w < m.nmaskw; w++) {
      unsigned bits = (unsigned)m.body_dofmask[b * m.nmaskw + w];
      while (bits) {
        const int d = 32 * w + rg_ctz(bits);
        bits &= bits - 1;
        const float qd = qvel[d];
        const float* Sd = s + L.S + 6 * d;
        const float* Sp = s + L.Sdot + 6 * d;
        for (int k = 0; k < 6; k++) { V[k] += Sd[k] * qd; A[k] += Sp[k] * qd; }
      }
    }
    const float* I = s + L.I10 + 10 * b;
    float F[6], H[6], G[6];
    rg_inertia_mul(F, I, A);
    rg_inertia_mul(H, I, V);
    rg_cross_force(G, V, H);
    float* ca = s + L.crb + 10 * b;   /* crb is dead once M is assembled: reuse it for the per-body force */
    for (int k = 0; k < 6; k++) ca[k] = b == 0 ? 0.0f : F[k] + G[k];
  }
  RG_PHASE_END
  /* subtree force sums (into the Sdot slot, dead after the previous phase) */
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < m.nbody * 6; i += 32) {
    const int b = i / 6, k = i - 6 * b;
    float acc = 0.0f;
    const int e = b + m.body_subtreesize[b];
    RG_NOUNROLL for (int bb = b; bb < e; bb++) acc += s[L.crb + 10 * bb + k];
    s[L.Sdot + 6 * b + k] = acc;
  }
  RG_PHASE_END
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int d = lane; d < m.nv; d += 32) s[L.bias + d] = rg_dot6(s + L.S + 6 * d, s + L.Sdot + 6 * m.dof_bodyid[d]);
  RG_PHASE_END
}

/* ---------------------------------------------------------------- S3 tendons (fixed + spatial with pulleys) */
RG_DEV int rg_seg_intersect(const float* p1, const float* p2, const float* p3, const float* p4) {
  const float det = (p4[1] - p3[1]) * (p2[0] - p1[0]) - (p4[0] - p3[0]) * (p2[1] - p1[1]);
  if (fabsf(det) < 1e-20f) return 0;
  const float a = ((p4[0] - p3[0]) * (p1[1] - p3[1]) - (p4[1] - p3[1]) * (p1[0] - p3[0])) / det;
  const float b = ((p2[0] - p1[0]) * (p1[1] - p3[1]) - (p2[1] - p1[1]) * (p1[0] - p3[0])) / det;
  return (a >= 0 && a <= 1 && b >= 0 && b <= 1);
}
/* 2-D wrap of d0 -> circle(rad, origin) -> d1; sd = unit side direction or nullptr */
RG_DEV float rg_wrap_circle(float* pnt, const float* d0, const float* d1, const float* sd, float rad) {
  const float sq0 = d0[0] * d0[0] + d0[1] * d0[1], sq1 = d1[0] * d1[0] + d1[1] * d1[1], sqr = rad * rad;
  const float dif[2] = {d1[0] - d0[0], d1[1] - d0[1]};
  const float dd = dif[0] * dif[0] + dif[1] * dif[1];
  if (sq0 < sqr || sq1 < sqr || rad < 1e-12f || dd < 1e-20f) return -1.0f;
  float a = rg_clamp(-(dif[0] * d0[0] + dif[1] * d0[1]) / dd, 0.0f, 1.0f);
  const float tmp[2] = {a * dif[0] + d0[0], a * dif[1] + d0[1]};
  if (tmp[0] * tmp[0] + tmp[1] * tmp[1] > sqr && (!sd || tmp[0] * sd[0] + tmp[1] * sd[1] >= 0)) return -1.0f;
  float sol[2][4], good[2];
  const float sqrt0 = sqrtf(sq0 - sqr), sqrt1 = sqrtf(sq1 - sqr);
  for (int i = 0; i < 2; i++) {
    const float sgn = i == 0 ? 1.0f : -1.0f;
    sol[i][0] = (d0[0] * sqr + sgn * rad * d0[1] * sqrt0) / sq0;
    sol[i][1] = (d0[1] * sqr - sgn * rad * d0[0] * sqrt0) / sq0;
    sol[i][2] = (d1[0] * sqr - sgn * rad * d1[1] * sqrt1) / sq1;
    sol[i][3] = (d1[1] * sqr + sgn * rad * d1[0] * sqrt1) / sq1;
    if (sd) {
      const float t0 = sol[i][0] + sol[i][2], t1 = sol[i][1] + sol[i][3];
      const float n = fmaxf(sqrtf(t0 * t0 + t1 * t1), 1e-20f);
      good[i] = (t0 * sd[0] + t1 * sd[1]) / n;
    } else {
      const float t0 = sol[i][0] - sol[i][2], t1 = sol[i][1] - sol[i][3];
      good[i] = -(t0 * t0 + t1 * t1);
    }
    if (rg_seg_intersect(d0, sol[i], d1, sol[i] + 2)) good[i] = -10000.0f;
  }
  const int i = good[0] > good[1] ? 0 : 1;
  pnt[0] = sol[i][0]; pnt[1] = sol[i][1]; pnt[2] = sol[i][2]; pnt[3] = sol[i][3];
  if (rg_seg_intersect(d0, pnt, d1, pnt + 2)) return -1.0f;
  return rad * acosf(rg_clamp((pnt[0] * pnt[2] + pnt[1] * pnt[3]) / sqr, -1.0f, 1.0f));
}
/* x0 -> wrap geom -> x1: tangent points wp[0..2], wp[3..5]; returns curved length or -1 */
RG_DEV_NOINLINE float rg_wrap_geom(float* wp, const float* x0, const float* x1, const float* gpos, const float* gmat, float rad, int type, const float* side) {
  float t[3], p0[3], p1[3];
  rg_sub3(t, x0, gpos); rg_mulmatT3(p0, gmat, t);
  rg_sub3(t, x1, gpos); rg_mulmatT3(p1, gmat, t);
  if (rg_dot3(p0, p0) < 1e-24f || rg_dot3(p1, p1) < 1e-24f) return -1.0f;
  float ax0[3] = {1, 0, 0}, ax1[3] = {0, 1, 0};
  if (type == RG_WRAP_SPHERE) {
    float nrm[3];
    rg_copy3(ax0, p0); rg_normalize3(ax0);
    rg_cross(nrm, p0, p1);
    if (rg_normalize3(nrm) < 1e-12f) {
      float e[3] = {1, 0, 0};
      if (fabsf(ax0[0]) > 0.9f) { e[0] = 0; e[1] = 1; }
      rg_cross(nrm, ax0, e); rg_normalize3(nrm);
    }
    rg_cross(ax1, nrm, ax0); rg_normalize3(ax1);
  }
  const float d0[2] = {rg_dot3(p0, ax0), rg_dot3(p0, ax1)}, d1[2] = {rg_dot3(p1, ax0), rg_dot3(p1, ax1)};
  float sd2[2];
  const float* sdp = nullptr;
  if (side) {
    float sl[3];
    rg_sub3(t, side, gpos); rg_mulmatT3(sl, gmat, t);
    sd2[0] = rg_dot3(sl, ax0); sd2[1] = rg_dot3(sl, ax1);
    const float n = sqrtf(sd2[0] * sd2[0] + sd2[1] * sd2[1]);
    if (n < rad) return -1.0f; /* inside wrap: not used by the robogym models */
    sd2[0] /= n; sd2[1] /= n;
    sdp = sd2;
  }
  float pnt[4];
  float wlen = rg_wrap_circle(pnt, d0, d1, sdp, rad);
  if (wlen < 0) return -1.0f;
  float r0[3], r1[3];
  for (int i = 0; i < 3; i++) { r0[i] = ax0[i] * pnt[0] + ax1[i] * pnt[1]; r1[i] = ax0[i] * pnt[2] + ax1[i] * pnt[3]; }
  if (type == RG_WRAP_CYLINDER) {
    const float L0 = sqrtf((d0[0] - pnt[0]) * (d0[0] - pnt[0]) + (d0[1] - pnt[1]) * (d0[1] - pnt[1]));
    const float L1 = sqrtf((d1[0] - pnt[2]) * (d1[0] - pnt[2]) + (d1[1] - pnt[3]) * (d1[1] - pnt[3]));
    const float tot = L0 + wlen + L1;
    r0[2] = p0[2] + (p1[2] - p0[2]) * L0 / tot;
    r1[2] = p0[2] + (p1[2] - p0[2]) * (L0 + wlen) / tot;
    const float h = r1[2] - r0[2];
    wlen = sqrtf(wlen * wlen + h * h);
  }
  rg_mulmat3(t, gmat, r0); rg_add3(wp, t, gpos);
  rg_mulmat3(t, gmat, r1); rg_add3(wp + 3, t, gpos);
  return wlen;
}
RG_DEV_NOINLINE void rg_tendon_seg_jac(const RgCtx c, float* J, int ba, const float* pa, int bb, const float* pb, const float* dir, float scale) {
  if (ba == bb) return;
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  /* only the dofs that move exactly one of the two bodies contribute: walk the set bits of the symmetric difference of their
     ancestor masks instead of testing every dof */
  RG_NOUNROLL for (int w = 0; w < m.nmaskw; w++) {
    const unsigned ma = (unsigned)m.body_dofmask[ba * m.nmaskw + w], mb = (unsigned)m.body_dofmask[bb * m.nmaskw + w];
    unsigned bits = ma ^ mb;
    while (bits) {
      const int bit = rg_ctz(bits);
      bits &= bits - 1;
      const int d = 32 * w + bit;
      const int inb = (mb >> bit) & 1u;
      float jp[3];
      rg_jacp_world(c, d, inb ? pb : pa, jp);
      J[d] += (inb ? scale : -scale) * rg_dot3(jp, dir);
    }
  }
}

/* entry (t, d) of the sparse tendon Jacobian */
RG_DEV float rg_tendon_J(const RgCtx c, int t, int d) {
  const int n = ((const int*)(RG_SCRATCH(c) + RG_CL(c).tJn))[t];
  const unsigned char* ji = (const unsigned char*)(RG_SCRATCH(c) + RG_CL(c).tJi) + RG_TJ * t;
  float v = 0.0f;
  RG_NOUNROLL for (int k = 0; k < n; k++) if (ji[k] == d) v = RG_SCRATCH(c)[RG_CL(c).tJv + RG_TJ * t + k];
  return v;
}
RG_DEV_NOINLINE void rg_tendon(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int nv = m.nv;
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int t = lane; t < m.ntendon; t += 32) {
    float* J = s + L.H + t * nv;   /* dense scratch row in the (currently dead) H region, compressed below */
    RG_NOUNROLL for (int k = 0; k < nv; k++) J[k] = 0.0f;
    const int adr = m.tendon_adr[t], num = m.tendon_num[t];
    float len = 0.0f, divisor = 1.0f;
    if (m.wrap_type[adr] == RG_WRAP_JOINT) {
      RG_NOUNROLL for (int w = adr; w < adr + num; w++) {
        const int j = m.wrap_objid[w];
        len += m.wrap_prm[w] * s[L.qpos + m.jnt_qposadr[j]];
        J[m.jnt_dofadr[j]] += m.wrap_prm[w];
      }
    } else {
      int w = adr;
      while (w < adr + num - 1) {
        const int t0 = m.wrap_type[w], t1 = m.wrap_type[w + 1];
        if (t0 == RG_WRAP_PULLEY) { divisor = m.wrap_prm[w]; w++; continue; }
        if (t1 == RG_WRAP_PULLEY) { w++; continue; }
        const int s0 = m.wrap_objid[w];
        const float* x0 = s + L.sxpos + 3 * s0;
        const int b0 = m.site_bodyid[s0];
        const float inv = 1.0f / divisor;
        if (t1 == RG_WRAP_SITE) {
          const int s1 = m.wrap_objid[w + 1];
          const float* x1 = s + L.sxpos + 3 * s1;
          float dir[3];
          rg_sub3(dir, x1, x0);
          len += rg_normalize3(dir) * inv;
          rg_tendon_seg_jac(c, J, b0, x0, m.site_bodyid[s1], x1, dir, inv);
          w += 1;
        } else {
          const int g = m.wrap_objid[w + 1], s1 = m.wrap_objid[w + 2], sid = (int)m.wrap_prm[w + 1];
          const float* x1 = s + L.sxpos + 3 * s1;
          const int b1 = m.site_bodyid[s1], bg = m.geom_bodyid[g];
          float gq[4], gmat[9], wp[6], dir[3];
          rg_quat_mul(gq, s + L.xquat + 4 * bg, m.geom_quat + 4 * g);
          rg_quat_norm(gq);
          rg_quat2mat(gmat, gq);
          const float wlen = rg_wrap_geom(wp, x0, x1, s + L.gxpos + 3 * g, gmat, m.geom_size[3 * g], t1, sid >= 0 ? s + L.sxpos + 3 * sid : nullptr);
          if (wlen < 0) {
            rg_sub3(dir, x1, x0);
            len += rg_normalize3(dir) * inv;
            rg_tendon_seg_jac(c, J, b0, x0, b1, x1, dir, inv);
          } else {
            rg_sub3(dir, wp, x0);
            len += rg_normalize3(dir) * inv;
            rg_tendon_seg_jac(c, J, b0, x0, bg, wp, dir, inv);
            len += wlen * inv;
            rg_sub3(dir, x1, wp + 3);
            len += rg_normalize3(dir) * inv;
            rg_tendon_seg_jac(c, J, bg, wp + 3, b1, x1, dir, inv);
          }
          w += 2;
        }
      }
    }
    s[L.tlen + t] = len;
    float v = 0.0f;
    int nnz = 0;
    unsigned char* ji = (unsigned char*)(s + L.tJi) + RG_TJ * t;
    float* jv = s + L.tJv + RG_TJ * t;
    RG_NOUNROLL for (int k = 0; k < nv; k++) {
      if (J[k] == 0.0f) continue;
      v += J[k] * s[L.qvel + k];
      if (nnz < RG_TJ) { ji[nnz] = (unsigned char)k; jv[nnz] = J[k]; nnz++; }
      else RG_SI(c, RG_S_WARN) |= RG_WARN_TENDON_NNZ;
    }
    ((int*)(s + L.tJn))[t] = nnz;
    s[L.tvel + t] = v;
  }
  RG_PHASE_END
}

/* Equality constraints (weld, joint coupling; robogym/assets/xmls/robot/ur16e/base.xml:52-54, gripper_actuators.xml:2-4): every
 * row becomes a "virtual tendon" behind the real ones -- residual in tlen, its rate in tvel, sparse Jacobian row in tJ* -- so the
 * constraint builder and the solver handle it with the tendon-limit machinery (rg_make_constraints).  One lane per constraint.
 * Weld (mj_instantiateEquality, mjEQ_WELD): body 2 keeps the pose eq_data relative to body 1; residual = position error in
 * world axes, then the vector part of conj(q2) q1 relquat; d/dt of that = 1/2 vec(conj(q2) [0, w1 - w2] q1 relquat). */
RG_DEV_NOINLINE void rg_equality(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int nt = m.ntendon;
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int r0 = lane; r0 < m.neqrow; r0 += 32) {
    const int code = m.eqrow[r0];
    if (code & 7) continue;               /* the lane of a constraint's first row does all of its rows */
    const int e = code >> 3, t0 = nt + r0;
    const float* data = m.eq_data + 7 * e;
    int* tJn = (int*)(s + L.tJn);
    unsigned char* ji = (unsigned char*)(s + L.tJi) + RG_TJ * t0;
    float* jv = s + L.tJv + RG_TJ * t0;
    if (m.eq_type[e] == RG_EQ_JOINT) {
      const int j1 = m.eq_obj1id[e], j2 = m.eq_obj2id[e];
      const int a1 = m.jnt_qposadr[j1], d1 = m.jnt_dofadr[j1];
      float cpos = s[L.qpos + a1] - m.qpos0[a1] - data[0], vel = s[L.qvel + d1];
      int n = 1;
      ji[0] = (unsigned char)d1; jv[0] = 1.0f;
      if (j2 >= 0) {
        const int a2 = m.jnt_qposadr[j2], d2 = m.jnt_dofadr[j2];
        const float dif = s[L.qpos + a2] - m.qpos0[a2];
        cpos -= dif * (data[1] + dif * (data[2] + dif * (data[3] + dif * data[4])));
        const float deriv = data[1] + dif * (2.0f * data[2] + dif * (3.0f * data[3] + dif * 4.0f * data[4]));
        ji[1] = (unsigned char)d2; jv[1] = -deriv; n = 2;
        vel -= deriv * s[L.qvel + d2];
      }
      tJn[t0] = n; s[L.tlen + t0] = cpos; s[L.tvel + t0] = vel;
      continue;
    }
    const int b1 = m.eq_obj1id[e], b2 = m.eq_obj2id[e];
    float p0[3], cpos[6], q[4], qc[4], qe[4], vel[6] = {0, 0, 0, 0, 0, 0};
    const float* p1 = s + L.xpos + 3 * b2;
    rg_rot(p0, s + L.xquat + 4 * b1, data);
    rg_add3(p0, p0, s + L.xpos + 3 * b1);
    rg_sub3(cpos, p0, p1);
    rg_quat_mul(q, s + L.xquat + 4 * b1, data + 3);
    const float* q2 = s + L.xquat + 4 * b2;
    qc[0] = q2[0]; qc[1] = -q2[1]; qc[2] = -q2[2]; qc[3] = -q2[3];
    rg_quat_mul(qe, qc, q);
    cpos[3] = qe[1]; cpos[4] = qe[2]; cpos[5] = qe[3];
    int n = 0;
    RG_NOUNROLL for (int w = 0; w < m.nmaskw; w++) {
      const unsigned m1 = (unsigned)m.body_dofmask[b1 * m.nmaskw + w], m2 = (unsigned)m.body_dofmask[b2 * m.nmaskw + w];
      unsigned bits = m1 | m2;
      while (bits && n < RG_TJ) {
        const int bit = rg_ctz(bits);
        bits &= bits - 1;
        const int d = 32 * w + bit;
        const float* S = s + L.S + 6 * d;
        float col[6] = {0, 0, 0, 0, 0, 0}, ang[3] = {0, 0, 0}, jp[3];
        if ((m1 >> bit) & 1u) { rg_jacp_world(c, d, p0, jp); rg_add3(col, col, jp); rg_add3(ang, ang, S); }
        if ((m2 >> bit) & 1u) { rg_jacp_world(c, d, p1, jp); rg_sub3(col, col, jp); rg_sub3(ang, ang, S); }
        float wq[4] = {0.0f, ang[0], ang[1], ang[2]}, t1[4], t2[4];
        rg_quat_mul(t1, qc, wq);
        rg_quat_mul(t2, t1, q);
        col[3] = 0.5f * t2[1]; col[4] = 0.5f * t2[2]; col[5] = 0.5f * t2[3];
        const float qd = s[L.qvel + d];
        for (int k = 0; k < 6; k++) { ji[RG_TJ * k + n] = (unsigned char)d; jv[RG_TJ * k + n] = col[k]; vel[k] += col[k] * qd; }
        n++;
      }
    }
    for (int k = 0; k < 6; k++) { tJn[t0 + k] = n; s[L.tlen + t0 + k] = cpos[k]; s[L.tvel + t0 + k] = vel[k]; }
  }
  RG_PHASE_END
}

/* ---------------------------------------------------------------- S4/S10/S11/S12 passive, PID actuation, smooth force */
RG_DEV_NOINLINE void rg_forces(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int nv = m.nv, flags = m.opt_disableflags[0];
  const float dt = c.timestep;
  /* actuators: transmission, mujoco-py PID bias callback (stateful), force clamp */
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < m.nu; i += 32) {
    const float gear = m.actuator_gear[6 * i];
    const int id = m.actuator_trnid[i];
    float len, vel;
    if (m.actuator_trntype[i] == RG_TRN_JOINT) { len = gear * s[L.qpos + m.jnt_qposadr[id]]; vel = gear * s[L.qvel + m.jnt_dofadr[id]]; }
    else { len = gear * s[L.tlen + id]; vel = gear * s[L.tvel + id]; }
    s[L.alen + i] = len;
    float ctrl = s[L.ctrl + i];
    if (m.actuator_ctrllimited[i] && !(flags & RG_DSBL_CLAMPCTRL)) ctrl = rg_clamp(ctrl, m.actuator_ctrlrange[2 * i], m.actuator_ctrlrange[2 * i + 1]);
    float gain = 0.0f, bias = 0.0f;
    if (m.actuator_gaintype[i] == RG_GAIN_FIXED) gain = m.actuator_gainprm[10 * i];
    const float* bp = m.actuator_biasprm + 10 * i;
    const float lo = m.actuator_forcerange[2 * i], hi = m.actuator_forcerange[2 * i + 1];
    if (m.actuator_biastype[i] == RG_BIAS_AFFINE) bias = bp[0] + bp[1] * len + bp[2] * vel;
    else if (m.actuator_biastype[i] == RG_BIAS_USER && m.opt_pid[0] && m.pidw * (i + 1) <= m.nuserdata && m.actuator_user0[i] == 1.0f) {
      /* mujoco-py's cascaded-PI callback (mjpid.pyx; UR16e default calibration, robogym/assets/xmls/robot/ur16e/jointspec/
         calibrations/cascaded_pi/joint_actuations.xml:4-10): gainprm = Kp_x Ti_x clamp_x Td_x dsmooth_x | Kp_v Ti_v clamp_v | ema max_vel.
         Smoothed set-point -> position PID -> desired velocity (clamped) -> velocity PI + bias-force compensation -> forcerange */
      const float* g = m.actuator_gainprm + 10 * i;
      float* ud = s + L.pid + 6 * i;
      const float raw = s[L.ctrl + i];
      const float ema = ud[5] != 0.0f ? g[8] * ud[4] + (1.0f - g[8]) * raw : raw;
      ud[4] = ema;
      float des = raw;
      if (g[0] != 0.0f) {
        const float err = ema - len;
        const float integ = rg_clamp(ud[0] + err * dt, -g[2], g[2]);
        const float deriv = (1.0f - g[4]) * ud[2] + g[4] * (err - ud[1]) / dt;
        des = g[0] * (err + (g[1] != 0.0f ? integ / g[1] : 0.0f) + g[3] * deriv);
        ud[0] = integ; ud[1] = err; ud[2] = deriv;
      }
      des = rg_clamp(des, -g[9], g[9]);
      const float errv = des - vel;
      const float integv = rg_clamp(ud[3] + errv * dt, -g[7], g[7]);
      ud[3] = integv;
      bias = g[5] * (errv + (g[6] != 0.0f ? integv / g[6] : 0.0f));
      if (m.actuator_trntype[i] == RG_TRN_JOINT) bias += s[L.bias + m.jnt_dofadr[id]];
      if (lo != 0.0f || hi != 0.0f) bias = rg_clamp(bias, lo, hi);
    }
    else if (m.actuator_biastype[i] == RG_BIAS_USER && m.opt_pid[0] && m.pidw * (i + 1) <= m.nuserdata) {
      const float* g = m.actuator_gainprm + 10 * i;
      float err = s[L.ctrl + i] - len;
      if (fabsf(err) < g[5]) err = 0.0f;
      float* ud = s + L.pid + m.pidw * i;
      const float integ = rg_clamp(ud[0] + err * dt, -g[2], g[2]);
      float deriv = (err - ud[1]) / dt;
      deriv = (1.0f - g[4]) * ud[2] + g[4] * deriv;
      bias = g[0] * (err + (g[1] != 0.0f ? integ / g[1] : 0.0f) + g[3] * deriv);
      ud[0] = integ; ud[1] = err; ud[2] = deriv;
      if (lo != 0.0f || hi != 0.0f) bias = rg_clamp(bias, lo, hi);
    }
    float f = gain * ctrl + bias;
    if (m.actuator_forcelimited[i]) f = rg_clamp(f, lo, hi);
    if (flags & RG_DSBL_ACTUATION) f = 0.0f;
    s[L.aforce + i] = f;
  }
  /* tendon spring-damper force, stored over tvel's twin slot in tmp */
  RG_NOUNROLL for (int t = lane; t < m.ntendon; t += 32)
    s[L.tmp + t] = -m.tendon_stiffness[t] * (s[L.tlen + t] - m.tendon_lengthspring[t]) - m.tendon_damping[t] * s[L.tvel + t];
  RG_PHASE_END
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int d = lane; d < nv; d += 32) {
    float passive = 0.0f;
    if (!(flags & RG_DSBL_PASSIVE)) {
      const int j = m.dof_jntid[d];
      if (m.jnt_stiffness[j] != 0.0f && (m.jnt_type[j] == RG_JNT_SLIDE || m.jnt_type[j] == RG_JNT_HINGE))
        passive -= m.jnt_stiffness[j] * (s[L.qpos + m.jnt_qposadr[j]] - m.qpos_spring[m.jnt_qposadr[j]]);
      passive -= m.dof_damping[d] * s[L.qvel + d];
      RG_NOUNROLL for (int t = 0; t < m.ntendon; t++) passive += rg_tendon_J(c, t, d) * s[L.tmp + t];
    }
    float act = 0.0f;
    RG_NOUNROLL for (int i = 0; i < m.nu; i++) {
      const float gear = m.actuator_gear[6 * i];
      const int id = m.actuator_trnid[i];
      if (m.actuator_trntype[i] == RG_TRN_JOINT) { if (m.jnt_dofadr[id] == d) act += gear * s[L.aforce + i]; }
      else act += gear * rg_tendon_J(c, id, d) * s[L.aforce + i];
    }
    float applied = 0.0f;
    if (c.xfrc) {
      RG_NOUNROLL for (int b = 1; b < m.nbody; b++) {
        if (!rg_dof_in_body(m, b, d)) continue;
        const float* x = c.xfrc + 6 * b;
        if (x[0] == 0 && x[1] == 0 && x[2] == 0 && x[3] == 0 && x[4] == 0 && x[5] == 0) continue;
        float jp[3];
        float xi[3];
        rg_body_xipos(c, b, xi);
        rg_jacp_world(c, d, xi, jp);
        applied += rg_dot3(jp, x) + rg_dot3(s + L.S + 6 * d, x + 3);
      }
    }
    s[L.smooth + d] = passive - s[L.bias + d] + act + applied;
  }
  RG_PHASE_END
}
