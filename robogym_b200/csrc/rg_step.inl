/* rg_step.inl -- per-environment driver of the fused step and the scratch layout.
 *
 * rg_env_step() is what one warp does for one environment in one launch of rg_step():
 *   load state -> nsub x { mj_step } -> optional mj_forward -> store state + derived outputs,
 * i.e. SimulationInterface.step() = sim.step() + sim.forward()
 * (robogym/mujoco/simulation_interface.py:176-189) for the whole batch in a single kernel.
 */
#pragma once
#include "rg_sol.inl"

/* Batched state: row-major [nenv][n] fp32 tensors owned by the caller (torch). */
struct RgBatchIO {
  int nenv;
  float* qpos; float* qvel; float* ctrl; float* pid; float* warm; float* time;
  const float* xfrc;        /* [nenv][nbody*6] or nullptr */
  const float* timestep;    /* [nenv] per-env opt.timestep override or nullptr */
  const float* mocap_pos;   /* [nenv][nmocap*3] data.mocap_pos or nullptr (then the mocap bodies keep their model pose) */
  const float* mocap_quat;  /* [nenv][nmocap*4] data.mocap_quat or nullptr */
  /* derived outputs (any may be nullptr) */
  float* site_xpos; float* body_xpos; float* body_xquat; float* geom_xpos; float* act_force; float* qacc;
  float* sensordata;        /* [nenv][nsensordata] data.sensordata after the last forward pass (joint positions, touch, force / torque) */
  float* body_xvel;         /* [nenv][nbody][6]: angular then linear velocity of the body frame, world axes (data.get_body_xvelr / get_body_xvelp) */
  float* contact;           /* [nenv][contact capacity][4] = geom1, geom2, dist, dim */
  int* ncon; int* warn;
  int* sep;                 /* [nenv][RG_NSEP] separating-axis cache carried from launch to launch (engine-internal, may be nullptr) */
  int* cost;                /* [nenv] work estimate of this launch (engine-internal, see rg_order_kernel) */
  float* dbg;               /* [nenv][rg_dbg_size] stage dump for the parity tests */
};

#ifdef RG_EMU
#define RG_HD static inline
#else
#define RG_HD __host__ __device__ static inline
#endif
template <class MT>
RG_HD int rg_dbg_size(const MT& m, int ncon_cap) {
  return m.nv * m.nv + 6 * m.nv + m.ntendon + 2 * m.nu + 4 + ncon_cap * RG_CON_STRIDE + m.ntendon * m.nv + RG_NPROF;
}

static inline int rg_imax(int a, int b) { return a > b ? a : b; }

/* host: compute the per-warp scratch layout for a model */
static inline RgLayout rg_make_layout(const RgModel& m, int ncon = RG_NCON, int nel = RG_NEL, int tile = RG_TILE) {
  RgLayout L;
  int o = 0;
  if (tile > 32) tile = 32;
  if (tile > m.nv) tile = m.nv > 0 ? m.nv : 1;
  if (ncon < 1) ncon = 1;
  if (nel < 1) nel = 1;
  if (nel > 1022) nel = 1022;   /* element ids travel in 10 bits (eldof) */
  L.ncon = ncon; L.nel = nel; L.tile = tile;
#define RG_ALLOC(field, n) do { L.field = o; o += (n); } while (0)   /* scalar 4-byte accesses only: no padding between arrays */
  RG_ALLOC(qpos, m.nq); RG_ALLOC(qvel, m.nv); RG_ALLOC(ctrl, m.nu); RG_ALLOC(pid, m.pidw * m.nu); RG_ALLOC(warm, m.nv);
  RG_ALLOC(xpos, 3 * m.nbody); RG_ALLOC(xquat, 4 * m.nbody);
  RG_ALLOC(gxpos, 3 * m.ngeom); RG_ALLOC(sxpos, 3 * m.nsite);
  const int ntri = ((m.ns + 1) * (m.ns + 2)) >> 1;   /* packed lower triangle of H plus one extra row (the right-hand side rides along in the factorisation) */
  RG_ALLOC(S, 6 * m.nv); RG_ALLOC(M, m.nM);   /* M: tree-sparse rows (rg_host.h), H: dense packed lower triangle */
  /* H aliases the smooth-dynamics temporaries */
  const int h0 = o;
  /* body inertias (I10, crb) live from the mass-matrix stage to the bias stage only: they sit in the contact records, which
     are rewritten by the collision stage afterwards (when those are big enough) */
  const int inertia_in_con = 20 * m.nbody <= ncon * RG_CON_STRIDE;
  RG_ALLOC(Sdot, 6 * rg_imax(m.nv, m.nbody));
  if (!inertia_in_con) { RG_ALLOC(I10, 10 * m.nbody); RG_ALLOC(crb, 10 * m.nbody); }
  L.H = h0;
  o = h0 + rg_imax(rg_imax(rg_imax(o - h0, (ntri + 3) & ~3), (m.ntendon * m.nv + 3) & ~3), (m.nM + 3) & ~3);   /* also the tree-sparse factor (Euler, solver-free trees) */
  RG_ALLOC(bias, m.nv); RG_ALLOC(smooth, m.nv); RG_ALLOC(qacc, m.nv);
  RG_ALLOC(Ma, m.nv); RG_ALLOC(search, m.nv); RG_ALLOC(Mv, m.nv); RG_ALLOC(qfc, m.nv);
  RG_ALLOC(tmp, rg_imax(m.nv, m.ntendon));
  const int nvt = m.ntendon + m.neqrow;   /* equality rows ride behind the tendons as "virtual tendons" (rg_equality) */
  RG_ALLOC(tlen, nvt); RG_ALLOC(tvel, nvt); RG_ALLOC(tJn, nvt); RG_ALLOC(tJi, (RG_TJ * nvt + 3) >> 2);   /* dof ids as bytes */ RG_ALLOC(tJv, RG_TJ * nvt); RG_ALLOC(alen, m.nu); RG_ALLOC(aforce, m.nu);
  RG_ALLOC(mocap, 7 * m.nmocap);
  RG_ALLOC(con, ncon * RG_CON_STRIDE); RG_ALLOC(cu, 6 * ncon); RG_ALLOC(cw, 6 * ncon); RG_ALLOC(cF, 6 * ncon); RG_ALLOC(cprm, RG_CPRM * ncon);
  const int nel64 = nel < 64 ? 64 : nel;   /* el_jv / el_f double as the broad-phase candidate lists (64 entries) */
  RG_ALLOC(el_i, nel); RG_ALLOC(el_D, nel);
  RG_ALLOC(el_jar, nel); RG_ALLOC(el_jv, nel64); RG_ALLOC(el_f, nel64);
  /* the Hessian tiles of one contact are built while cw / el_jv are dead (they are written after the factorisation) */
  const int tile_in_cw = 12 * tile <= 6 * ncon && tile <= nel64;
  if (tile_in_cw) { L.tileJ = L.cw; L.tileWJ = L.cw + 6 * tile; L.tileDof = L.el_jv; }
  else { RG_ALLOC(tileJ, 6 * tile); RG_ALLOC(tileWJ, 6 * tile); RG_ALLOC(tileDof, tile); }
  RG_ALLOC(scal, 8 + RG_NPROF);
  if (inertia_in_con) { L.I10 = L.con; L.crb = L.con + 10 * m.nbody; }
  RG_ALLOC(eldof, m.nv); RG_ALLOC(env, m.nv);   /* eldof: the (up to) three single-dof elements of a dof, 10 bits each */ RG_ALLOC(cdof, ((tile + 3) >> 2) * ncon);   /* dof ids as bytes */
  RG_ALLOC(sep, RG_NSEP);   /* lives across the substeps of a launch, so it cannot share storage */
#undef RG_ALLOC
  /* lifetimes that never overlap share storage: local frames (kinematics only) sit in the contact
     solver vectors, the broad-phase candidate lists in the row work arrays */
  if (((3 * m.nbody + 3) & ~3) + 4 * m.nbody <= 12 * ncon) { L.lpos = L.cu; L.lquat = L.cu + ((3 * m.nbody + 3) & ~3); }
  else { L.lpos = o; o += (3 * m.nbody + 3) & ~3; L.lquat = o; o += 4 * m.nbody; }
  L.cand = L.el_jv; L.cand2 = L.el_f;
  /* narrow-phase staging (32 x 8 results + 32 slots): in the contact solver vectors, which are idle during collision */
  if (18 * ncon >= 288) L.stage = L.cu; else { L.stage = o; o += 288; }
  L.total = (o + 3) & ~3;   /* per-warp areas (and the per-warp model views behind them) stay 16-byte aligned */
  return L;
}

/* mj_forward: everything but the integrator.  RG_SYNC_LEVEL picks which stage boundaries are CTA barriers (lock-step keeps
   the warps of a CTA in the same code): 2 = every stage (default), 1 = only before the big stages (collision, constraint
   rows, solver), 0 = before collision and the solver only. */
#ifndef RG_SYNC_LEVEL
#define RG_SYNC_LEVEL 2
#endif
#if RG_SYNC_LEVEL >= 2
#define RG_SYNC_SMALL() RG_CTA_SYNC()
#else
#define RG_SYNC_SMALL()
#endif
#if RG_SYNC_LEVEL >= 1
#define RG_SYNC_MID() RG_CTA_SYNC()
#else
#define RG_SYNC_MID()
#endif
RG_DEV_NOINLINE void rg_forward(const RgCtx c) {
  RG_PROF_BEGIN
  RG_SYNC_SMALL(); rg_kinematics(c); RG_PROF(c, 0)
  RG_SYNC_SMALL(); rg_massmatrix(c); RG_PROF(c, 1)
  RG_SYNC_SMALL(); rg_bias(c); RG_PROF(c, 2)
  RG_SYNC_SMALL(); rg_tendon(c); if (RG_MDEREF(c.mref).neqrow) rg_equality(c); RG_PROF(c, 3)
  RG_SYNC_SMALL(); rg_forces(c); RG_PROF(c, 4)
  RG_CTA_SYNC(); rg_collision(c); RG_PROF(c, 5)
  RG_SYNC_MID(); rg_make_constraints(c); RG_PROF(c, 6)
  RG_CTA_SYNC(); rg_solve(c); RG_PROF(c, 7)
}

/* ---- S14 sensors (evaluated once per launch, after the last forward pass, when RG_FIELD_SENSORDATA is bound) */
/* smallest non-negative root of a t^2 + 2 b t + c = 0 (both roots in xx), or -1 */
RG_DEV float rg_ray_quad(float a, float b, float c, float* xx) {
  float det = b * b - a * c;
  if (det < 1e-15f || a < 1e-15f) { xx[0] = xx[1] = -1.0f; return -1.0f; }
  det = sqrtf(det);
  xx[0] = (-b - det) / a; xx[1] = (-b + det) / a;
  if (xx[0] >= 0.0f) return xx[0];
  if (xx[1] >= 0.0f) return xx[1];
  return -1.0f;
}
/* distance along the ray pnt + t vec to a sphere / capsule site volume, -1 = miss (a point inside always hits) */
RG_DEV_NOINLINE float rg_ray_site(const float* pos, const float* quat, const float* size, const float* pnt, const float* vec, int type) {
  float mat[9], dif[3], lp[3], lv[3], xx[2], x = -1.0f;
  rg_quat2mat(mat, quat);
  rg_sub3(dif, pnt, pos);
  rg_mulmatT3(lp, mat, dif);
  rg_mulmatT3(lv, mat, vec);
  if (type == RG_GEOM_SPHERE) return rg_ray_quad(rg_dot3(lv, lv), rg_dot3(lv, lp), rg_dot3(lp, lp) - size[0] * size[0], xx);
  if (type != RG_GEOM_CAPSULE) return -1.0f;
  const float sol = rg_ray_quad(lv[0] * lv[0] + lv[1] * lv[1], lv[0] * lp[0] + lv[1] * lp[1], lp[0] * lp[0] + lp[1] * lp[1] - size[0] * size[0], xx);
  if (sol >= 0.0f && fabsf(lp[2] + sol * lv[2]) <= size[1]) x = sol;
  for (int cap = 0; cap < 2; cap++) {
    const float zc = cap == 0 ? size[1] : -size[1];
    const float q[3] = {lp[0], lp[1], lp[2] - zc};
    rg_ray_quad(rg_dot3(lv, lv), rg_dot3(lv, q), rg_dot3(q, q) - size[0] * size[0], xx);
    for (int i = 0; i < 2; i++) {
      if (xx[i] < 0.0f) continue;
      const float z = lp[2] + xx[i] * lv[2];
      if ((cap == 0 ? z >= size[1] : z <= -size[1]) && (x < 0.0f || xx[i] < x)) x = xx[i];
    }
  }
  return x;
}
/* one lane per sensor: joint position, or the normal forces of the contacts on the site's body whose point (or the ray from it
   along the contact normal) lies in the site's volume (mjSENS_TOUCH; robogym/assets/xmls/robot/shadowhand/assets.xml:135-142) */
/* mj_rnePostConstraint for one body: the wrench its parent transmits to the subtree rooted at `body` (cfrc_int: sum over the
   subtree of I a + v x* I v minus the external wrench -- xfrc_applied and the contact forces; the world accelerates with
   -gravity), as [torque about the tree's reference point, force] in world axes.  One lane walks the subtree: it runs once
   per launch, for the two sensors of the UR16e tool flange (robogym/assets/xmls/robot/ur16e/base.xml:48-49). */
RG_DEV_NOINLINE void rg_subtree_wrench(const RgCtx c, int body, float* W) {
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int end = body + m.body_subtreesize[body];
  const float* ref = rg_body_ref(c, body);
  for (int k = 0; k < 6; k++) W[k] = 0.0f;
  RG_NOUNROLL for (int b = body; b < end; b++) {
    float V[6] = {0, 0, 0, 0, 0, 0}, Vs[6] = {0, 0, 0, 0, 0, 0}, A[6] = {0, 0, 0, 0, 0, 0};
    if (!(m.opt_disableflags[0] & RG_DSBL_GRAVITY)) { A[3] = -m.opt_gravity[0]; A[4] = -m.opt_gravity[1]; A[5] = -m.opt_gravity[2]; }
    int rot0 = -1, rot1 = -1;   /* rotational dofs of the ball / free joint being walked: their axes turn with the velocity before them */
    RG_NOUNROLL for (int w = 0; w < m.nmaskw; w++) {
      unsigned bits = (unsigned)m.body_dofmask[b * m.nmaskw + w];
      while (bits) {           /* ascending dof ids = root to leaf */
        const int d = 32 * w + rg_ctz(bits);
        bits &= bits - 1;
        if (d >= rot1) {
          const int j = m.dof_jntid[d], jt = m.jnt_type[j], j0 = m.jnt_dofadr[j];
          if (jt == RG_JNT_BALL) { rot0 = j0; rot1 = j0 + 3; }
          else if (jt == RG_JNT_FREE) { rot0 = j0 + 3; rot1 = j0 + 6; }
          else rot0 = rot1 = -1;
        }
        if (d == rot0) for (int k = 0; k < 6; k++) Vs[k] = V[k];
        const float* Sd = s + L.S + 6 * d;
        float Sp[6];
        rg_cross_motion(Sp, (d >= rot0 && d < rot1) ? Vs : V, Sd);
        const float qd = s[L.qvel + d], qa = s[L.qacc + d];
        for (int k = 0; k < 6; k++) { A[k] += Sp[k] * qd + Sd[k] * qa; V[k] += Sd[k] * qd; }
      }
    }
    float I[10], F[6], H[6], G[6];
    rg_body_inertia10(c, b, I);
    rg_inertia_mul(F, I, A);
    rg_inertia_mul(H, I, V);
    rg_cross_force(G, V, H);
    for (int k = 0; k < 6; k++) W[k] += F[k] + G[k];
    if (c.xfrc) {              /* xfrc_applied: force, torque at the body's centre of mass */
      const float* x = c.xfrc + 6 * b;
      float xi[3], t[3];
      rg_body_xipos(c, b, xi);
      rg_sub3(xi, xi, ref);
      rg_cross(t, xi, x);
      for (int k = 0; k < 3; k++) { W[k] -= x[3 + k] + t[k]; W[3 + k] -= x[k]; }
    }
  }
  const int ncon = RG_SI(c, RG_S_NCON);
  RG_NOUNROLL for (int k = 0; k < ncon; k++) {
    const float* r = s + L.con + RG_CON_STRIDE * k;
    const int b1 = (int)r[18], b2 = (int)r[19], dim = (int)s[L.cprm + RG_CPRM * k + 1];
    const int in1 = b1 >= body && b1 < end, in2 = b2 >= body && b2 < end;
    if (in1 == in2 || dim == 0) continue;        /* outside, internal to the subtree, or not in the solver */
    const float* F = s + L.cF + 6 * k;
    float f[3] = {0, 0, 0}, tq[3] = {0, 0, 0}, p[3], t[3];
    for (int a = 0; a < 3 && a < dim; a++) for (int i = 0; i < 3; i++) f[i] += F[a] * r[4 + 3 * a + i];
    for (int a = 3; a < dim; a++) for (int i = 0; i < 3; i++) tq[i] += F[a] * r[4 + 3 * (a - 3) + i];
    const float sgn = in2 ? 1.0f : -1.0f;        /* the contact force acts on geom2's body, its opposite on geom1's */
    rg_sub3(p, r + 1, ref);
    rg_cross(t, p, f);
    for (int i = 0; i < 3; i++) { W[i] -= sgn * (tq[i] + t[i]); W[3 + i] -= sgn * f[i]; }
  }
}

RG_DEV_NOINLINE void rg_sensors(const RgCtx c, float* out) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int ncon = RG_SI(c, RG_S_NCON);
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < m.nsensor; i += 32) {
    const int adr = m.sensor_adr[i], obj = m.sensor_objid[i], type = m.sensor_type[i];
    for (int k = 0; k < m.sensor_dim[i]; k++) out[adr + k] = 0.0f;
    if (type == 8) { const int j = obj, qa = m.jnt_qposadr[j]; out[adr] = s[L.qpos + qa]; }
    if (type == 4 || type == 5) {   /* mjSENS_FORCE / mjSENS_TORQUE: the wrench between the site's body and its parent, in the site's frame */
      const int body = m.site_bodyid[obj];
      float W[6], sq[4], R[9], t[3], rel[3];
      rg_subtree_wrench(c, body, W);
      if (type == 5) {             /* torque about the site */
        rg_sub3(rel, s + L.sxpos + 3 * obj, rg_body_ref(c, body));
        rg_cross(t, rel, W + 3);
        for (int k = 0; k < 3; k++) W[k] -= t[k];
      }
      rg_quat_mul(sq, s + L.xquat + 4 * body, m.site_quat + 4 * obj);
      rg_quat_norm(sq);
      rg_quat2mat(R, sq);
      const float* v = type == 4 ? W + 3 : W;
      for (int k = 0; k < 3; k++) out[adr + k] = R[k] * v[0] + R[3 + k] * v[1] + R[6 + k] * v[2];
    }
    if (type != 0) continue;
    const int body = m.site_bodyid[obj];
    float sq[4], total = 0.0f;
    rg_quat_mul(sq, s + L.xquat + 4 * body, m.site_quat + 4 * obj);
    rg_quat_norm(sq);
    RG_NOUNROLL for (int k = 0; k < ncon; k++) {
      const float* r = s + L.con + RG_CON_STRIDE * k;
      const int b1 = (int)r[18], b2 = (int)r[19];
      if ((body != b1 && body != b2) || (int)s[L.cprm + RG_CPRM * k + 1] == 0) continue;
      const float f = s[L.cF + 6 * k];
      if (!(f > 0.0f)) continue;
      float ray[3] = {r[4], r[5], r[6]};
      if (body == b2) rg_scl3(ray, ray, -1.0f);
      if (rg_ray_site(s + L.sxpos + 3 * obj, sq, m.site_size + 3 * obj, r + 1, ray, m.site_type[obj]) >= 0.0f) total += f;
    }
    out[adr] = total;
  }
  RG_PHASE_END
}

/* `store` == 0: a padding iteration that keeps this warp in step with its CTA (same barriers), results discarded */
RG_DEV_NOINLINE void rg_env_step(RgMRef mr, const RgLayout& L_in, float* s_in, int soff, const RgBatchIO& io, int env, int nsub, int final_forward, int store) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(mr);
  const float* xfrc_env = io.xfrc ? io.xfrc + (size_t)env * m.nbody * 6 : nullptr;
  const float dt_env = io.timestep ? io.timestep[env] : m.opt_timestep[0];
#ifdef RG_EMU
  const RgCtx c = {mr, &L_in, s_in, xfrc_env, dt_env};
#else
  const RgCtx c = {mr, soff, xfrc_env, dt_env};   /* the layout travels inside the shared-memory model view */
#endif
  const RgLayout& L = RG_CL(c);
  float* s = RG_SCRATCH(c);
  const int npid = m.pidw * m.nu;
  /* ---- load (coalesced: consecutive lanes read consecutive floats of this env's rows) */
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < m.nq; i += 32) s[L.qpos + i] = io.qpos[(size_t)env * m.nq + i];
  RG_NOUNROLL for (int i = lane; i < m.nv; i += 32) { s[L.qvel + i] = io.qvel[(size_t)env * m.nv + i]; s[L.warm + i] = io.warm[(size_t)env * m.nv + i]; }
  RG_NOUNROLL for (int i = lane; i < m.nu; i += 32) s[L.ctrl + i] = io.ctrl[(size_t)env * m.nu + i];
  RG_NOUNROLL for (int i = lane; i < npid; i += 32) s[L.pid + i] = io.pid[(size_t)env * npid + i];
  if (lane < 8 + RG_NPROF) RG_SI(c, lane) = 0;
  RG_NOUNROLL for (int i = lane; i < RG_NSEP; i += 32) ((int*)(s + L.sep))[i] = io.sep ? io.sep[(size_t)env * RG_NSEP + i] : 0xfff;
  RG_PHASE_END
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int j = lane; j < m.njnt; j += 32)
    if (m.jnt_type[j] == RG_JNT_FREE) for (int a = 0; a < 3; a++) s[L.qpos + m.jnt_qposadr[j] + a] -= m.origin[a];
  /* mocap poses: data.mocap_pos / mocap_quat of this environment, or the model pose of the mocap bodies */
  RG_NOUNROLL for (int b = lane; b < m.nbody; b += 32) {
    const int k = m.body_mocapid[b];
    if (k < 0) continue;
    float* mp = s + L.mocap + 7 * k;
    for (int a = 0; a < 3; a++) mp[a] = io.mocap_pos ? io.mocap_pos[((size_t)env * m.nmocap + k) * 3 + a] - m.origin[a] : m.body_pos[3 * b + a];
    for (int a = 0; a < 4; a++) mp[3 + a] = io.mocap_quat ? io.mocap_quat[((size_t)env * m.nmocap + k) * 4 + a] : m.body_quat[4 * b + a];
  }
  RG_PHASE_END
  for (int sub = 0; sub < nsub; sub++) {
    rg_forward(c);
    { RG_PROF_BEGIN RG_SYNC_SMALL(); rg_euler(c); RG_PROF(c, 8) }
    if (m.pidw == 6) {   /* cascaded-PI "a step has been taken" flag (mjpid.pyx: d.time > 0) */
      RG_PHASE_BEGIN
      RG_NOUNROLL for (int i = lane; i < m.nu; i += 32) s[L.pid + 6 * i + 5] = 1.0f;
      RG_PHASE_END
    }
    /* mj_checkPos / mj_checkVel: reset on a bad state, like mj_step does */
    LANEVAR(int, badl);
    RG_PHASE_BEGIN
    int bad = 0;
    RG_NOUNROLL for (int i = lane; i < m.nq; i += 32) { const float v = s[L.qpos + i]; bad |= !(v == v) || fabsf(v) > 1e10f; }
    RG_NOUNROLL for (int i = lane; i < m.nv; i += 32) { const float v = s[L.qvel + i]; bad |= !(v == v) || fabsf(v) > 1e10f; }
    LV(badl) = bad;
    RG_PHASE_END
    if (RG_WARP_OR(badl)) {
      RG_PHASE_BEGIN
      RG_NOUNROLL for (int i = lane; i < m.nq; i += 32) s[L.qpos + i] = m.qpos0[i];
      RG_NOUNROLL for (int i = lane; i < m.nv; i += 32) { s[L.qvel + i] = 0.0f; s[L.warm + i] = 0.0f; }
      RG_NOUNROLL for (int i = lane; i < npid; i += 32) s[L.pid + i] = 0.0f;
      if (lane == 0) RG_SI(c, RG_S_WARN) |= RG_WARN_BAD_STATE;
      RG_PHASE_END
      RG_PHASE_BEGIN
      RG_NOUNROLL for (int j = lane; j < m.njnt; j += 32)
        if (m.jnt_type[j] == RG_JNT_FREE) for (int a = 0; a < 3; a++) s[L.qpos + m.jnt_qposadr[j] + a] -= m.origin[a];
      RG_PHASE_END
    }
  }
  for (int f = 0; f < final_forward; f++) rg_forward(c);   /* robogym steps, then observes: sim.forward() runs again in RobotEnv._observe_sync */
  /* ---- store */
  if (!store) return;
  if (io.sensordata && m.nsensor > 0) rg_sensors(c, io.sensordata + (size_t)env * m.nsensordata);
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int j = lane; j < m.njnt; j += 32)
    if (m.jnt_type[j] == RG_JNT_FREE) for (int a = 0; a < 3; a++) s[L.qpos + m.jnt_qposadr[j] + a] += m.origin[a];
  RG_PHASE_END
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < m.nq; i += 32) io.qpos[(size_t)env * m.nq + i] = s[L.qpos + i];
  RG_NOUNROLL for (int i = lane; i < m.nv; i += 32) { io.qvel[(size_t)env * m.nv + i] = s[L.qvel + i]; io.warm[(size_t)env * m.nv + i] = s[L.warm + i]; }
  RG_NOUNROLL for (int i = lane; i < npid; i += 32) io.pid[(size_t)env * npid + i] = s[L.pid + i];
  if (lane == 0 && io.time) io.time[env] += c.timestep * (float)nsub;
  if (io.sep) RG_NOUNROLL for (int i = lane; i < RG_NSEP; i += 32) io.sep[(size_t)env * RG_NSEP + i] = ((const int*)(s + L.sep))[i];
  if (io.site_xpos) RG_NOUNROLL for (int i = lane; i < 3 * m.nsite; i += 32) io.site_xpos[(size_t)env * 3 * m.nsite + i] = s[L.sxpos + i] + m.origin[i % 3];
  if (io.body_xpos) RG_NOUNROLL for (int i = lane; i < 3 * m.nbody; i += 32) io.body_xpos[(size_t)env * 3 * m.nbody + i] = s[L.xpos + i] + m.origin[i % 3];
  if (io.body_xquat) RG_NOUNROLL for (int i = lane; i < 4 * m.nbody; i += 32) io.body_xquat[(size_t)env * 4 * m.nbody + i] = s[L.xquat + i];
  if (io.geom_xpos) RG_NOUNROLL for (int i = lane; i < 3 * m.ngeom; i += 32) io.geom_xpos[(size_t)env * 3 * m.ngeom + i] = s[L.gxpos + i] + m.origin[i % 3];
  if (io.body_xvel) RG_NOUNROLL for (int b = lane; b < m.nbody; b += 32) {
    /* mj_objectVelocity(body, world axes): spatial velocity about the tree's reference point, moved to the body origin */
    float V[6] = {0, 0, 0, 0, 0, 0};
    RG_NOUNROLL for (int w = 0; w < m.nmaskw; w++) {
      unsigned bits = (unsigned)m.body_dofmask[b * m.nmaskw + w];
      while (bits) {
        const int d = 32 * w + rg_ctz(bits);
        bits &= bits - 1;
        const float qd = s[L.qvel + d];
        const float* Sd = s + L.S + 6 * d;
        for (int k = 0; k < 6; k++) V[k] += Sd[k] * qd;
      }
    }
    float rel[3], t[3];
    rg_sub3(rel, s + L.xpos + 3 * b, rg_body_ref(c, b));
    rg_cross(t, V, rel);
    float* o = io.body_xvel + ((size_t)env * m.nbody + b) * 6;
    o[0] = V[0]; o[1] = V[1]; o[2] = V[2]; o[3] = V[3] + t[0]; o[4] = V[4] + t[1]; o[5] = V[5] + t[2];
  }
  if (io.act_force) RG_NOUNROLL for (int i = lane; i < m.nu; i += 32) io.act_force[(size_t)env * m.nu + i] = s[L.aforce + i];
  if (io.qacc) RG_NOUNROLL for (int i = lane; i < m.nv; i += 32) io.qacc[(size_t)env * m.nv + i] = s[L.qacc + i];
  const int ncon = RG_SI(c, RG_S_NCON);
  if (io.contact) RG_NOUNROLL for (int k = lane; k < L.ncon; k += 32) {
    float* o = io.contact + ((size_t)env * L.ncon + k) * 4;
    const float* r = s + L.con + RG_CON_STRIDE * k;
    if (k < ncon) { o[0] = r[20]; o[1] = r[21]; o[2] = r[0]; o[3] = r[17]; } else { o[0] = o[1] = -1.0f; o[2] = 0.0f; o[3] = 0.0f; }
  }
  if (lane == 0) { if (io.ncon) io.ncon[env] = ncon; if (io.warn) io.warn[env] |= RG_SI(c, RG_S_WARN); if (io.cost) io.cost[env] = RG_SI(c, RG_S_WORK); }
  if (io.dbg) {
    float* g = io.dbg + (size_t)env * rg_dbg_size(m, L.ncon);
    const int nv = m.nv;
    int o = 0;
    RG_NOUNROLL for (int i = lane; i < nv * nv; i += 32) {   /* dense nv x nv from the tree-sparse rows */
      const int r_ = i / nv, c_ = i - r_ * nv;
      const int hi = r_ > c_ ? r_ : c_, lo = r_ > c_ ? c_ : r_;
      const int anc = hi < lo + m.dof_mrow[3 * lo + 1];       /* lo is an ancestor of hi (or hi itself) */
      g[o + i] = anc ? s[L.M + m.dof_mrow[3 * hi] + m.dof_mrow[3 * hi + 2] - m.dof_mrow[3 * lo + 2]] : 0.0f;
    }
    o += nv * nv;
    RG_NOUNROLL for (int i = lane; i < nv; i += 32) {
      g[o + i] = s[L.bias + i]; g[o + nv + i] = 0.0f; g[o + 2 * nv + i] = 0.0f;   /* passive / actuator split is not kept */
      g[o + 3 * nv + i] = s[L.smooth + i]; g[o + 4 * nv + i] = s[L.qacc + i]; g[o + 5 * nv + i] = s[L.qfc + i];
    }
    o += 6 * nv;
    RG_NOUNROLL for (int i = lane; i < m.ntendon; i += 32) g[o + i] = s[L.tlen + i];
    o += m.ntendon;
    RG_NOUNROLL for (int i = lane; i < m.nu; i += 32) { g[o + i] = s[L.alen + i]; g[o + m.nu + i] = s[L.aforce + i]; }
    o += 2 * m.nu;
    if (lane == 0) { g[o] = (float)ncon; g[o + 1] = (float)RG_SI(c, RG_S_NEL); g[o + 2] = (float)RG_SI(c, RG_S_NITER); g[o + 3] = (float)RG_SI(c, RG_S_WARN); }
    o += 4;
    RG_NOUNROLL for (int i = lane; i < L.ncon * RG_CON_STRIDE; i += 32) g[o + i] = s[L.con + i];
    o += L.ncon * RG_CON_STRIDE;
    RG_NOUNROLL for (int i = lane; i < m.ntendon * nv; i += 32) g[o + i] = rg_tendon_J(c, i / nv, i % nv);
    o += m.ntendon * nv;
    RG_NOUNROLL for (int i = lane; i < RG_NPROF; i += 32) g[o + i] = s[L.scal + 8 + i];
  }
  RG_PHASE_END
}

/* ---- mj_setConst for one environment (SimulationInterface.set_constants, robogym/mujoco/simulation_interface.py:197-201, after
 * the randomisers edited masses / inertias / armatures, robogym/wrappers/randomizations.py:72-306): the constants MuJoCo derives
 * from the model at qpos0, recomputed from THIS environment's parameter view and written to its rows of the per-environment
 * arrays (each pointer may be null).  A^-1 = M(qpos0)^-1 by the tree-sparse factorisation:
 *   dof_invweight0[i]   = A^-1(i, i), averaged over the 3 rotations / 3 translations of a ball or free joint
 *   body_invweight0[b]  = tr(Jp A^-1 Jp') / 3, tr(Jr A^-1 Jr') / 3 with the Jacobians at the body's centre of mass (0 for bodies
 *                         welded to the world)
 *   tendon_invweight0[t]= J_t A^-1 J_t',   tendon_length0[t] = length at qpos0
 *   body_subtreemass[b] = mass of b and its descendants,   opt_meaninertia = mean diag M(qpos0) */
struct RgSetConstOut { float* dof_invweight0; float* body_invweight0; float* tendon_invweight0; float* tendon_length0; float* body_subtreemass; float* opt_meaninertia; };

RG_DEV_NOINLINE void rg_env_setconst(RgMRef mr, const RgLayout& L_in, float* s_in, int soff, const RgSetConstOut& o) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(mr);
#ifdef RG_EMU
  const RgCtx c = {mr, &L_in, s_in, nullptr, m.opt_timestep[0]};
#else
  const RgCtx c = {mr, soff, nullptr, m.opt_timestep[0]};
#endif
  const RgLayout& L = RG_CL(c);
  float* s = RG_SCRATCH(c);
  const int nv = m.nv;
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < m.nq; i += 32) s[L.qpos + i] = m.qpos0[i];
  RG_NOUNROLL for (int i = lane; i < nv; i += 32) s[L.qvel + i] = 0.0f;
  if (lane < 8 + RG_NPROF) RG_SI(c, lane) = 0;
  RG_PHASE_END
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int j = lane; j < m.njnt; j += 32)
    if (m.jnt_type[j] == RG_JNT_FREE) for (int a = 0; a < 3; a++) s[L.qpos + m.jnt_qposadr[j] + a] -= m.origin[a];
  RG_NOUNROLL for (int b = lane; b < m.nbody; b += 32) {
    const int k = m.body_mocapid[b];
    if (k < 0) continue;
    for (int a = 0; a < 3; a++) s[L.mocap + 7 * k + a] = m.body_pos[3 * b + a];
    for (int a = 0; a < 4; a++) s[L.mocap + 7 * k + 3 + a] = m.body_quat[4 * b + a];
  }
  RG_PHASE_END
  rg_kinematics(c);
  rg_massmatrix(c);
  if (m.ntendon > 0) rg_tendon(c);
  const int F = L.H, invD = L.Mv, x = L.search, rhs = L.qfc, diag = L.Ma;
  rg_sparse_factor(c, F, invD, nullptr, 0.0f, m.dof_lvl + 0);
  /* diagonal of the inverse, one unit vector at a time */
  for (int i = 0; i < nv; i++) {
    RG_PHASE_BEGIN
    RG_NOUNROLL for (int d = lane; d < nv; d += 32) s[x + d] = d == i ? 1.0f : 0.0f;
    RG_PHASE_END
    rg_sparse_solve(c, F, invD, x, m.dof_lvl + 0);
    RG_PHASE_BEGIN
    if (lane == 0) s[diag + i] = s[x + i];
    RG_PHASE_END
  }
  RG_PHASE_BEGIN
  if (o.dof_invweight0) RG_NOUNROLL for (int d = lane; d < nv; d += 32) {
    const int j = m.dof_jntid[d], type = m.jnt_type[j], a = m.jnt_dofadr[j];
    float v = s[diag + d];
    if (type == RG_JNT_BALL || type == RG_JNT_FREE) {
      const int g = a + 3 * ((d - a) / 3);
      v = (s[diag + g] + s[diag + g + 1] + s[diag + g + 2]) * (1.0f / 3.0f);
    }
    o.dof_invweight0[d] = v;
  }
  if (o.body_subtreemass) RG_NOUNROLL for (int b = lane; b < m.nbody; b += 32) {
    float acc = 0.0f;
    RG_NOUNROLL for (int k = b; k < b + m.body_subtreesize[b]; k++) acc += m.body_mass[k];
    o.body_subtreemass[b] = acc;
  }
  if (o.tendon_length0) RG_NOUNROLL for (int t = lane; t < m.ntendon; t += 32) o.tendon_length0[t] = s[L.tlen + t];
  RG_PHASE_END
  if (o.opt_meaninertia) {
    LANEVAR(float, part);
    RG_PHASE_BEGIN
    float a = 0.0f;
    RG_NOUNROLL for (int d = lane; d < nv; d += 32) a += s[L.M + m.dof_mrow[3 * d]];
    LV(part) = a;
    RG_PHASE_END
    const float tot = RG_WARP_SUM(part);
    RG_PHASE_BEGIN
    if (lane == 0) o.opt_meaninertia[0] = tot / (float)(nv > 0 ? nv : 1);
    RG_PHASE_END
  }
  /* rows of the body Jacobians and of the tendon Jacobian: y = A^-1 r, then r . y */
  const int nrow = (o.body_invweight0 ? 6 * m.nbody : 0) + (o.tendon_invweight0 ? m.ntendon : 0);
  float tr = 0.0f;
  for (int q = 0; q < nrow; q++) {
    const int isbody = o.body_invweight0 && q < 6 * m.nbody;
    const int b = isbody ? q / 6 : 0, a = isbody ? q - 6 * b : 0, t = isbody ? 0 : q - (o.body_invweight0 ? 6 * m.nbody : 0);
    const int skip = isbody && (b == 0 || m.body_weldid[b] == 0);
    if (!skip) {
      LANEVAR(float, part);
      RG_PHASE_BEGIN
      float com[3] = {0, 0, 0};
      if (isbody) rg_body_xipos(c, b, com);
      RG_NOUNROLL for (int d = lane; d < nv; d += 32) {
        float v = 0.0f;
        if (!isbody) v = rg_tendon_J(c, t, d);
        else if (rg_dof_in_body(m, b, d)) {
          if (a < 3) { float jp[3]; rg_jacp_world(c, d, com, jp); v = jp[a]; }
          else v = s[L.S + 6 * d + a - 3];
        }
        s[rhs + d] = v; s[x + d] = v;
      }
      RG_PHASE_END
      rg_sparse_solve(c, F, invD, x, m.dof_lvl + 0);
      RG_PHASE_BEGIN
      float acc = 0.0f;
      RG_NOUNROLL for (int d = lane; d < nv; d += 32) acc += s[rhs + d] * s[x + d];
      LV(part) = acc;
      RG_PHASE_END
      tr += RG_WARP_SUM(part);
    }
    if (!isbody || a == 2 || a == 5) {
      RG_PHASE_BEGIN
      if (lane == 0) {
        if (isbody) o.body_invweight0[2 * b + (a == 5)] = skip ? 0.0f : fmaxf(tr * (1.0f / 3.0f), RG_MINVAL);
        else o.tendon_invweight0[t] = fmaxf(tr, RG_MINVAL);
      }
      RG_PHASE_END
      tr = 0.0f;
    }
  }
}
