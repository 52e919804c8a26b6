/* rg_sol.inl -- S8/S9/S13/S15 of the fused step: constraint elements (dof friction loss, joint and
 * tendon limits, pyramidal frictional contacts) with solref/solimp impedance, Newton solver with
 * exact line search on the primal convex problem, semi-implicit Euler with implicit joint damping.
 * Replaces mj_makeConstraint / mj_fwdConstraint / mj_Euler inside the mj_step behind
 * sim.step() (robogym/mujoco/simulation_interface.py:184; solver options from
 * robogym/assets/xmls/robot/shadowhand/assets.xml:14-15).
 *
 * Jacobians are never materialised as nefc x nv: single-dof rows are handled by index, tendon
 * rows reuse the tendon Jacobian, contact rows are rebuilt from the motion axes S on the fly.
 */
#pragma once
#include "rg_col.inl"

#ifndef RG_NEWTON_TOL
#define RG_NEWTON_TOL 1e-6f   /* fp32: below ~1e-6 the scaled cost differences are rounding noise */
#endif
#ifndef RG_LS_TOL
#define RG_LS_TOL 1e-6f
#endif
#define RG_MINIMP 0.0001f
#define RG_MAXIMP 0.9999f

RG_DEV float rg_impedance(const float* solimp, float xabs) {
  const float dmin = rg_clamp(solimp[0], RG_MINIMP, RG_MAXIMP), dmax = rg_clamp(solimp[1], RG_MINIMP, RG_MAXIMP);
  const float width = fmaxf(1e-12f, solimp[2]), mid = rg_clamp(solimp[3], RG_MINIMP, RG_MAXIMP), power = fmaxf(1.0f, solimp[4]);
  const float x = xabs / width;
  if (x >= 1.0f) return dmax;
  if (x <= 0.0f) return dmin;
  float y;
  if (power == 1.0f) y = x;
  else if (x <= mid) y = powf(x, power) / powf(mid, power - 1.0f);
  else y = 1.0f - powf(1.0f - x, power) / powf(1.0f - mid, power - 1.0f);
  return dmin + y * (dmax - dmin);
}
/* R (regulariser) and aref for one row; K is dropped for friction rows */
RG_DEV_NOINLINE void rg_row_params(const RgCtx c, const float* solref_in, const float* solimp, float pos, float margin, float vel, float diagApprox,
                          int isfriction, float* R, float* aref, float* Bout, float* KIout) {
  float sr0 = solref_in[0];
  const float sr1 = solref_in[1];
  if (!(RG_MDEREF(c.mref).opt_disableflags[0] & RG_DSBL_REFSAFE) && sr0 > 0) sr0 = fmaxf(sr0, 2.0f * c.timestep);
  const float p = pos - margin;
  const float imp = rg_impedance(solimp, fabsf(p));
  const float dmax = rg_clamp(solimp[1], RG_MINIMP, RG_MAXIMP);
  *R = fmaxf(1e-12f, (1.0f - imp) * diagApprox / imp);
  float K, B;
  if (sr0 > 0) { K = 1.0f / fmaxf(1e-12f, dmax * dmax * sr0 * sr0 * sr1 * sr1); B = 2.0f / fmaxf(1e-12f, dmax * sr0); }
  else { K = -sr0 / fmaxf(1e-12f, dmax * dmax); B = -sr1 / fmaxf(1e-12f, dmax); }
  if (isfriction) K = 0.0f;
  *aref = -B * vel - K * imp * p;
  *Bout = B;
  *KIout = K * imp * p;
}

/* contact-frame Jacobian column of dof d for contact record r (dim components), sign included; 0 if untouched */
RG_DEV_NOINLINE int rg_contact_col(const RgCtx c, const float* r, int d, int dim, float* col) {
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  const int b1 = (int)r[18], b2 = (int)r[19];
  const int in1 = rg_dof_in_body(m, b1, d), in2 = rg_dof_in_body(m, b2, d);
  if (in1 == in2) return 0;
  const float* S = RG_SCRATCH(c) + RG_CL(c).S + 6 * d;
  float jp[3];
  rg_jacp_world(c, d, r + 1, jp);
  const float sg = in2 ? 1.0f : -1.0f;
  col[0] = sg * rg_dot3(r + 4, jp);
  if (dim > 1) { col[1] = sg * rg_dot3(r + 7, jp); col[2] = sg * rg_dot3(r + 10, jp); }
  if (dim > 3) col[3] = sg * rg_dot3(r + 4, S);
  if (dim > 4) { col[4] = sg * rg_dot3(r + 7, S); col[5] = sg * rg_dot3(r + 10, S); }
  return 1;
}
/* same, for a dof already known to be touched (sg = +1 if it moves body 2, -1 if body 1) */
RG_DEV void rg_contact_col_list(const RgCtx c, const float* r, int d, float sg, int dim, float* col) {
  const float* S = RG_SCRATCH(c) + RG_CL(c).S + 6 * d;
  float jp[3];
  rg_jacp_world(c, d, r + 1, jp);
  col[0] = sg * rg_dot3(r + 4, jp);
  if (dim > 1) { col[1] = sg * rg_dot3(r + 7, jp); col[2] = sg * rg_dot3(r + 10, jp); }
  if (dim > 3) col[3] = sg * rg_dot3(r + 4, S);
  if (dim > 4) { col[4] = sg * rg_dot3(r + 7, S); col[5] = sg * rg_dot3(r + 10, S); }
}
RG_DEV float rg_contact_mu(const float* r, int k) { return k <= 2 ? r[14] : (k == 3 ? r[15] : r[16]); }

/* ---- elliptic friction cones (option cone="elliptic", robogym/assets/xmls/robot/ur16e/base.xml:4).  Kept out of line: the
 * pyramidal models (dactyl) must not pay for this code in their Newton loop's instruction footprint.  One contact has dim
 * rows (normal + friction directions) with D_0 = 1/R_normal and D_k = D_0 friction_k^2 / mu^2, mu = friction_0 / sqrt(impratio).
 * With N = mu u_0, U_k = friction_k u_k, T = |U| (coordinates in which the force cone is circular with coefficient mu) the
 * cost has three zones: top (N >= mu T: nothing), bottom (mu N + T <= 0: every row quadratic) and the middle zone
 * 1/2 Dm (N - mu T)^2, Dm = D_0 / (mu^2 (1 + mu^2)) = half the squared length of the projection onto the cone's boundary. */
RG_DEV float rg_cone_mu(const RG_MODEL_T& m, const float* r) { return fmaxf(r[14] * sqrtf(1.0f / m.opt_impratio[0]), 1e-5f); }
/* zone (0 top, 1 bottom, 2 middle), cost and force F = -gradient at u */
RG_DEV_NOINLINE int rg_cone_eval(const float* r, float mu, int dim, float D0, const float* u, float* cost, float* F) {
  float U[6] = {0, 0, 0, 0, 0, 0}, T2 = 0.0f;
  const float N = u[0] * mu;
  for (int k = 1; k < 6; k++) if (k < dim) { U[k] = u[k] * rg_contact_mu(r, k); T2 += U[k] * U[k]; }
  const float T = sqrtf(T2);
  for (int k = 0; k < 6; k++) F[k] = 0.0f;
  *cost = 0.0f;
  if (N >= mu * T) return 0;
  if (mu * N + T <= 0.0f) {
    float cst = 0.0f;
    for (int k = 0; k < 6; k++) if (k < dim) {
      const float fk = k == 0 ? mu : rg_contact_mu(r, k);
      const float Dk = D0 * fk * fk / (mu * mu);
      F[k] = -Dk * u[k]; cst += 0.5f * Dk * u[k] * u[k];
    }
    *cost = cst;
    return 1;
  }
  const float Dm = D0 / (mu * mu * (1.0f + mu * mu)), NT = N - mu * T;
  *cost = 0.5f * Dm * NT * NT;
  F[0] = -Dm * NT * mu;
  const float q = Dm * NT * mu / T;
  for (int k = 1; k < 6; k++) if (k < dim) F[k] = q * rg_contact_mu(r, k) * U[k];
  return 2;
}
/* first and second derivative of the cone cost along u + alpha w */
RG_DEV_NOINLINE void rg_cone_ray(const float* r, float mu, int dim, float D0, const float* u, const float* w, float alpha, float* g, float* h) {
  float U[6], V[6], T2 = 0.0f, UV = 0.0f, VV = 0.0f;
  const float N = (u[0] + alpha * w[0]) * mu, Nd = w[0] * mu;
  for (int k = 1; k < 6; k++) if (k < dim) {
    const float fk = rg_contact_mu(r, k);
    U[k] = (u[k] + alpha * w[k]) * fk; V[k] = w[k] * fk;
    T2 += U[k] * U[k]; UV += U[k] * V[k]; VV += V[k] * V[k];
  }
  const float T = sqrtf(T2);
  if (N >= mu * T) return;
  if (mu * N + T <= 0.0f) {
    for (int k = 0; k < 6; k++) if (k < dim) {
      const float fk = k == 0 ? mu : rg_contact_mu(r, k);
      const float Dk = D0 * fk * fk / (mu * mu);
      *g += Dk * (u[k] + alpha * w[k]) * w[k]; *h += Dk * w[k] * w[k];
    }
    return;
  }
  const float Dm = D0 / (mu * mu * (1.0f + mu * mu)), NT = N - mu * T;
  const float Td = UV / T, Tdd = (VV - Td * Td) / T;
  const float NTd = Nd - mu * Td;
  *g += Dm * NT * NTd;
  *h += Dm * (NTd * NTd - NT * mu * Tdd);
}
/* dim x dim Hessian block of the middle zone at u (row-major W[6][6], only [0, dim) x [0, dim) is written) */
RG_DEV_NOINLINE void rg_cone_hessian(const float* r, float mu, int dim, float D0, const float* u, float W[6][6]) {
  float U[6] = {0, 0, 0, 0, 0, 0}, f[6] = {0, 0, 0, 0, 0, 0}, T2 = 0.0f;
  const float N = u[0] * mu;
  for (int k = 1; k < 6; k++) if (k < dim) { f[k] = rg_contact_mu(r, k); U[k] = u[k] * f[k]; T2 += U[k] * U[k]; }
  const float T = sqrtf(T2), iT = 1.0f / T;
  const float Dm = D0 / (mu * mu * (1.0f + mu * mu)), NT = N - mu * T;
  W[0][0] = Dm * mu * mu;
  for (int k = 1; k < 6; k++) if (k < dim) {
    W[0][k] = W[k][0] = -Dm * mu * mu * f[k] * U[k] * iT;
    for (int l = 1; l < 6; l++) if (l < dim)
      W[k][l] = Dm * mu * f[k] * f[l] * (mu * U[k] * U[l] * iT * iT - NT * ((k == l ? iT : 0.0f) - U[k] * U[l] * iT * iT * iT));
  }
}

/* y = M x with the tree-sparse M: row i couples dof i with its ancestors (stored in row i) and with the dofs of its
   subtree, which are the dofs right behind it (entry (d, i) sits depth(d) - depth(i) into row d) */
RG_DEV_NOINLINE void rg_matvec_phase(const RgCtx c, int y, int x) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  const int nv = m.nv;
  float* s = RG_SCRATCH(c);
  const float* M = s + RG_CL(c).M;
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < nv; i += 32) {
    const int adr = m.dof_mrow[3 * i], nsub = m.dof_mrow[3 * i + 1], dep = m.dof_mrow[3 * i + 2];
    float acc = 0.0f;
    int j = i;
    RG_NOUNROLL for (int k = 0; k <= dep; k++) { acc += M[adr + k] * s[x + j]; j = m.dof_parentid[j]; }
    RG_NOUNROLL for (int d = i + 1; d < i + nsub; d++) acc += M[m.dof_mrow[3 * d] + m.dof_mrow[3 * d + 2] - dep] * s[x + d];
    s[y + i] = acc;
  }
  RG_PHASE_END
}

/* dense H (packed lower triangle over the solver's dofs, leaves first) <- tree-sparse M */
RG_DEV void rg_H_from_M(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  const RgLayout& L = RG_CL(c);
  const int nv = m.nv, ns = m.ns;
  float* s = RG_SCRATCH(c);
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < ((ns * (ns + 1)) >> 1); i += 32) s[L.H + i] = 0.0f;
  RG_PHASE_END
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < nv; i += 32) {
    const int si = m.dof_sidx[i];
    if (si < 0) continue;
    const float* row = s + L.M + m.dof_mrow[3 * i];
    int j = i;
    RG_NOUNROLL for (int k = 0; j >= 0; k++) {
      s[L.H + RG_TRI(m.dof_sidx[j], si)] = row[k];     /* an ancestor sits later in the solver's order */
      j = m.dof_parentid[j];
    }
  }
  RG_PHASE_END
}

/* in-place envelope Cholesky of the lower triangle of A (packed rows); env[i] = first nonzero column of row i.
 * Row n (one past the matrix) carries the right-hand side: treating it as one more row of the matrix makes the factorisation
 * deliver the forward substitution y = L^-1 b in that row for free (the same dot products, one more lane). */
RG_DEV_NOINLINE void rg_cholesky(const RgCtx c, int A, const int* env) {
  RG_LANE_DECL
  const int n = RG_MDEREF(c.mref).ns;
  float* s = RG_SCRATCH(c);
  for (int j = 0; j < n; j++) {
    LANEVAR(float, sumv);
    RG_PHASE_BEGIN
    const int i = j + lane;
    float acc = 0.0f;
    const int ei = i < n ? env[i] : 0;
    if (i <= n && ei <= j) {
      const float* ri = s + A + RG_TRI(i, 0); const float* rj = s + A + RG_TRI(j, 0);
      acc = ri[j];
      float acc1 = 0.0f;
      int k = ei > env[j] ? ei : env[j];
      RG_UNROLL2 for (; k + 1 < j; k += 2) { acc -= ri[k] * rj[k]; acc1 -= ri[k + 1] * rj[k + 1]; }
      if (k < j) acc -= ri[k] * rj[k];
      acc += acc1;
    }
    LV(sumv) = acc;
    RG_PHASE_END
    /* the diagonal slot keeps 1/L[j][j]: the factor is only ever used to solve (rg_chol_back) */
    const float inv = RG_RSQRT(fmaxf(RG_WARP_BCAST(sumv, 0), 1e-20f));
    RG_PHASE_BEGIN
    const int i = j + lane;
    if (i == j) s[A + RG_TRI(j, j)] = inv;
    else if (i <= n) s[A + RG_TRI(i, j)] = LV(sumv) * inv;
    RG_NOUNROLL for (int i2 = i + 32; i2 <= n; i2 += 32) {
      float acc = 0.0f;
      const int e2 = i2 < n ? env[i2] : 0;
      if (e2 <= j) {
        const float* ri = s + A + RG_TRI(i2, 0); const float* rj = s + A + RG_TRI(j, 0);
        acc = ri[j];
        float acc1 = 0.0f;
        int k = e2 > env[j] ? e2 : env[j];
        RG_UNROLL2 for (; k + 1 < j; k += 2) { acc -= ri[k] * rj[k]; acc1 -= ri[k + 1] * rj[k + 1]; }
        if (k < j) acc -= ri[k] * rj[k];
        acc += acc1;
      }
      s[A + RG_TRI(i2, j)] = acc * inv;
    }
    RG_PHASE_END
  }
}
/* y <- L^-1 y for the right-hand side sitting in row n of A (only needed when a factor is REUSED: rg_cholesky does it
   on the fly) */
RG_DEV_NOINLINE void rg_chol_forward(const RgCtx c, int A, const int* env) {
  RG_LANE_DECL
  const int n = RG_MDEREF(c.mref).ns;
  float* s = RG_SCRATCH(c);
  const int x = A + RG_TRI(n, 0);
  for (int j = 0; j < n; j++) {
    RG_PHASE_BEGIN
    const float xj = s[x + j] * s[A + RG_TRI(j, j)];   /* diagonal slot = 1/L[j][j] */
    RG_NOUNROLL for (int i = j + 1 + lane; i < n; i += 32)
      if (env[i] <= j) s[x + i] -= s[A + RG_TRI(i, j)] * xj;
    RG_PHASE_END
    RG_PHASE_BEGIN   /* only now: the other lanes have read the old value */
    if (lane == 0) s[x + j] *= s[A + RG_TRI(j, j)];
    RG_PHASE_END
  }
}
/* out[dof at solver position j] <- (L^-T y)[j] with y in row n of A (row n is consumed): the solution leaves in model dof order */
RG_DEV_NOINLINE void rg_chol_back(const RgCtx c, int A, const int* env, int out) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  const int n = m.ns;
  const int* sdof = m.dof_sidx + m.nv;
  float* s = RG_SCRATCH(c);
  const int x = A + RG_TRI(n, 0);
  for (int j = n - 1; j >= 0; j--) {
    RG_PHASE_BEGIN
    const float xj = s[x + j] * s[A + RG_TRI(j, j)];
    if (lane == 0) s[out + sdof[j]] = xj;
    RG_NOUNROLL for (int i = env[j] + lane; i < j; i += 32) s[x + i] -= s[A + RG_TRI(j, i)] * xj;
    RG_PHASE_END
  }
}

/* ---------------------------------------------------------------- tree-sparse factorisation of M + diag */
/* M couples a dof only with its ancestors and descendants, so M + diag = L' D L with L as sparse as M (the layout MuJoCo
 * calls qLD).  F (nM floats, rows like M) receives D(i) in slot 0 of row i and the UNSCALED couplings D(i) L(i, a) behind
 * it; invD (nv floats) the reciprocals of D.  Work is organised by depth level, leaves first: entry (i, a) subtracts the
 * contributions of the dofs in the subtree of i, all of which are deeper and therefore final ("pull" form: lanes never
 * write the same slot, no atomics, fixed summation order). */
RG_DEV_NOINLINE void rg_sparse_factor(const RgCtx c, int F, int invD, const float* diag, float scale, const int* lvl) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  float* s = RG_SCRATCH(c);
  const float* M = s + RG_CL(c).M;
  const int nv = m.nv;
  const int* order = lvl;          /* dof ids by depth level (all dofs, or the trees outside the constraint solver) */
  const int* start = lvl + nv;
  for (int lvl = m.ndoflevel - 1; lvl >= 0; lvl--) {
    const int l0 = start[lvl], cnt = start[lvl + 1] - l0, w = lvl + 1;
    RG_PHASE_BEGIN
    RG_NOUNROLL for (int it = lane; it < cnt * w; it += 32) {
      const int q = it / w, e = it - q * w;
      const int i = order[l0 + q];
      const int adr = m.dof_mrow[3 * i], nsub = m.dof_mrow[3 * i + 1];
      float acc = M[adr + e] + (e == 0 && diag ? scale * diag[i] : 0.0f);
      RG_NOUNROLL for (int k = i + 1; k < i + nsub; k++) {
        const int ak = m.dof_mrow[3 * k], dk = m.dof_mrow[3 * k + 2] - lvl;   /* column i sits dk entries into row k */
        acc -= s[F + ak + dk] * s[F + ak + dk + e] * s[invD + k];
      }
      s[F + adr + e] = acc;
      if (e == 0) s[invD + i] = 1.0f / acc;
    }
    RG_PHASE_END
  }
}
/* x <- (M + diag)^-1 x with the factor above */
RG_DEV_NOINLINE void rg_sparse_solve(const RgCtx c, int F, int invD, int x, const int* lvl) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  float* s = RG_SCRATCH(c);
  const int nv = m.nv;
  const int* order = lvl;          /* dof ids by depth level (all dofs, or the trees outside the constraint solver) */
  const int* start = lvl + nv;
  /* x <- L^-T x: a dof collects from its subtree (deeper levels are final) */
  for (int lvl = m.ndoflevel - 2; lvl >= 0; lvl--) {
    RG_PHASE_BEGIN
    RG_NOUNROLL for (int q = start[lvl] + lane; q < start[lvl + 1]; q += 32) {
      const int i = order[q];
      const int nsub = m.dof_mrow[3 * i + 1];
      float acc = s[x + i];
      RG_NOUNROLL for (int k = i + 1; k < i + nsub; k++)
        acc -= s[F + m.dof_mrow[3 * k] + m.dof_mrow[3 * k + 2] - lvl] * s[invD + k] * s[x + k];
      s[x + i] = acc;
    }
    RG_PHASE_END
  }
  /* x <- D^-1 x, then x <- L^-1 x: a dof collects from its ancestors (shallower levels are final) */
  for (int lvl = 0; lvl < m.ndoflevel; lvl++) {
    RG_PHASE_BEGIN
    RG_NOUNROLL for (int q = start[lvl] + lane; q < start[lvl + 1]; q += 32) {
      const int i = order[q];
      const int adr = m.dof_mrow[3 * i];
      float acc = s[x + i];
      int a = m.dof_parentid[i];
      RG_NOUNROLL for (int e = 1; e <= lvl; e++) { acc -= s[F + adr + e] * s[x + a]; a = m.dof_parentid[a]; }
      s[x + i] = acc * s[invD + i];
    }
    RG_PHASE_END
  }
}

/* ---------------------------------------------------------------- S8/S9 constraint elements */
RG_DEV_NOINLINE void rg_make_constraints(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int nv = m.nv, flags = m.opt_disableflags[0];
  int* el_i = (int*)(s + L.el_i);
  int* eldof = (int*)(s + L.eldof);
  int nel = 0, warn = 0;
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int i = lane; i < nv; i += 32) eldof[i] = 0x3fffffff;   /* three empty 10-bit slots */
  RG_PHASE_END
  const int on = !(flags & RG_DSBL_CONSTRAINT);
  /* dof friction loss: one element per dof with frictionloss > 0 */
  for (int base = 0; on && !(flags & RG_DSBL_FRICTIONLOSS) && base < nv; base += 32) {
    LANEVAR(int, cnt); LANEVAR(int, pos);
    int tot;
    RG_PHASE_BEGIN
    const int d = base + lane;
    LV(cnt) = (d < nv && m.dof_frictionloss[d] > 0.0f) ? 1 : 0;
    RG_PHASE_END
    RG_WARP_SCAN(cnt, pos, tot);
    RG_PHASE_BEGIN
    const int d = base + lane;
    const int e = nel + LV(pos);
    if (LV(cnt) && e < L.nel) {
      float R, aref, B, KI;
      rg_row_params(c, m.dof_solref + 2 * d, m.dof_solimp + 5 * d, 0.0f, 0.0f, s[L.qvel + d], m.dof_invweight0[d], 1, &R, &aref, &B, &KI);
      el_i[e] = RG_EL_FLOSS + 8 * d;
      s[L.el_D + e] = 1.0f / R; s[L.el_jar + e] = -aref;
      eldof[d] = (eldof[d] & ~0x3ff) | e;
    }
    RG_PHASE_END
    nel += tot;
  }
  /* joint limits (hinge / slide) */
  for (int base = 0; on && !(flags & RG_DSBL_LIMIT) && base < m.njnt; base += 32) {
    LANEVAR(int, cnt); LANEVAR(int, pos);
    int tot;
    RG_PHASE_BEGIN
    const int j = base + lane;
    int cn = 0;
    if (j < m.njnt && m.jnt_limited[j] && (m.jnt_type[j] == RG_JNT_SLIDE || m.jnt_type[j] == RG_JNT_HINGE)) {
      const float q = s[L.qpos + m.jnt_qposadr[j]];
      if (q - m.jnt_range[2 * j] < m.jnt_margin[j]) cn++;
      if (m.jnt_range[2 * j + 1] - q < m.jnt_margin[j]) cn++;
    }
    LV(cnt) = cn;
    RG_PHASE_END
    RG_WARP_SCAN(cnt, pos, tot);
    RG_PHASE_BEGIN
    const int j = base + lane;
    if (LV(cnt)) {
      const int d = m.jnt_dofadr[j];
      const float q = s[L.qpos + m.jnt_qposadr[j]];
      int e = nel + LV(pos);
      for (int side = 0; side < 2; side++) {
        const float dist = side == 0 ? q - m.jnt_range[2 * j] : m.jnt_range[2 * j + 1] - q;
        if (!(dist < m.jnt_margin[j]) || e >= L.nel) continue;
        const float sg = side == 0 ? 1.0f : -1.0f;
        float R, aref, B, KI;
        rg_row_params(c, m.jnt_solref + 2 * j, m.jnt_solimp + 5 * j, dist, m.jnt_margin[j], sg * s[L.qvel + d], m.dof_invweight0[d], 0, &R, &aref, &B, &KI);
        el_i[e] = RG_EL_JLIMIT + 4 * side + 8 * d;
        s[L.el_D + e] = 1.0f / R; s[L.el_jar + e] = -aref;
        eldof[d] = (eldof[d] & ~(0x3ff << (10 + 10 * side))) | (e << (10 + 10 * side));
        e++;
      }
    }
    RG_PHASE_END
    nel += tot;
  }
  const int tl0 = nel < L.nel ? nel : L.nel;
  /* tendon limits, then the equality rows: every row of a weld / joint coupling is a "virtual tendon" (rg_equality) whose
     residual must be zero -- the quadratic penalty 1/2 D jar^2 of an always-active row is exactly the sum of the two one-sided
     penalties of a lower and an upper limit at the same place, so an equality row is a pair of limit rows that are both there */
  const int nt = m.ntendon, nvt = nt + m.neqrow;
  for (int base = 0; on && base < nvt; base += 32) {
    LANEVAR(int, cnt); LANEVAR(int, pos);
    int tot;
    RG_PHASE_BEGIN
    const int t = base + lane;
    int cn = 0;
    if (t < nt) {
      if (m.tendon_limited[t] && !(flags & RG_DSBL_LIMIT)) {
        const float len = s[L.tlen + t];
        if (len - m.tendon_range[2 * t] < m.tendon_margin[t]) cn++;
        if (m.tendon_range[2 * t + 1] - len < m.tendon_margin[t]) cn++;
      }
    } else if (t < nvt && !(flags & RG_DSBL_EQUALITY) && m.eq_active[m.eqrow[t - nt] >> 3]) cn = 2;
    LV(cnt) = cn;
    RG_PHASE_END
    RG_WARP_SCAN(cnt, pos, tot);
    RG_PHASE_BEGIN
    const int t = base + lane;
    if (LV(cnt)) {
      const float len = s[L.tlen + t];
      int e = nel + LV(pos);
      const float* solref = m.tendon_solref_lim + 2 * (t < nt ? t : 0);
      const float* solimp = m.tendon_solimp_lim + 5 * (t < nt ? t : 0);
      float lo = 0.0f, hi = 0.0f, margin = 0.0f, invw = 0.0f;
      if (t < nt) { lo = m.tendon_range[2 * t]; hi = m.tendon_range[2 * t + 1]; margin = m.tendon_margin[t]; invw = m.tendon_invweight0[t]; }
      else {
        const int q = m.eqrow[t - nt] >> 3, k = m.eqrow[t - nt] & 7;
        solref = m.eq_solref + 2 * q; solimp = m.eq_solimp + 5 * q;
        if (m.eq_type[q] == RG_EQ_WELD) {
          const int b1 = m.eq_obj1id[q], b2 = m.eq_obj2id[q], o = k < 3 ? 0 : 1;
          invw = m.body_invweight0[2 * b1 + o] + m.body_invweight0[2 * b2 + o];
        } else {
          invw = m.dof_invweight0[m.jnt_dofadr[m.eq_obj1id[q]]];
          if (m.eq_obj2id[q] >= 0) invw += m.dof_invweight0[m.jnt_dofadr[m.eq_obj2id[q]]];
        }
      }
      for (int side = 0; side < 2; side++) {
        const float dist = side == 0 ? len - lo : hi - len;
        if ((t < nt && !(dist < margin)) || e >= L.nel) continue;
        const float sg = side == 0 ? 1.0f : -1.0f;
        float R, aref, B, KI;
        rg_row_params(c, solref, solimp, dist, margin, sg * s[L.tvel + t], invw, 0, &R, &aref, &B, &KI);
        el_i[e] = RG_EL_TLIMIT + 4 * side + 8 * t;
        s[L.el_D + e] = 1.0f / R; s[L.el_jar + e] = -aref;
        e++;
      }
    }
    RG_PHASE_END
    nel += tot;
  }
  if (nel > L.nel) { nel = L.nel; warn |= RG_WARN_ROWS_FULL; }
  /* contacts: per-contact solver parameters; cu = B * (Jc qvel) + [K imp (dist-margin)] on the normal row */
  const int ncon = RG_SI(c, RG_S_NCON);
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int k = lane; k < ncon; k += 32) {
    float* r = s + L.con + RG_CON_STRIDE * k;
    float* prm = s + L.cprm + RG_CPRM * k;
    int dim = (int)r[17];
    if ((flags & (RG_DSBL_CONTACT | RG_DSBL_CONSTRAINT)) || !(r[0] < r[13])) dim = 0; /* inactive: outside includemargin */
    const int b1 = (int)r[18], b2 = (int)r[19];
    const float tran = m.body_invweight0[2 * b1] + m.body_invweight0[2 * b2];
    float R, aref, B, KI;
    /* first row's diagApprox: tran + mu^2 tran (dim>1) or tran */
    const float mu0 = r[14];
    const float solimp[5] = {prm[0], prm[1], prm[2], prm[3], prm[4]};   /* staged there by rg_collision */
    const int elliptic = m.opt_cone[0] == 1;
    rg_row_params(c, r + 22, solimp, r[0], r[13], 0.0f, (dim > 1 && !elliptic) ? tran + mu0 * mu0 * tran : tran, 0, &R, &aref, &B, &KI);
    if (dim > 1 && !elliptic) {   /* pyramid edges share R = 2 mu^2 R_normal; an elliptic contact keeps D_0 = 1 / R_normal here */
      float mu = mu0 * sqrtf(1.0f / m.opt_impratio[0]);
      if (mu < 1e-5f) mu = 1e-5f;
      R = fmaxf(1e-12f, 2.0f * mu * mu * R);
    }
    prm[0] = 1.0f / R; prm[1] = (float)dim; prm[2] = B; prm[3] = KI;
    /* list of the dofs this contact touches (symmetric difference of the two bodies' ancestor sets) */
    unsigned char* list = (unsigned char*)(s + L.cdof + ((L.tile + 3) >> 2) * k);
    int nd = 0;
    unsigned sgn = 0u;
    RG_NOUNROLL for (int w = 0; w < m.nmaskw && dim > 0; w++) {
      const unsigned m1 = (unsigned)m.body_dofmask[b1 * m.nmaskw + w], m2 = (unsigned)m.body_dofmask[b2 * m.nmaskw + w];
      unsigned bits = m1 ^ m2;
      while (bits) {
        const int bit = rg_ctz(bits);
        bits &= bits - 1;
        if (nd < L.tile) { list[nd] = (unsigned char)(32 * w + bit); if ((m2 >> bit) & 1u) sgn |= 1u << nd; nd++; }
        else { dim = 0; RG_SI(c, RG_S_WARN) |= RG_WARN_DOFS_FULL; } /* touches more dofs than the batch's per-contact capacity: dropped, flagged */
      }
    }
    if (dim == 0) { prm[1] = 0.0f; nd = 0; }
    prm[4] = (float)nd; prm[5] = rg_i2f((int)sgn);   /* sign bits travel as raw bits */
    /* velocity part of jar */
    float v[6] = {0, 0, 0, 0, 0, 0};
    RG_NOUNROLL for (int i = 0; i < nd; i++) {
      float col[6];
      const int d = list[i];
      rg_contact_col_list(c, r, d, (sgn >> i) & 1u ? 1.0f : -1.0f, dim, col);
      const float qd = s[L.qvel + d];
      RG_NOUNROLL for (int a = 0; a < dim; a++) v[a] += col[a] * qd;
    }
    float* cb = s + L.cF + 6 * k; /* staged here until the solver initialises cu */
    for (int a = 0; a < 6; a++) cb[a] = B * v[a];
    cb[0] += KI;
  }
  RG_PHASE_END
  RG_PHASE_BEGIN
  if (lane == 0) { RG_SI(c, RG_S_NEL) = nel; RG_SI(c, RG_S_WARN) |= warn; RG_SI(c, RG_S_TL0) = tl0; }
  RG_PHASE_END
}

/* ---------------------------------------------------------------- S13 Newton solver */
/* jar of single-row element e for a candidate acceleration vector at offset x */
RG_DEV float rg_el_Jx(const RgCtx c, int code, int x) {
  const int type = code & 3, side = (code >> 2) & 1, id = code >> 3;
  const float* s = RG_SCRATCH(c);
  if (type == RG_EL_FLOSS) return s[x + id];
  if (type == RG_EL_JLIMIT) return side ? -s[x + id] : s[x + id];
  float acc = 0.0f;
  const int n = ((const int*)(s + RG_CL(c).tJn))[id];
  const unsigned char* ji = (const unsigned char*)(s + RG_CL(c).tJi) + RG_TJ * id;
  RG_NOUNROLL for (int k = 0; k < n; k++) acc += s[RG_CL(c).tJv + RG_TJ * id + k] * s[x + ji[k]];
  return side ? -acc : acc;
}

/* well-mixed per-row hash: the active-set signature is a SUM of these, so it must not be linear in the row id */
RG_DEV unsigned rg_mix(unsigned h) { h *= 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16; return h; }
/* forces + cost at the current jar (el_jar, cu); returns total constraint cost; fills el_f and cF */
RG_DEV_NOINLINE float rg_solver_update(const RgCtx c, int nel, int ncon) {
  RG_LANE_DECL
  const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int* el_i = (const int*)(s + L.el_i);
  LANEVAR(float, part); LANEVAR(int, sigp); LANEVAR(int, conep);
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  const int elliptic = m.opt_cone[0] == 1;
  RG_STAT(memset(rg_stat_act, 0, sizeof rg_stat_act);)
  RG_PHASE_BEGIN
  float cost = 0.0f;
  int cone = 0;
  unsigned sig = 0u;   /* which rows are in their quadratic zone: decides whether H must be rebuilt */
  RG_NOUNROLL for (int e = lane; e < nel; e += 32) {
    const float jar = s[L.el_jar + e], D = s[L.el_D + e];
    float f;
    if ((el_i[e] & 3) == RG_EL_FLOSS) {
      const float fl = m.dof_frictionloss[el_i[e] >> 3], rf = fl / D;
      if (jar <= -rf) { f = fl; cost += -0.5f * rf * fl - fl * jar; }
      else if (jar >= rf) { f = -fl; cost += -0.5f * rf * fl + fl * jar; }
      else { f = -D * jar; cost += 0.5f * D * jar * jar; sig += rg_mix((unsigned)(e + 1)); RG_STAT(rg_stat_act[e] = 1;) }
    } else if (jar < 0.0f) { f = -D * jar; cost += 0.5f * D * jar * jar; sig += rg_mix((unsigned)(e + 1)); RG_STAT(rg_stat_act[e] = 1;) }
    else f = 0.0f;
    s[L.el_f + e] = f;
  }
  RG_NOUNROLL for (int k = lane; k < ncon; k += 32) {
    const float* r = s + L.con + RG_CON_STRIDE * k;
    const float* prm = s + L.cprm + RG_CPRM * k;
    const float* u = s + L.cu + 6 * k;
    const int dim = (int)prm[1];
    const float D = prm[0];
    float F[6] = {0, 0, 0, 0, 0, 0};
    if (dim == 1) { if (u[0] < 0.0f) { F[0] = -D * u[0]; cost += 0.5f * D * u[0] * u[0]; sig += rg_mix((unsigned)(1000 + 16 * k)); } }
    else if (elliptic) {
      float cc;
      const int zone = rg_cone_eval(r, rg_cone_mu(m, r), dim, D, u, &cc, F);
      cost += cc;
      if (zone) sig += rg_mix((unsigned)(1000 + 16 * k + zone));
      if (zone == 2) cone = 1;
    }
    else for (int a = 1; a < dim; a++) {
      const float mu = rg_contact_mu(r, a);
      const float jp = u[0] + mu * u[a], jm = u[0] - mu * u[a];
      if (jp < 0.0f) { const float f = -D * jp; F[0] += f; F[a] += mu * f; cost += 0.5f * D * jp * jp; sig += rg_mix((unsigned)(1000 + 16 * k + 2 * a)); RG_STAT(rg_stat_act[512 + 16 * k + 2 * a] = 1;) }
      if (jm < 0.0f) { const float f = -D * jm; F[0] += f; F[a] -= mu * f; cost += 0.5f * D * jm * jm; sig += rg_mix((unsigned)(1000 + 16 * k + 2 * a + 1)); RG_STAT(rg_stat_act[512 + 16 * k + 2 * a + 1] = 1;) }
    }
    float* cf = s + L.cF + 6 * k;
    for (int a = 0; a < 6; a++) cf[a] = F[a];
  }
  LV(part) = cost;
  LV(sigp) = (int)(sig & 0x00ffffffu);
  LV(conep) = cone;
  RG_PHASE_END
  RG_SI(c, RG_S_SIG) = RG_WARP_ISUM(sigp);
  if (elliptic) RG_SI(c, RG_S_CONE) = RG_WARP_OR(conep);
  return RG_WARP_SUM(part);
}

/* out[d] = sum_rows J^T f for every dof (single-row elements, tendon rows, contacts) */
RG_DEV_NOINLINE void rg_JT_force_phase(const RgCtx c, int out, int nel, int tl0, int ncon) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int* el_i = (const int*)(s + L.el_i);
  const int* eldof = (const int*)(s + L.eldof);
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int d = lane; d < m.nv; d += 32) {
    float acc = 0.0f;
    const int ew = eldof[d], e0 = ew & 0x3ff, e1 = (ew >> 10) & 0x3ff, e2 = (ew >> 20) & 0x3ff;   /* 0x3ff = none */
    if (e0 != 0x3ff) acc += s[L.el_f + e0];
    if (e1 != 0x3ff) acc += s[L.el_f + e1];
    if (e2 != 0x3ff) acc -= s[L.el_f + e2];
    for (int e = tl0; e < nel; e++) {
      const int code = el_i[e];
      const float f = s[L.el_f + e];
      if (f != 0.0f) acc += ((code >> 2) & 1 ? -f : f) * rg_tendon_J(c, code >> 3, d);
    }
    for (int k = 0; k < ncon; k++) {
      const float* r = s + L.con + RG_CON_STRIDE * k;
      const int dim = (int)s[L.cprm + RG_CPRM * k + 1];
      float col[6];
      if (dim == 0 || !rg_contact_col(c, r, d, dim, col)) continue;
      const float* F = s + L.cF + 6 * k;
      RG_NOUNROLL for (int a = 0; a < dim; a++) acc += col[a] * F[a];
    }
    s[out + d] = acc;
  }
  RG_PHASE_END
}

/* el_x[e] = J_e x, cx[k] = Jc_k x  (x = vector at offset xoff); if init, adds -aref / staged velocity terms */
RG_DEV_NOINLINE void rg_J_mul_phase(const RgCtx c, int xoff, int el_out, int c_out, int nel, int ncon, int init) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int* el_i = (const int*)(s + L.el_i);
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int e = lane; e < nel; e += 32) {
    float v = rg_el_Jx(c, el_i[e], xoff);
    if (init) v += s[el_out + e];   /* make_constraints left -aref there */
    s[el_out + e] = v;
  }
  RG_NOUNROLL for (int k = lane; k < ncon; k += 32) {
    const float* r = s + L.con + RG_CON_STRIDE * k;
    const int dim = (int)s[L.cprm + RG_CPRM * k + 1];
    float v[6] = {0, 0, 0, 0, 0, 0};
    const unsigned char* list = (const unsigned char*)(s + L.cdof + ((L.tile + 3) >> 2) * k);
    const int nd = (int)s[L.cprm + RG_CPRM * k + 4];
    const unsigned sgn = (unsigned)rg_f2i(s[L.cprm + RG_CPRM * k + 5]);
    RG_NOUNROLL for (int i = 0; i < nd; i++) {
      float col[6];
      const int d = list[i];
      rg_contact_col_list(c, r, d, (sgn >> i) & 1u ? 1.0f : -1.0f, dim, col);
      const float xd = s[xoff + d];
      RG_NOUNROLL for (int a = 0; a < dim; a++) v[a] += col[a] * xd;
    }
    if (init) for (int a = 0; a < 6; a++) v[a] += s[L.cF + 6 * k + a];
    for (int a = 0; a < 6; a++) s[c_out + 6 * k + a] = v[a];
  }
  RG_PHASE_END
}

RG_DEV_NOINLINE void rg_solve(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int nv = m.nv;
  const int nel = RG_SI(c, RG_S_NEL), tl0 = RG_SI(c, RG_S_TL0), ncon = RG_SI(c, RG_S_NCON);
  const int* el_i = (const int*)(s + L.el_i);
  const int* eldof = (const int*)(s + L.eldof);
  int* env = (int*)(s + L.env);
#if defined(RG_EMU) && defined(RG_DEBUG_NEWTON)
  printf("solve: nel %d tl0 %d ncon %d:", nel, tl0, ncon);
  for (int e = 0; e < nel; e++) printf(" %d/%d/%d", el_i[e] & 3, (el_i[e] >> 2) & 1, el_i[e] >> 3);
  printf("\n");
#endif
  RG_STAT(rg_stat_x[0]++; rg_stat_x[8] += nel;)
  RG_PROFS_BEGIN
  const int ns = m.ns;
  const int* sidx = m.dof_sidx + 0;
  /* trees no constraint can touch: qacc = M^-1 qfrc_smooth, exactly, by the tree-sparse factorisation (the Hessian region
     is still free); they take no part in the iteration below (their search direction stays zero) */
  if (ns < nv) {
    rg_sparse_factor(c, L.H, L.tmp, nullptr, 0.0f, m.dof_xlvl + 0);
    RG_PHASE_BEGIN
    RG_NOUNROLL for (int d = lane; d < nv; d += 32) s[L.search + d] = s[L.smooth + d];
    RG_PHASE_END
    rg_sparse_solve(c, L.H, L.tmp, L.search, m.dof_xlvl + 0);
  }
  /* start from the previous solution (warm start) */
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int d = lane; d < nv; d += 32) {
    if (sidx[d] < 0) { s[L.qacc + d] = s[L.search + d]; s[L.search + d] = 0.0f; }
    else s[L.qacc + d] = (m.opt_disableflags[0] & RG_DSBL_WARMSTART) ? 0.0f : s[L.warm + d];
  }
  RG_PHASE_END
  rg_matvec_phase(c, L.Ma, L.qacc);
  rg_J_mul_phase(c, L.qacc, L.el_jar, L.cu, nel, ncon, 1);
  float cost_con = rg_solver_update(c, nel, ncon);
  const float scale = 1.0f / (m.opt_meaninertia[0] * (float)(nv > 1 ? nv : 1));
  const float tol = fmaxf(m.opt_tolerance[0], RG_NEWTON_TOL); /* fp32: below ~1e-6 the cost differences are rounding noise */
  int iter = 0, have_factor = 0, factor_sig = 0;
  float cost;
  {
    LANEVAR(float, gp);
    RG_PHASE_BEGIN
    float a = 0.0f;
    RG_NOUNROLL for (int d = lane; d < nv; d += 32) a += s[L.qacc + d] * (0.5f * s[L.Ma + d] - s[L.smooth + d]);
    LV(gp) = a;
    RG_PHASE_END
    cost = RG_WARP_SUM(gp) + cost_con;
  }
  RG_PROFS(c, 9)
  int active = 1;
  for (;;) {
  const int go = active && iter < m.opt_iterations[0];
  if (!go) break;   /* (a CTA-wide barrier per Newton iteration was measured: no gain over the per-stage barriers) */
  int done = 1;
  do {
    /* gradient */
    rg_JT_force_phase(c, L.qfc, nel, tl0, ncon);
    LANEVAR(float, gn);
    RG_PHASE_BEGIN
    float a = 0.0f;
    RG_NOUNROLL for (int d = lane; d < nv; d += 32) {
      if (sidx[d] < 0) continue;
      const float g = s[L.Ma + d] - s[L.smooth + d] - s[L.qfc + d];
      s[L.search + d] = -g;
      a += g * g;
    }
    LV(gn) = a;
    RG_PHASE_END
    const float gnorm = sqrtf(RG_WARP_SUM(gn));
    RG_PROFS(c, 10)
    if (scale * gnorm < tol) { RG_STAT(rg_stat_x[10]++;) break; }
    /* Hessian H = M + J' diag(D active) J: rebuilt and refactored only when the active set changed */
#ifdef RG_NO_REUSE
    const int refactor = 1;
#else
    const int refactor = !(have_factor && RG_SI(c, RG_S_SIG) == factor_sig) || RG_SI(c, RG_S_CONE);
#endif
    RG_STAT(rg_stat_x[1]++;)
    if (refactor) {
    RG_STAT(rg_stat_x[2]++; rg_stat_flips(have_factor);)
    rg_H_from_M(c);
    RG_PHASE_BEGIN
    RG_NOUNROLL for (int d = lane; d < nv; d += 32) {
      float add = 0.0f;
      const int ew = eldof[d], e0 = ew & 0x3ff;
      if (e0 != 0x3ff) { const float rf = m.dof_frictionloss[d] / s[L.el_D + e0]; if (fabsf(s[L.el_jar + e0]) < rf) add += s[L.el_D + e0]; }
      for (int q = 1; q < 3; q++) { const int e = (ew >> (10 * q)) & 0x3ff; if (e != 0x3ff && s[L.el_jar + e] < 0.0f) add += s[L.el_D + e]; }
      if (add != 0.0f) s[L.H + RG_TRI(sidx[d], sidx[d])] += add;
    }
    RG_PHASE_END
    for (int e = tl0; e < nel; e++) {
      if (!(s[L.el_jar + e] < 0.0f)) continue;
      const int t = el_i[e] >> 3;
      const float D = s[L.el_D + e];
      const int tn = ((const int*)(s + L.tJn))[t];
      const unsigned char* tji = (const unsigned char*)(s + L.tJi) + RG_TJ * t;
      RG_PHASE_BEGIN
      RG_NOUNROLL for (int p = lane; p < tn * tn; p += 32) {
        const int a = p / tn, b = p - a * tn;
        if (tji[a] >= tji[b]) s[L.H + RG_HS(sidx[tji[a]], sidx[tji[b]])] += D * s[L.tJv + RG_TJ * t + a] * s[L.tJv + RG_TJ * t + b];
      }
      RG_PHASE_END
    }
    for (int k = 0; k < ncon; k++) {
      const float* r = s + L.con + RG_CON_STRIDE * k;
      const float* prm = s + L.cprm + RG_CPRM * k;
      const int dim = (int)prm[1];
      if (dim == 0) continue;
      /* W = sum over active pyramid rows of D c c^T, c = e0 +- mu e_a */
      const float* u = s + L.cu + 6 * k;
      const float D = prm[0];
      float W00 = 0.0f, W0a[6] = {0, 0, 0, 0, 0, 0}, Waa[6] = {0, 0, 0, 0, 0, 0};
      int anyact = 0, zone = 0;
      if (dim == 1) { if (u[0] < 0.0f) { W00 = D; anyact = 1; } }
      else if (m.opt_cone[0] == 1) {
        /* elliptic cone: nothing in the top zone, diag(D_k) in the bottom zone (fits the arrow form below), a dense block in
           the middle zone (handled separately) */
        float cc, F[6];
        const float mu = rg_cone_mu(m, r);
        zone = rg_cone_eval(r, mu, dim, D, u, &cc, F);
        if (zone == 1) {
          W00 = D; anyact = 1;
          for (int a = 1; a < 6; a++) if (a < dim) { const float fa = rg_contact_mu(r, a); Waa[a] = D * fa * fa / (mu * mu); }
        } else if (zone == 2) anyact = 1;
      }
      else for (int a = 1; a < dim; a++) {
        const float mu = rg_contact_mu(r, a);
        if (u[0] + mu * u[a] < 0.0f) { W00 += D; W0a[a] += D * mu; Waa[a] += D * mu * mu; anyact = 1; }
        if (u[0] - mu * u[a] < 0.0f) { W00 += D; W0a[a] -= D * mu; Waa[a] += D * mu * mu; anyact = 1; }
      }
      if (!anyact) continue;
      /* tile: columns of Jc for the dofs this contact touches, and W Jc */
      int* tdof = (int*)(s + L.tileDof);
      const unsigned char* list = (const unsigned char*)(s + L.cdof + ((L.tile + 3) >> 2) * k);
      int nd = (int)prm[4];
      const unsigned sgn = (unsigned)rg_f2i(prm[5]);
      RG_PHASE_BEGIN
      if (lane < nd) {
        float col[6] = {0, 0, 0, 0, 0, 0};
        const int d = list[lane];
        rg_contact_col_list(c, r, d, (sgn >> lane) & 1u ? 1.0f : -1.0f, dim, col);
        tdof[lane] = d;
        float* tj = s + L.tileJ + 6 * lane;
        float* tw = s + L.tileWJ + 6 * lane;
        if (zone == 2) {
          float W[6][6];
          rg_cone_hessian(r, rg_cone_mu(m, r), dim, D, u, W);
          for (int a = 0; a < 6; a++) {
            float w = 0.0f;
            for (int b = 0; b < 6; b++) if (a < dim && b < dim) w += W[a][b] * col[b];
            tj[a] = a < dim ? col[a] : 0.0f; tw[a] = w;
          }
        } else {
        float w0 = W00 * col[0];
        RG_NOUNROLL for (int a = 1; a < dim; a++) w0 += W0a[a] * col[a];
        tj[0] = col[0]; tw[0] = w0;
        for (int a = 1; a < 6; a++) { tj[a] = a < dim ? col[a] : 0.0f; tw[a] = a < dim ? W0a[a] * col[0] + Waa[a] * col[a] : 0.0f; }
        }
      }
      RG_PHASE_END
      RG_PHASE_BEGIN
      RG_NOUNROLL for (int p = lane; p < nd * nd; p += 32) {
        const int i = p / nd, j = p - i * nd;
        if (tdof[i] < tdof[j]) continue;
        const float* tj = s + L.tileJ + 6 * i;
        const float* tw = s + L.tileWJ + 6 * j;
        float acc = 0.0f;
        RG_NOUNROLL for (int a = 0; a < dim; a++) acc += tj[a] * tw[a];
        if (tdof[i] >= tdof[j]) s[L.H + RG_HS(sidx[tdof[i]], sidx[tdof[j]])] += acc;
      }
      RG_PHASE_END
    }
    /* envelope, factor, Newton direction */
    RG_PHASE_BEGIN
    RG_NOUNROLL for (int i = lane; i < ns; i += 32) {
      int e = 0;
      while (e < i && s[L.H + RG_TRI(i, e)] == 0.0f) e++;
      env[i] = e;
    }
    RG_PHASE_END
    RG_PROFS(c, 11)
    RG_PHASE_BEGIN   /* right-hand side -> row nv of H, in the solver's reversed dof order */
    RG_NOUNROLL for (int d = lane; d < nv; d += 32) if (sidx[d] >= 0) s[L.H + RG_TRI(ns, sidx[d])] = s[L.search + d];
    RG_PHASE_END
    rg_cholesky(c, L.H, env);
    RG_PROFS(c, 12)
    have_factor = 1; factor_sig = RG_SI(c, RG_S_SIG);
    } else {
    RG_PHASE_BEGIN
    RG_NOUNROLL for (int d = lane; d < nv; d += 32) if (sidx[d] >= 0) s[L.H + RG_TRI(ns, sidx[d])] = s[L.search + d];
    RG_PHASE_END
    rg_chol_forward(c, L.H, env);
    }
#if defined(RG_EMU) && defined(RG_DEBUG_NEWTON)
    { double cs = 0; for (int i = 0; i < (ns * (ns + 1)) / 2; i++) cs += s[L.H + i] * (1 + (i % 7)); int es = 0; for (int i = 0; i < ns; i++) es += env[i] * (i + 1); printf("    L checksum %.9g env %d\n", cs, es); }
#endif
    rg_chol_back(c, L.H, env, L.search);
    RG_PROFS(c, 13)
#if defined(RG_EMU) && defined(RG_DEBUG_NEWTON)
    { /* finite-difference check of the Newton direction: g(q + eps*s) should be ~ (1-eps) g(q) when the active set holds */
      static float q0[256], g0v[256], ma0[256], jar0[256], cu0[6 * 512], f0[256], cf0[6 * 512], qfc0[256];
      for (int d = 0; d < nv; d++) { q0[d] = s[L.qacc + d]; ma0[d] = s[L.Ma + d]; qfc0[d] = s[L.qfc + d]; g0v[d] = s[L.Ma + d] - s[L.smooth + d] - s[L.qfc + d]; }
      for (int e = 0; e < nel; e++) { jar0[e] = s[L.el_jar + e]; f0[e] = s[L.el_f + e]; }
      for (int i = 0; i < 6 * ncon; i++) { cu0[i] = s[L.cu + i]; cf0[i] = s[L.cF + i]; }
      const float eps = 1e-3f;
      for (int d = 0; d < nv; d++) s[L.qacc + d] = q0[d] + eps * s[L.search + d];
      rg_matvec_phase(c, L.Ma, L.qacc);
      for (int e = 0; e < nel; e++) s[L.el_jar + e] = jar0[e] - rg_el_Jx(c, el_i[e], L.qacc) + rg_el_Jx(c, el_i[e], L.qacc);   /* placeholder */
      /* recompute jar from scratch: jar = J q - aref ; we know jar0 = J q0 - aref */
      rg_J_mul_phase(c, L.search, L.el_jv, L.cw, nel, ncon, 0);
      for (int e = 0; e < nel; e++) s[L.el_jar + e] = jar0[e] + eps * s[L.el_jv + e];
      for (int i = 0; i < 6 * ncon; i++) s[L.cu + i] = cu0[i] + eps * s[L.cw + i];
      const int sig_before = RG_SI(c, RG_S_SIG);
      rg_solver_update(c, nel, ncon);
      rg_JT_force_phase(c, L.qfc, nel, tl0, ncon);
      float num = 0, den = 0, dotg = 0;
      for (int d = 0; d < nv; d++) { const float g1 = s[L.Ma + d] - s[L.smooth + d] - s[L.qfc + d]; const float pred = (1.0f - eps) * g0v[d]; num += (g1 - pred) * (g1 - pred); den += g0v[d] * g0v[d]; dotg += g0v[d] * s[L.search + d]; }
      printf("    FD check: |g(q+eps s) - (1-eps) g|/|g| = %.3g  (sig %d -> %d)  g.s %.4g\n", sqrtf(num / den), sig_before, RG_SI(c, RG_S_SIG), dotg);
      for (int d = 0; d < nv; d++) { s[L.qacc + d] = q0[d]; s[L.Ma + d] = ma0[d]; s[L.qfc + d] = qfc0[d]; }
      for (int e = 0; e < nel; e++) { s[L.el_jar + e] = jar0[e]; s[L.el_f + e] = f0[e]; }
      for (int i = 0; i < 6 * ncon; i++) { s[L.cu + i] = cu0[i]; s[L.cF + i] = cf0[i]; }
      RG_SI(c, RG_S_SIG) = sig_before;
    }
#endif
    /* line search along `search` */
    rg_matvec_phase(c, L.Mv, L.search);
    rg_J_mul_phase(c, L.search, L.el_jv, L.cw, nel, ncon, 0);
    float q1, q2;
    {
      LANEVAR(float, p1); LANEVAR(float, p2);
      RG_PHASE_BEGIN
      float a = 0.0f, b = 0.0f;
      RG_NOUNROLL for (int d = lane; d < nv; d += 32) { a += s[L.search + d] * (s[L.Ma + d] - s[L.smooth + d]); b += s[L.search + d] * s[L.Mv + d]; }
      LV(p1) = a; LV(p2) = 0.5f * b;
      RG_PHASE_END
      q1 = RG_WARP_SUM(p1); q2 = RG_WARP_SUM(p2);
    }
    float alpha = 0.0f, lo = 0.0f, hi = -1.0f, g0 = 0.0f;
    for (int ls = 0; ls < 12; ls++) {
      LANEVAR(float, pg); LANEVAR(float, ph);
      RG_STAT(rg_stat_x[3]++;)
      RG_PHASE_BEGIN
      float g = 0.0f, h = 0.0f;
      RG_NOUNROLL for (int e = lane; e < nel; e += 32) {
        const float jv = s[L.el_jv + e], x = s[L.el_jar + e] + alpha * jv, D = s[L.el_D + e];
        if ((el_i[e] & 3) == RG_EL_FLOSS) {
          const float fl = m.dof_frictionloss[el_i[e] >> 3], rf = fl / D;
          if (x <= -rf) g -= fl * jv;
          else if (x >= rf) g += fl * jv;
          else { g += D * x * jv; h += D * jv * jv; }
        } else if (x < 0.0f) { g += D * x * jv; h += D * jv * jv; }
      }
      RG_NOUNROLL for (int k = lane; k < ncon; k += 32) {
        const float* r = s + L.con + RG_CON_STRIDE * k;
        const float* prm = s + L.cprm + RG_CPRM * k;
        const float* u = s + L.cu + 6 * k;
        const float* w = s + L.cw + 6 * k;
        const int dim = (int)prm[1];
        const float D = prm[0];
        if (dim == 1) { const float x = u[0] + alpha * w[0]; if (x < 0.0f) { g += D * x * w[0]; h += D * w[0] * w[0]; } }
        else if (m.opt_cone[0] == 1) rg_cone_ray(r, rg_cone_mu(m, r), dim, D, u, w, alpha, &g, &h);
        else for (int a = 1; a < dim; a++) {
          const float mu = rg_contact_mu(r, a);
          const float xp = u[0] + alpha * w[0] + mu * (u[a] + alpha * w[a]), vp = w[0] + mu * w[a];
          const float xm = u[0] + alpha * w[0] - mu * (u[a] + alpha * w[a]), vm = w[0] - mu * w[a];
          if (xp < 0.0f) { g += D * xp * vp; h += D * vp * vp; }
          if (xm < 0.0f) { g += D * xm * vm; h += D * vm * vm; }
        }
      }
      LV(pg) = g; LV(ph) = h;
      RG_PHASE_END
      const float g = RG_WARP_SUM(pg) + q1 + 2.0f * alpha * q2;
      const float h = RG_WARP_SUM(ph) + 2.0f * q2;
      if (ls == 0) { g0 = g; if (!(g0 < 0.0f)) break; }
      else {
        if (fabsf(g) <= RG_LS_TOL * fabsf(g0)) break;
        if (g < 0.0f) lo = alpha; else hi = alpha;
        if (hi >= 0.0f && hi - lo <= 1e-7f * hi) break;
      }
      float a2 = alpha - g / fmaxf(h, 1e-30f);
      if (hi >= 0.0f && (a2 <= lo || a2 >= hi)) a2 = 0.5f * (lo + hi);
      alpha = a2;
    }
    RG_PROFS(c, 14)
    if (!(alpha > 0.0f)) { RG_STAT(rg_stat_x[11]++;) break; }
    /* take the step */
    RG_PHASE_BEGIN
    RG_NOUNROLL for (int d = lane; d < nv; d += 32) { s[L.qacc + d] += alpha * s[L.search + d]; s[L.Ma + d] += alpha * s[L.Mv + d]; }
    RG_NOUNROLL for (int e = lane; e < nel; e += 32) s[L.el_jar + e] += alpha * s[L.el_jv + e];
    RG_NOUNROLL for (int i = lane; i < 6 * ncon; i += 32) s[L.cu + i] += alpha * s[L.cw + i];
    RG_PHASE_END
    cost_con = rg_solver_update(c, nel, ncon);
    float newcost;
    {
      LANEVAR(float, gp);
      RG_PHASE_BEGIN
      float a = 0.0f;
      RG_NOUNROLL for (int d = lane; d < nv; d += 32) a += s[L.qacc + d] * (0.5f * s[L.Ma + d] - s[L.smooth + d]);
      LV(gp) = a;
      RG_PHASE_END
      newcost = RG_WARP_SUM(gp) + cost_con;
    }
    const float improvement = scale * (cost - newcost);
    RG_PROFS(c, 15)
    RG_STAT(if (getenv("RG_TRACE") && rg_stat_x[0] % 97 == 0) printf("  fw %lld it %d g %.3g alpha %.4g impr %.3g refac %d nel %d ncon %d\n", rg_stat_x[0], iter, scale * gnorm, alpha, improvement, refactor, nel, ncon);)
#if defined(RG_EMU) && defined(RG_DEBUG_NEWTON)
    printf("  it %d cost %.9g new %.9g impr %.3g alpha %.6g gnorm %.3g refactor %d sig %d\n", iter, cost, newcost, improvement, alpha, gnorm, refactor, RG_SI(c, RG_S_SIG));
#endif
    cost = newcost;
    /* fp32: cost differences below ~2 ulp of the cost itself are rounding noise, not progress */
    if (improvement < tol + 2.4e-7f * fabsf(cost) * scale) { RG_STAT(rg_stat_x[12]++;) iter++; break; }
    done = 0;
  } while (0);
  if (done) active = 0; else iter++;
  }
  rg_JT_force_phase(c, L.qfc, nel, tl0, ncon);
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int d = lane; d < nv; d += 32) s[L.warm + d] = s[L.qacc + d];
  RG_STAT(if (lane == 0) rg_stat_iterhist[iter < 15 ? iter : 15]++;)
  if (lane == 0) { RG_SI(c, RG_S_NITER) = iter; RG_SI(c, RG_S_WORK) += RG_COST_ITER * iter; }
  RG_PHASE_END
}

/* ---------------------------------------------------------------- S15 semi-implicit Euler */
RG_DEV_NOINLINE void rg_euler(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  const int nv = m.nv;
  const float h = c.timestep;
  /* (M + h B) qacc_damped = qfrc_smooth + qfrc_constraint; the dense Hessian region is free here and holds the factor */
  rg_sparse_factor(c, L.H, L.tmp, m.dof_damping + 0, h, m.dof_lvl + 0);
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int d = lane; d < nv; d += 32) s[L.search + d] = s[L.smooth + d] + s[L.qfc + d];
  RG_PHASE_END
  rg_sparse_solve(c, L.H, L.tmp, L.search, m.dof_lvl + 0);
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int d = lane; d < nv; d += 32) s[L.qvel + d] += h * s[L.search + d];
  RG_PHASE_END
  RG_PHASE_BEGIN
  RG_NOUNROLL for (int j = lane; j < m.njnt; j += 32) {
    const int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j], type = m.jnt_type[j];
    if (type == RG_JNT_SLIDE || type == RG_JNT_HINGE) s[L.qpos + qa] += h * s[L.qvel + da];
    else {
      int qq = qa, dd = da;
      if (type == RG_JNT_FREE) { for (int a = 0; a < 3; a++) s[L.qpos + qa + a] += h * s[L.qvel + da + a]; qq += 3; dd += 3; }
      float w[3] = {s[L.qvel + dd], s[L.qvel + dd + 1], s[L.qvel + dd + 2]};
      const float ang = sqrtf(rg_dot3(w, w)) * h;
      if (ang > 1e-15f) {
        rg_normalize3(w);
        float sn, cs;
        RG_SINCOS(0.5f * ang, &sn, &cs);
        const float dq[4] = {cs, w[0] * sn, w[1] * sn, w[2] * sn};
        float q[4] = {s[L.qpos + qq], s[L.qpos + qq + 1], s[L.qpos + qq + 2], s[L.qpos + qq + 3]}, rq[4];
        rg_quat_mul(rq, q, dq);
        rg_quat_norm(rq);
        for (int a = 0; a < 4; a++) s[L.qpos + qq + a] = rq[a];
      }
    }
  }
  RG_PHASE_END
}
