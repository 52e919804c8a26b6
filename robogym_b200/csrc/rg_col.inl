/* rg_col.inl -- S7 collision of the fused step: streaming broad phase over the statically
 * filtered pair list + lane-per-pair narrow phase (plane-box, plane-sphere, plane-convex,
 * Minkowski Portal Refinement for every convex-convex pair, each geom inflated by margin/2).
 * Replaces mj_collision inside the mj_step that robogym runs through sim.step()
 * (robogym/mujoco/simulation_interface.py:184).  Hull support mapping is a hill climb on the
 * hull's edge graph (model arrays mesh_adjadr/mesh_adj) warm-started from the previous answer.
 * Contact order is the pair-list order: deterministic and independent of the batch slot.
 */
#pragma once
#include "rg_dyn.inl"

#if defined(RG_EMU) && defined(RG_STATS)
#include <stdio.h>
#include <stdlib.h>
static long long rg_stat_hist[2][64] = {{0}}; static long long rg_stat_support = 0, rg_stat_climb = 0, rg_stat_mpr = 0, rg_stat_mpr_hit = 0, rg_stat_narrow = 0, rg_stat_maxsup = 0, rg_stat_cur = 0;
/* solver / pipeline counters: forwards, newton iterations, refactorisations, line-search evaluations, triangular solves, broad-phase survivors, obb survivors, contacts, rows, mpr iterations */
static long long rg_stat_x[16] = {0};
static long long rg_stat_iterhist[16] = {0};
static long long rg_stat_fliphist[16] = {0};
static unsigned char rg_stat_act[2048], rg_stat_prev[2048];
static void rg_stat_flips(int have_factor) {
  if (have_factor) {
    int fl = 0, flc = 0;
    for (int q = 0; q < 512; q++) fl += rg_stat_act[q] != rg_stat_prev[q];
    for (int q = 512; q < 2048; q++) flc += rg_stat_act[q] != rg_stat_prev[q];
    rg_stat_fliphist[fl + flc < 15 ? fl + flc : 15]++;
    rg_stat_x[13] += fl; rg_stat_x[14] += flc; rg_stat_x[15]++;
  }
  memcpy(rg_stat_prev, rg_stat_act, sizeof rg_stat_act);
}
#define RG_STAT(x) x
#else
#define RG_STAT(x)
#endif
struct RgGeomView {
  float pos[3], mat[9], size[3];
  int type, vadr, vnum;
  float halfmargin;
};

RG_DEV void rg_geom_view(const RgCtx c, int g, float margin, RgGeomView& v) {
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  const int b = m.geom_bodyid[g];
  float q[4];
  rg_quat_mul(q, RG_SCRATCH(c) + RG_CL(c).xquat + 4 * b, m.geom_quat + 4 * g);
  rg_quat_norm(q);
  rg_quat2mat(v.mat, q);
  rg_copy3(v.pos, RG_SCRATCH(c) + RG_CL(c).gxpos + 3 * g);
  rg_copy3(v.size, m.geom_size + 3 * g);
  v.type = m.geom_type[g];
  v.halfmargin = 0.5f * margin;
  v.vadr = 0; v.vnum = 0;
  if (v.type == RG_GEOM_MESH) { const int mid = m.geom_dataid[g]; v.vadr = m.mesh_vertadr[mid]; v.vnum = m.mesh_vertnum[mid]; }
}

/* support point of a primitive in its own frame */
RG_DEV void rg_support_prim(const RgGeomView& v, const float* dl, float* loc) {
  loc[0] = loc[1] = loc[2] = 0.0f;
  switch (v.type) {
    case RG_GEOM_SPHERE: rg_scl3(loc, dl, v.size[0]); break;
    case RG_GEOM_BOX:
      loc[0] = dl[0] >= 0 ? v.size[0] : -v.size[0];
      loc[1] = dl[1] >= 0 ? v.size[1] : -v.size[1];
      loc[2] = dl[2] >= 0 ? v.size[2] : -v.size[2];
      break;
    case RG_GEOM_CAPSULE:
      rg_scl3(loc, dl, v.size[0]);
      loc[2] += dl[2] >= 0 ? v.size[1] : -v.size[1];
      break;
    case RG_GEOM_CYLINDER: {
      const float n = sqrtf(dl[0] * dl[0] + dl[1] * dl[1]);
      if (n > 1e-20f) { loc[0] = dl[0] / n * v.size[0]; loc[1] = dl[1] / n * v.size[0]; }
      loc[2] = dl[2] >= 0 ? v.size[1] : -v.size[1];
    } break;
    case RG_GEOM_ELLIPSOID: {
      float t[3] = {dl[0] * v.size[0], dl[1] * v.size[1], dl[2] * v.size[2]};
      const float n = sqrtf(rg_dot3(t, t));
      if (n > 1e-20f) { loc[0] = t[0] * v.size[0] / n; loc[1] = t[1] * v.size[1] / n; loc[2] = t[2] * v.size[2] / n; }
    } break;
    default: break;
  }
}

/* ---- hull support mapping: exhaustive scan of the hull's vertices (16-byte loads, no dependent chain).  In the
 * convex-convex narrow phase the scan of one hull is shared by the RG_GRP lanes that work on the pair: lane `gl` looks
 * at vertices gl, gl + RG_GRP, ... and the group's arg-max picks the winner (ties: the lower vertex id, which is what a
 * plain first-maximum loop returns). */
RG_DEV void rg_hull_scan(const RG_MODEL_T& m, const RgGeomView& v, const float* dl, int first, int stride, float& best, int& idx) {
  best = -3.0e38f; idx = 0x7fffffff;
  RG_STAT(rg_stat_climb += (v.vnum - first + stride - 1) / stride;)
  RG_UNROLL4 for (int k = first; k < v.vnum; k += stride) {
    float p[4];
    RG_LDG4(m.mesh_vert4, v.vadr + k, p);
    const float d = p[0] * dl[0] + p[1] * dl[1] + p[2] * dl[2];
    if (d > best) { best = d; idx = k; }
  }
}
RG_DEV void rg_support_world(const RgGeomView& v, const float* loc, const float* dir, float* res) {
  rg_mulmat3(res, v.mat, loc);
  res[0] += v.pos[0] + dir[0] * v.halfmargin;
  res[1] += v.pos[1] + dir[1] * v.halfmargin;
  res[2] += v.pos[2] + dir[2] * v.halfmargin;
}
/* one lane, whole hull (plane-convex pairs only: rare and cheap enough) */
RG_DEV void rg_support(const RG_MODEL_T& m, const RgGeomView& v, const float* dir, float* res) {
  float dl[3], loc[3];
  RG_STAT(rg_stat_support++; rg_stat_cur++;)
  rg_mulmatT3(dl, v.mat, dir);
  if (v.type == RG_GEOM_MESH) {
    float best; int idx;
    rg_hull_scan(m, v, dl, 0, 1, best, idx);
    float p[4];
    RG_LDG4(m.mesh_vert4, v.vadr + idx, p);
    rg_copy3(loc, p);
  } else rg_support_prim(v, dl, loc);
  rg_support_world(v, loc, dir, res);
}

struct RgSup { float v[3], v1[3], v2[3]; };
RG_DEV int rg_mpr_zero(float x) { return fabsf(x) < RG_EPS; }
/* ---- Minkowski Portal Refinement (XenoCollide), written as ONE loop around the support call:
 * the portal lives in named registers (no indexed local arrays) and lanes that are in different
 * stages of the algorithm (portal discovery / refinement / penetration) still execute the expensive
 * support evaluation together -- only the cheap bookkeeping diverges.  The sequence of operations
 * per pair is the textbook one (discover portal -> refine until the origin is inside -> push the
 * portal to the surface), identical to the oracle's three-loop formulation. */
RG_DEV void rg_portal_dir3(const RgSup& a, const RgSup& b, const RgSup& cc, float* dir) {
  float u[3], w[3];
  rg_sub3(u, b.v, a.v); rg_sub3(w, cc.v, a.v);
  rg_cross(dir, u, w);
  rg_normalize3(dir);
}
RG_DEV void rg_expand_portal3(RgSup& p1, RgSup& p2, RgSup& p3, const float* v0, const RgSup& v4) {
  float c4[3];
  rg_cross(c4, v4.v, v0);
  if (rg_dot3(p1.v, c4) > 0) { if (rg_dot3(p2.v, c4) > 0) p1 = v4; else p3 = v4; }
  else { if (rg_dot3(p3.v, c4) > 0) p2 = v4; else p1 = v4; }
}
RG_DEV int rg_reach_tol3(const RgSup& p1, const RgSup& p2, const RgSup& p3, const RgSup& v4, const float* dir, float tol) {
  const float dv4 = rg_dot3(v4.v, dir);
  const float mn = fminf(fminf(dv4 - rg_dot3(p1.v, dir), dv4 - rg_dot3(p2.v, dir)), dv4 - rg_dot3(p3.v, dir));
  return mn <= tol || rg_mpr_zero(mn - tol);
}
/* closest point of triangle abc to the origin */
RG_DEV float rg_tri_closest(const float* a, const float* b, const float* cc, float* w) {
  float ab[3], ac[3];
  rg_sub3(ab, b, a); rg_sub3(ac, cc, a);
  const float d1 = -rg_dot3(ab, a), d2 = -rg_dot3(ac, a);
  if (d1 <= 0 && d2 <= 0) { rg_copy3(w, a); return rg_dot3(w, w); }
  const float d3 = -rg_dot3(ab, b), d4 = -rg_dot3(ac, b);
  if (d3 >= 0 && d4 <= d3) { rg_copy3(w, b); return rg_dot3(w, w); }
  const float vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { rg_copy3(w, a); rg_addscl3(w, ab, d1 / (d1 - d3)); return rg_dot3(w, w); }
  const float d5 = -rg_dot3(ab, cc), d6 = -rg_dot3(ac, cc);
  if (d6 >= 0 && d5 <= d6) { rg_copy3(w, cc); return rg_dot3(w, w); }
  const float vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { rg_copy3(w, a); rg_addscl3(w, ac, d2 / (d2 - d6)); return rg_dot3(w, w); }
  const float va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    float bc[3];
    rg_sub3(bc, cc, b);
    rg_copy3(w, b); rg_addscl3(w, bc, (d4 - d3) / ((d4 - d3) + (d5 - d6)));
    return rg_dot3(w, w);
  }
  const float den = 1.0f / (va + vb + vc);
  rg_copy3(w, a); rg_addscl3(w, ab, vb * den); rg_addscl3(w, ac, vc * den);
  return rg_dot3(w, w);
}
RG_DEV void rg_find_pos3(const float* v0, const float* c1, const float* c2, const RgSup& p1, const RgSup& p2, const RgSup& p3, float* pos) {
  float dir[3], b[4], t[3];
  rg_portal_dir3(p1, p2, p3, dir);
  rg_cross(t, p1.v, p2.v); b[0] = rg_dot3(t, p3.v);
  rg_cross(t, p3.v, p2.v); b[1] = rg_dot3(t, v0);
  rg_cross(t, v0, p1.v); b[2] = rg_dot3(t, p3.v);
  rg_cross(t, p2.v, p1.v); b[3] = rg_dot3(t, v0);
  float sum = b[0] + b[1] + b[2] + b[3];
  if (sum <= 1e-30f) {
    b[0] = 0;
    rg_cross(t, p2.v, p3.v); b[1] = rg_dot3(t, dir);
    rg_cross(t, p3.v, p1.v); b[2] = rg_dot3(t, dir);
    rg_cross(t, p1.v, p2.v); b[3] = rg_dot3(t, dir);
    sum = b[1] + b[2] + b[3];
  }
  const float inv = 0.5f / sum;
  for (int i = 0; i < 3; i++)
    pos[i] = inv * (b[0] * (c1[i] + c2[i]) + b[1] * (p1.v1[i] + p1.v2[i]) + b[2] * (p2.v1[i] + p2.v2[i]) + b[3] * (p3.v1[i] + p3.v2[i]));
}

/* ---- MPR as a resumable state machine.  The support evaluation sits OUTSIDE: the narrow phase runs RG_GRP lanes per
 * pair, scans both hulls cooperatively, then every lane of the group advances its (identical) copy of this state with
 * the new support point.  The sequence of operations per pair is the textbook one (discover portal -> refine until the
 * origin is inside -> push the portal to the surface), identical to the oracle's three-loop formulation. */
enum { RG_MPR_V1 = 0, RG_MPR_V2 = 1, RG_MPR_V3 = 2, RG_MPR_REFINE = 3, RG_MPR_PENETR = 4, RG_MPR_PRE = 5 /* trying the cached separating axis */, RG_MPR_DONE = 6, RG_MPR_IDLE = 7, RG_MPR_FINISHED = 8 };
struct RgMpr {
  RgGeomView o1, o2;
  RgSup P1, P2, P3;
  float v0[3], dir[3];
  float depth, odir[3], opos[3];   /* result: depth >= 0 with direction and position when the inflated geoms intersect, -1 otherwise */
  int state, pen_it, it, slot, pair;
  float lastdot;                   /* support . dir of the last iteration (< 0: dir separates the geoms) */
};
RG_DEV void rg_mpr_begin(RgMpr& S) {
  rg_sub3(S.v0, S.o1.pos, S.o2.pos);
  if (rg_mpr_zero(S.v0[0]) && rg_mpr_zero(S.v0[1]) && rg_mpr_zero(S.v0[2])) S.v0[0] = 1e-5f;
  rg_scl3(S.dir, S.v0, -1.0f); rg_normalize3(S.dir);
  S.state = RG_MPR_V1; S.pen_it = 0; S.it = 0;
  S.depth = -1.0f;
  S.odir[0] = S.odir[1] = S.odir[2] = 0.0f; S.opos[0] = S.opos[1] = S.opos[2] = 0.0f;
}
/* consume the support point `sp` of the Minkowski difference in direction S.dir; leaves the next direction in S.dir */
RG_DEV void rg_mpr_advance(RgMpr& S, const RgSup& sp, float tol, int maxiter) {
  float va[3], vb[3];
  RgSup& P1 = S.P1; RgSup& P2 = S.P2; RgSup& P3 = S.P3;
  float* dir = S.dir; const float* v0 = S.v0;
  const float dot = rg_dot3(sp.v, dir);
  S.it++;
  S.lastdot = dot;
  if (S.state == RG_MPR_PRE) {
    if (dot < 0 && !rg_mpr_zero(dot)) { S.state = RG_MPR_DONE; return; }   /* still separated along last time's axis */
    rg_mpr_begin(S);
    return;
  }
  if (S.state == RG_MPR_V1) {
    P1 = sp;
    if (dot < 0 || rg_mpr_zero(dot)) { S.state = RG_MPR_DONE; return; }
    rg_cross(dir, v0, P1.v);
    /* fp32: the parallel test must be scale-free (libccd compares the raw squared norm with DBL_EPSILON) */
    if (rg_dot3(dir, dir) <= 1e-12f * rg_dot3(v0, v0) * rg_dot3(P1.v, P1.v)) {
      const float n = sqrtf(rg_dot3(P1.v, P1.v));
      for (int i = 0; i < 3; i++) S.opos[i] = 0.5f * (P1.v1[i] + P1.v2[i]);
      if (n < RG_EPS) { S.odir[0] = S.odir[1] = S.odir[2] = 0; S.depth = 0.0f; }
      else { rg_scl3(S.odir, P1.v, 1.0f / n); S.depth = n; }
      S.state = RG_MPR_DONE;
      return;
    }
    rg_normalize3(dir);
    S.state = RG_MPR_V2;
  } else if (S.state == RG_MPR_V2) {
    P2 = sp;
    if (dot < 0 || rg_mpr_zero(dot)) { S.state = RG_MPR_DONE; return; }
    rg_sub3(va, P1.v, v0); rg_sub3(vb, P2.v, v0);
    rg_cross(dir, va, vb); rg_normalize3(dir);
    if (rg_dot3(dir, v0) > 0) { const RgSup t = P1; P1 = P2; P2 = t; rg_scl3(dir, dir, -1.0f); }
    S.state = RG_MPR_V3;
  } else if (S.state == RG_MPR_V3) {
    P3 = sp;
    if (dot < 0 || rg_mpr_zero(dot)) { S.state = RG_MPR_DONE; return; }
    int cont = 0;
    rg_cross(va, P1.v, P3.v);
    if (rg_dot3(va, v0) < 0) { P2 = P3; cont = 1; }     /* triple products are ~1e-6: sign only */
    if (!cont) {
      rg_cross(va, P3.v, P2.v);
      if (rg_dot3(va, v0) < 0) { P1 = P3; cont = 1; }
    }
    if (cont) {
      rg_sub3(va, P1.v, v0); rg_sub3(vb, P2.v, v0);
      rg_cross(dir, va, vb); rg_normalize3(dir);
    } else {
      rg_portal_dir3(P1, P2, P3, dir);
      const float d1 = rg_dot3(dir, P1.v);
      S.state = (d1 > 0 || rg_mpr_zero(d1)) ? RG_MPR_PENETR : RG_MPR_REFINE;
    }
  } else if (S.state == RG_MPR_REFINE) {
    if (!(dot > 0 || rg_mpr_zero(dot)) || rg_reach_tol3(P1, P2, P3, sp, dir, tol)) { S.state = RG_MPR_DONE; return; }
    rg_expand_portal3(P1, P2, P3, v0, sp);
    rg_portal_dir3(P1, P2, P3, dir);
    const float d1 = rg_dot3(dir, P1.v);
    if (d1 > 0 || rg_mpr_zero(d1)) S.state = RG_MPR_PENETR;
  } else { /* RG_MPR_PENETR */
    if (rg_reach_tol3(P1, P2, P3, sp, dir, tol) || S.pen_it > maxiter) {
      float w[3];
      const float depth = sqrtf(rg_tri_closest(P1.v, P2.v, P3.v, w));
      if (depth < RG_EPS) { S.odir[0] = S.odir[1] = S.odir[2] = 0; }
      else rg_scl3(S.odir, w, 1.0f / depth);
      rg_find_pos3(v0, S.o1.pos, S.o2.pos, P1, P2, P3, S.opos);
      S.depth = depth;
      S.state = RG_MPR_DONE;
      return;
    }
    rg_expand_portal3(P1, P2, P3, v0, sp);
    rg_portal_dir3(P1, P2, P3, dir);
    S.pen_it++;
  }
  if (S.it >= 200 + maxiter) S.state = RG_MPR_DONE;
}

RG_DEV void rg_make_frame(const float* n, float* t1, float* t2) {
  float tmp[3] = {0, 0, 0};
  if (fabsf(n[1]) < 0.5f) tmp[1] = 1.0f; else tmp[2] = 1.0f;
  const float dd = rg_dot3(n, tmp);
  t1[0] = tmp[0] - dd * n[0]; t1[1] = tmp[1] - dd * n[1]; t1[2] = tmp[2] - dd * n[2];
  rg_normalize3(t1);
  rg_cross(t2, n, t1);
}

/* narrow phase of one pair; writes up to 4 (dist,pos,normal) records to out[7*i..]; returns count */
/* oriented-box overlap (separating-axis test on the geoms' local bounding boxes, box 1 grown by margin):
 * a conservative cull between the bounding-sphere test and MPR; it never removes a pair that could touch */
RG_DEV_NOINLINE int rg_obb_overlap(const RgCtx c, int g1, int g2, float margin) {
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  float q[4], A[9], B[9], ca[3], cb[3], t[3], d[3];
  rg_quat_mul(q, RG_SCRATCH(c) + RG_CL(c).xquat + 4 * m.geom_bodyid[g1], m.geom_quat + 4 * g1); rg_quat_norm(q); rg_quat2mat(A, q);
  rg_quat_mul(q, RG_SCRATCH(c) + RG_CL(c).xquat + 4 * m.geom_bodyid[g2], m.geom_quat + 4 * g2); rg_quat_norm(q); rg_quat2mat(B, q);
  rg_mulmat3(ca, A, m.geom_aabb + 6 * g1); rg_add3(ca, ca, RG_SCRATCH(c) + RG_CL(c).gxpos + 3 * g1);
  rg_mulmat3(cb, B, m.geom_aabb + 6 * g2); rg_add3(cb, cb, RG_SCRATCH(c) + RG_CL(c).gxpos + 3 * g2);
  const float a[3] = {m.geom_aabb[6 * g1 + 3] + margin, m.geom_aabb[6 * g1 + 4] + margin, m.geom_aabb[6 * g1 + 5] + margin};
  const float* b = m.geom_aabb + 6 * g2 + 3;
  rg_sub3(d, cb, ca);
  rg_mulmatT3(t, A, d);
  float R[9], AR[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
      AR[3 * i + j] = fabsf(R[3 * i + j]) + 1e-6f;
    }
  for (int i = 0; i < 3; i++)
    if (fabsf(t[i]) > a[i] + b[0] * AR[3 * i] + b[1] * AR[3 * i + 1] + b[2] * AR[3 * i + 2]) return 0;
  for (int j = 0; j < 3; j++)
    if (fabsf(t[0] * R[j] + t[1] * R[3 + j] + t[2] * R[6 + j]) > a[0] * AR[j] + a[1] * AR[3 + j] + a[2] * AR[6 + j] + b[j]) return 0;
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    for (int j = 0; j < 3; j++) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const float ra = a[i1] * AR[3 * i2 + j] + a[i2] * AR[3 * i1 + j];
      const float rb = b[j1] * AR[3 * i + j2] + b[j2] * AR[3 * i + j1];
      if (fabsf(t[i2] * R[3 * i1 + j] - t[i1] * R[3 * i2 + j]) > ra + rb) return 0;
    }
  }
  return 1;
}

/* plane against anything: one lane per pair; writes up to 4 (dist,pos,normal) records to out[7*i..]; returns count */
RG_DEV_NOINLINE int rg_narrow_plane(const RgCtx c, int g1, int g2, float margin, float* out) {
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  const int t1 = m.geom_type[g1], t2 = m.geom_type[g2];
  int cnt = 0;
  {
    RgGeomView pl, o;
    rg_geom_view(c, g1, 0.0f, pl);
    rg_geom_view(c, g2, 0.0f, o);
    const float n[3] = {pl.mat[2], pl.mat[5], pl.mat[8]};
    if (t2 == RG_GEOM_BOX) {
      for (int k = 0; k < 8 && cnt < 4; k++) {
        const float loc[3] = {(k & 1) ? o.size[0] : -o.size[0], (k & 2) ? o.size[1] : -o.size[1], (k & 4) ? o.size[2] : -o.size[2]};
        float w[3], dif[3];
        rg_mulmat3(w, o.mat, loc); rg_add3(w, w, o.pos);
        rg_sub3(dif, w, pl.pos);
        const float dist = rg_dot3(dif, n);
        if (dist > margin) continue;
        float* r = out + 7 * cnt;
        r[0] = dist;
        r[1] = w[0] - 0.5f * dist * n[0]; r[2] = w[1] - 0.5f * dist * n[1]; r[3] = w[2] - 0.5f * dist * n[2];
        rg_copy3(r + 4, n);
        cnt++;
      }
    } else if (t2 == RG_GEOM_SPHERE) {
      float dif[3];
      rg_sub3(dif, o.pos, pl.pos);
      const float dist = rg_dot3(dif, n) - o.size[0];
      if (dist <= margin) {
        float* r = out;
        r[0] = dist;
        for (int i = 0; i < 3; i++) r[1 + i] = o.pos[i] - (o.size[0] + 0.5f * dist) * n[i];
        rg_copy3(r + 4, n);
        cnt = 1;
      }
    } else {
      float f1[3], f2[3];
      rg_make_frame(n, f1, f2);
      for (int k = 0; k < 4; k++) {
        float dir[3] = {-n[0], -n[1], -n[2]};
        if (k > 0) {
          const float ang = 2.0943951f * (float)(k - 1);
          rg_addscl3(dir, f1, 0.05f * cosf(ang));
          rg_addscl3(dir, f2, 0.05f * sinf(ang));
          rg_normalize3(dir);
        }
        float w[3], dif[3];
        rg_support(m, o, dir, w);
        rg_sub3(dif, w, pl.pos);
        const float dist = rg_dot3(dif, n);
        if (dist > margin) continue;
        int dup = 0;
        for (int q = 0; q < cnt; q++) {
          /* recorded pos = w - 0.5 dist n: compare the vertices */
          float e[3] = {out[7 * q + 1] + 0.5f * out[7 * q] * n[0] - w[0], out[7 * q + 2] + 0.5f * out[7 * q] * n[1] - w[1], out[7 * q + 3] + 0.5f * out[7 * q] * n[2] - w[2]};
          if (rg_dot3(e, e) < 1e-12f) dup = 1;
        }
        if (dup) continue;
        float* r = out + 7 * cnt;
        r[0] = dist;
        r[1] = w[0] - 0.5f * dist * n[0]; r[2] = w[1] - 0.5f * dist * n[1]; r[3] = w[2] - 0.5f * dist * n[2];
        rg_copy3(r + 4, n);
        cnt++;
      }
    }
    return cnt;
  }
  return 0;
}

/* ---- box-box: separating-axis test over the 15 candidate axes, then a contact manifold (same construction as the oracle's
 * box_box, oracle/rgo_collision.inc): face contact = the opposing face of the other box clipped against the reference
 * face's side planes, every vertex within `margin` of the reference face is a contact (at most 4: the extreme ones along
 * the face's two tangents); edge-edge contact = the closest points of the two edges.  The 1-point MPR answer cannot hold a
 * block flat on a table or on another block (robogym/envs/rearrange/holdouts/tests/test_stability.py:215-261).  One lane
 * per pair; out[7*i..] = dist, pos[3], normal[3] (from box 1 to box 2); returns the number of contacts. */
RG_DEV_NOINLINE int rg_box_box(const RgCtx c, int g1, int g2, float margin, float* out) {
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  RgGeomView va, vb;
  rg_geom_view(c, g1, 0.0f, va);
  rg_geom_view(c, g2, 0.0f, vb);
  const float* pa = va.pos; const float* pb = vb.pos; const float* a = va.size; const float* b = vb.size;
  float A[3][3], B[3][3], d[3], C[3][3], AC[3][3];
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { A[i][k] = va.mat[3 * k + i]; B[i][k] = vb.mat[3 * k + i]; }
  rg_sub3(d, pb, pa);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { C[i][j] = rg_dot3(A[i], B[j]); AC[i][j] = fabsf(C[i][j]); }
  float best = -3.0e38f, bestn[3] = {0, 0, 0};
  int code = -1;
  for (int i = 0; i < 3; i++) {
    const float t = rg_dot3(A[i], d), s = fabsf(t) - (a[i] + b[0] * AC[i][0] + b[1] * AC[i][1] + b[2] * AC[i][2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = i; for (int k = 0; k < 3; k++) bestn[k] = t < 0 ? -A[i][k] : A[i][k]; }
  }
  for (int j = 0; j < 3; j++) {
    const float t = rg_dot3(B[j], d), s = fabsf(t) - (b[j] + a[0] * AC[0][j] + a[1] * AC[1][j] + a[2] * AC[2][j]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = 3 + j; for (int k = 0; k < 3; k++) bestn[k] = t < 0 ? -B[j][k] : B[j][k]; }
  }
  float beste = -3.0e38f, en[3] = {0, 0, 0};
  int ecode = -1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      float Lx[3];
      rg_cross(Lx, A[i], B[j]);
      const float len = sqrtf(rg_dot3(Lx, Lx));
      if (len < 1e-6f) continue;
      for (int k = 0; k < 3; k++) Lx[k] /= len;
      const float t = rg_dot3(Lx, d);
      float ra = 0.0f, rb = 0.0f;
      for (int k = 0; k < 3; k++) { ra += a[k] * fabsf(rg_dot3(Lx, A[k])); rb += b[k] * fabsf(rg_dot3(Lx, B[k])); }
      const float s = fabsf(t) - (ra + rb);
      if (s > margin) return 0;
      if (s > beste) { beste = s; ecode = 6 + 3 * i + j; for (int k = 0; k < 3; k++) en[k] = t < 0 ? -Lx[k] : Lx[k]; }
    }
  if (ecode >= 0 && beste > best + 1e-6f) {
    const int i = (ecode - 6) / 3, j = (ecode - 6) % 3;
    float ea[3], eb[3];
    rg_copy3(ea, pa); rg_copy3(eb, pb);
    for (int k = 0; k < 3; k++) {
      if (k != i) rg_addscl3(ea, A[k], rg_dot3(en, A[k]) > 0 ? a[k] : -a[k]);
      if (k != j) rg_addscl3(eb, B[k], rg_dot3(en, B[k]) > 0 ? -b[k] : b[k]);
    }
    float w[3];
    rg_sub3(w, ea, eb);
    const float uv = C[i][j], uw = rg_dot3(A[i], w), vw = rg_dot3(B[j], w), den = 1.0f - uv * uv;
    float sa = den > 1e-12f ? (uv * vw - uw) / den : 0.0f, tb = den > 1e-12f ? (vw - uv * uw) / den : 0.0f;
    sa = rg_clamp(sa, -a[i], a[i]); tb = rg_clamp(tb, -b[j], b[j]);
    out[0] = beste;
    for (int k = 0; k < 3; k++) { out[1 + k] = 0.5f * (ea[k] + sa * A[i][k] + eb[k] + tb * B[j][k]); out[4 + k] = en[k]; }
    return 1;
  }
  const int refA = code < 3, ax = refA ? code : code - 3;
  const float* pr = refA ? pa : pb; const float* pi = refA ? pb : pa;
  const float* hr = refA ? a : b; const float* hi = refA ? b : a;
  float (*Rr)[3] = refA ? A : B; float (*Ri)[3] = refA ? B : A;
  float nr[3];
  for (int k = 0; k < 3; k++) nr[k] = refA ? bestn[k] : -bestn[k];
  int iax = 0; float mind = 3.0e38f, isg = 1.0f;
  for (int k = 0; k < 3; k++) { const float t = rg_dot3(Ri[k], nr); if (-fabsf(t) < mind) { mind = -fabsf(t); iax = k; isg = t > 0 ? -1.0f : 1.0f; } }
  const int i1 = (iax + 1) % 3, i2 = (iax + 2) % 3, r1 = (ax + 1) % 3, r2 = (ax + 2) % 3;
  float poly[16][3], tmp[16][3], fc[3];
  rg_copy3(fc, pi); rg_addscl3(fc, Ri[iax], isg * hi[iax]);
  for (int q = 0; q < 4; q++) {
    rg_copy3(poly[q], fc);
    rg_addscl3(poly[q], Ri[i1], (q == 0 || q == 3) ? hi[i1] : -hi[i1]);
    rg_addscl3(poly[q], Ri[i2], (q < 2) ? hi[i2] : -hi[i2]);
  }
  int np = 4;
  for (int side = 0; side < 4 && np > 0; side++) {
    const int r = side < 2 ? r1 : r2;
    const float sg = (side & 1) ? -1.0f : 1.0f;
    int nn = 0;
    for (int q = 0; q < np; q++) {
      const float* P = poly[q]; const float* Q = poly[(q + 1) % np];
      float dp[3], dq[3];
      rg_sub3(dp, P, pr); rg_sub3(dq, Q, pr);
      const float fp = sg * rg_dot3(dp, Rr[r]) - hr[r], fq = sg * rg_dot3(dq, Rr[r]) - hr[r];
      const int inp = fp <= 1e-6f, inq = fq <= 1e-6f;   /* a vertex within a micron of the plane counts as inside (see the oracle) */
      if (inp) { rg_copy3(tmp[nn], P); nn++; }
      if (inp != inq) {
        const float t = fp / (fp - fq);
        for (int k = 0; k < 3; k++) tmp[nn][k] = P[k] + t * (Q[k] - P[k]);
        nn++;
      }
    }
    np = nn;
    for (int q = 0; q < np; q++) rg_copy3(poly[q], tmp[q]);
  }
  float dist[16], u[16], v[16];
  int keep[16], nk = 0;
  for (int q = 0; q < np; q++) {
    float dq[3];
    rg_sub3(dq, poly[q], pr);
    dist[q] = rg_dot3(dq, nr) - hr[ax];
    u[q] = rg_dot3(dq, Rr[r1]); v[q] = rg_dot3(dq, Rr[r2]);
    if (dist[q] <= margin) keep[nk++] = q;
  }
  if (nk > 4) {
    int sel[4] = {keep[0], keep[0], keep[0], keep[0]};
    for (int k = 0; k < nk; k++) {
      const int q = keep[k];
      /* extremes along two skewed directions: the edges of the usual polygons (axis-aligned or 45-degree rectangles)
         are never perpendicular to them, so every extreme is a single vertex and rounding cannot pick another one */
      const float k1 = u[q] + 0.31f * v[q], k2 = v[q] - 0.31f * u[q];
      if (k1 < u[sel[0]] + 0.31f * v[sel[0]]) sel[0] = q;
      if (k1 > u[sel[1]] + 0.31f * v[sel[1]]) sel[1] = q;
      if (k2 < v[sel[2]] - 0.31f * u[sel[2]]) sel[2] = q;
      if (k2 > v[sel[3]] - 0.31f * u[sel[3]]) sel[3] = q;
    }
    nk = 0;
    for (int k = 0; k < 4; k++) { int dup = 0; for (int l = 0; l < nk; l++) dup |= keep[l] == sel[k]; if (!dup) keep[nk++] = sel[k]; }
  }
  for (int k = 0; k < nk; k++) {
    const int q = keep[k];
    float* o = out + 7 * k;
    o[0] = dist[q];
    for (int cc = 0; cc < 3; cc++) { o[1 + cc] = poly[q][cc] - 0.5f * dist[q] * nr[cc]; o[4 + cc] = bestn[cc]; }
  }
  return nk;
}

/* append the first `n` survivors (flag per lane) of list `src` to list `dst`, then drop them from `src` */
RG_DEV void rg_pair(const RG_MODEL_T& m, int k, int& g1, int& g2) {
  if (RG_HAS_PAIRS(m)) { const unsigned p = m.pair_packed[k]; g1 = (int)(p & 255u); g2 = (int)(p >> 8); }
  else { g1 = RG_LDG(m.pair_geom1 + k); g2 = RG_LDG(m.pair_geom2 + k); }
}
/* Separating-axis cache: pairs that pass the bounding-box cull but do not touch (most of them: neighbouring finger links)
 * stay separated along nearly the same axis from one substep to the next.  The axis MPR stopped with is remembered per
 * pair and tried first next time: one support evaluation proves the separation instead of a portal search.  The test
 * is the same "support . dir < 0" MPR itself exits on, so the outcome (no contact) is the one MPR would reach.
 * Table: RG_NSEP / 2 sets x 2 ways, one word per entry = pair id (12 bits, 0xfff = empty) | axis in octahedral
 * coordinates (2 x 10 bits); most recently written entry in way 0. */
RG_DEV int rg_sep_set(int pair) { return 2 * ((int)(((unsigned)pair * 2654435761u) >> 16) & (RG_NSEP / 2 - 1)); }
RG_DEV int rg_sep_pack(int pair, const float* d) {
  const float inv = 1.0f / (fabsf(d[0]) + fabsf(d[1]) + fabsf(d[2]) + 1e-30f);
  float px = d[0] * inv, py = d[1] * inv;
  if (d[2] < 0.0f) { const float qx = (1.0f - fabsf(py)) * (px >= 0.0f ? 1.0f : -1.0f), qy = (1.0f - fabsf(px)) * (py >= 0.0f ? 1.0f : -1.0f); px = qx; py = qy; }
  const int x = (int)(px * 511.0f + 511.5f), y = (int)(py * 511.0f + 511.5f);
  return (pair & 0xfff) | (x << 12) | (y << 22);
}
RG_DEV void rg_sep_unpack(int w, float* d) {
  const float px = (float)(((w >> 12) & 1023) - 511) * (1.0f / 511.0f), py = (float)(((w >> 22) & 1023) - 511) * (1.0f / 511.0f);
  const float z = 1.0f - fabsf(px) - fabsf(py);
  d[0] = px; d[1] = py; d[2] = z;
  if (z < 0.0f) { d[0] = (1.0f - fabsf(py)) * (px >= 0.0f ? 1.0f : -1.0f); d[1] = (1.0f - fabsf(px)) * (py >= 0.0f ? 1.0f : -1.0f); }
  rg_normalize3(d);
}

/* Convex-convex narrow phase of one batch of candidates: RG_GRP lanes per pair, 32 / RG_GRP pairs in flight.  Every loop
 * trip is one MPR iteration of each pair in flight: the lanes of a group scan the two hulls together (support mapping),
 * reduce to the winning vertices, and then each of them advances its copy of the pair's MPR state.  A group that finishes
 * a pair takes the next one from the list, so the trip count is about (total iterations) / (pairs in flight) rather than
 * the iteration count of the slowest pair times the number of rounds.
 * `list[0..nconv)` = candidate slots to process; results go to stage[8 * slot ..] = {count, dist, pos[3], normal[3]}. */
RG_DEV_NOINLINE void rg_mpr_batch(const RgCtx c, const int* cand2, const int* list, int nconv, float* stage) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  const float tol = m.opt_mpr_tolerance[0];
  const int maxiter = m.opt_mpr_iterations[0];
  int* sep = (int*)(RG_SCRATCH(c) + RG_CL(c).sep);
  LANEVAR(RgMpr, st);
  LANEVAR(int, want); LANEVAR(int, wpos); LANEVAR(int, busy);
  LANEVAR(float, b1); LANEVAR(int, i1); LANEVAR(float, b2); LANEVAR(int, i2);
  int next = 0, wtot;
  RG_PHASE_BEGIN
  LV(st).state = RG_MPR_IDLE;
  RG_PHASE_END
  for (;;) {
    RG_PHASE_BEGIN
    LV(want) = LV(st).state == RG_MPR_IDLE ? 1 : 0;   /* all RG_GRP lanes of a group agree, so the scan counts groups x RG_GRP */
    RG_PHASE_END
    RG_WARP_SCAN(want, wpos, wtot);
    RG_PHASE_BEGIN
    RgMpr& S = LV(st);
    if (S.state == RG_MPR_IDLE) {
      const int idx = next + LV(wpos) / RG_GRP;
      if (idx < nconv) {
        S.slot = list[idx];
        int g1, g2;
        rg_pair(m, cand2[S.slot], g1, g2);
        const float margin = fmaxf(m.geom_margin[g1], m.geom_margin[g2]);
        rg_geom_view(c, g1, margin, S.o1);
        rg_geom_view(c, g2, margin, S.o2);
        rg_mpr_begin(S);
        S.pair = cand2[S.slot];
        if (S.pair < 0xfff) {
          const int* se = sep + rg_sep_set(S.pair);
          const int w0 = se[0], w1 = se[1];
          if ((w0 & 0xfff) == S.pair) { rg_sep_unpack(w0, S.dir); S.state = RG_MPR_PRE; }
          else if ((w1 & 0xfff) == S.pair) { rg_sep_unpack(w1, S.dir); S.state = RG_MPR_PRE; }
        }
        RG_STAT(if ((lane & (RG_GRP - 1)) == 0) rg_stat_mpr++;)
      } else S.state = RG_MPR_FINISHED;
    }
    LV(b1) = 0.0f; LV(i1) = 0; LV(b2) = 0.0f; LV(i2) = 0;
    if (S.state < RG_MPR_DONE) {
      /* this lane's share of both hull scans */
      const int gl = lane & (RG_GRP - 1);
      float d1[3], d2[3];
      rg_mulmatT3(d1, S.o1.mat, S.dir);
      rg_mulmatT3(d2, S.o2.mat, S.dir);
      d2[0] = -d2[0]; d2[1] = -d2[1]; d2[2] = -d2[2];
      /* one hull after the other: a fused loop over both (more loads in flight) measured SLOWER (2.17 M vs 1.84 M cycles per
         env-step in the narrow phase): a box on one side makes half of its loads useless, and the body gets bigger */
      if (S.o1.type == RG_GEOM_MESH) rg_hull_scan(m, S.o1, d1, gl, RG_GRP, LV(b1), LV(i1));
      if (S.o2.type == RG_GEOM_MESH) rg_hull_scan(m, S.o2, d2, gl, RG_GRP, LV(b2), LV(i2));
    }
    LV(busy) = S.state < RG_MPR_DONE;
    RG_PHASE_END
    next += wtot / RG_GRP;
#ifdef RG_MPR_CTA_TRIPS
    if (!RG_CTA_ANY(RG_WARP_OR(busy))) break;   /* every warp of the CTA runs the same number of trips (one barrier per trip) */
#else
    if (!RG_WARP_OR(busy)) break;
#endif
    RG_STAT(rg_stat_x[4]++;)
    RG_GROUP_ARGMAX(b1, i1);
    RG_GROUP_ARGMAX(b2, i2);
    RG_PHASE_BEGIN
    RgMpr& S = LV(st);
    if (S.state < RG_MPR_DONE) {
      RgSup sp;
      float dl[3], loc[3], nd[3] = {-S.dir[0], -S.dir[1], -S.dir[2]};
      RG_STAT(if ((lane & (RG_GRP - 1)) == 0) { rg_stat_support += 2; rg_stat_x[9]++; })
      if (S.o1.type == RG_GEOM_MESH) { float p[4]; RG_LDG4(m.mesh_vert4, S.o1.vadr + LV(i1), p); rg_copy3(loc, p); }
      else { rg_mulmatT3(dl, S.o1.mat, S.dir); rg_support_prim(S.o1, dl, loc); }
      rg_support_world(S.o1, loc, S.dir, sp.v1);
      if (S.o2.type == RG_GEOM_MESH) { float p[4]; RG_LDG4(m.mesh_vert4, S.o2.vadr + LV(i2), p); rg_copy3(loc, p); }
      else { rg_mulmatT3(dl, S.o2.mat, nd); rg_support_prim(S.o2, dl, loc); }
      rg_support_world(S.o2, loc, nd, sp.v2);
      rg_sub3(sp.v, sp.v1, sp.v2);
      rg_mpr_advance(S, sp, tol, maxiter);
      if (S.state == RG_MPR_DONE) {
        RG_STAT(if ((lane & (RG_GRP - 1)) == 0) { rg_stat_hist[S.depth >= 0][S.it < 63 ? S.it : 63]++; if (S.depth >= 0) rg_stat_mpr_hit++; })
        if ((lane & (RG_GRP - 1)) == 0) {
          if (S.depth < 0 && S.lastdot < 0 && S.pair < 0xfff) {
            int* se = sep + rg_sep_set(S.pair);
            if ((se[0] & 0xfff) != S.pair) se[1] = se[0];
            se[0] = rg_sep_pack(S.pair, S.dir);
          }
          float* o = stage + 8 * S.slot;
          const int hit = S.depth >= 0 && rg_dot3(S.odir, S.odir) >= 0.5f;
          o[0] = hit ? 1.0f : 0.0f;
          o[1] = 2.0f * S.o1.halfmargin - S.depth;
          rg_copy3(o + 2, S.opos);
          rg_copy3(o + 5, S.odir);
        }
        S.state = RG_MPR_IDLE;
      }
    }
    RG_PHASE_END
  }
}

RG_DEV_NOINLINE void rg_collision(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  int* cand = (int*)(s + L.cand);
  int* cand2 = (int*)(s + L.cand2);
  int ncon = 0, n1 = 0, n2 = 0, k0 = 0, warn = 0, work = 0;
  const int cap = (m.nconmax > 0 && m.nconmax < L.ncon) ? m.nconmax : L.ncon;
  const int enabled = !(m.opt_disableflags[0] & (RG_DSBL_CONTACT | RG_DSBL_CONSTRAINT));
  RG_PROF_BEGIN
  while (enabled && (k0 < m.npair || n1 > 0 || n2 > 0)) {
    /* stage A: bounding spheres over the static pair list -> cand */
    RG_PROFC(c, 15)
    while (n1 < 32 && k0 < m.npair) {
      LANEVAR(int, pred); LANEVAR(int, pos);
      int tot;
      RG_PHASE_BEGIN
      const int k = k0 + lane;
      int pr = 0;
      if (k < m.npair) {
        int g1, g2;
        rg_pair(m, k, g1, g2);
        const float margin = fmaxf(m.geom_margin[g1], m.geom_margin[g2]);
        float dif[3];
        rg_sub3(dif, s + L.gxpos + 3 * g2, s + L.gxpos + 3 * g1);
        if (m.geom_type[g1] == RG_GEOM_PLANE) {
          const int b = m.geom_bodyid[g1];
          float q[4], z[3] = {0, 0, 1}, n[3];
          rg_quat_mul(q, s + L.xquat + 4 * b, m.geom_quat + 4 * g1);
          rg_rot(n, q, z);
          pr = rg_dot3(dif, n) <= m.geom_rbound[g2] + margin;
        } else {
          const float bound = m.geom_rbound[g1] + m.geom_rbound[g2] + margin;
          pr = rg_dot3(dif, dif) <= bound * bound;
        }
      }
      LV(pred) = pr;
      RG_PHASE_END
      RG_WARP_SCAN(pred, pos, tot);
      RG_PHASE_BEGIN
      if (LV(pred)) cand[n1 + LV(pos)] = k0 + lane;
      RG_PHASE_END
      n1 += tot;
      RG_STAT(rg_stat_x[5] += tot;)
      k0 += 32;
    }
    RG_PROFC(c, 9)
    /* stage B: oriented bounding boxes on up to 32 candidates -> cand2 */
    if (n1 > 0 && n2 < 32) {
      const int n = n1 < 32 ? n1 : 32;
      LANEVAR(int, pred); LANEVAR(int, pos); LANEVAR(int, keep); LANEVAR(int, mine);
      int tot;
      RG_PHASE_BEGIN
      int pr = 0, k = -1;
      if (lane < n) {
        k = cand[lane];
        int g1, g2;
        rg_pair(m, k, g1, g2);
        pr = m.geom_type[g1] == RG_GEOM_PLANE ? 1 : rg_obb_overlap(c, g1, g2, fmaxf(m.geom_margin[g1], m.geom_margin[g2]));
      }
      LV(pred) = pr; LV(mine) = k;
      LV(keep) = (lane + 32 < n1) ? cand[lane + 32] : -1;
      RG_PHASE_END
      RG_WARP_SCAN(pred, pos, tot);
      RG_PHASE_BEGIN
      if (LV(pred)) cand2[n2 + LV(pos)] = LV(mine);
      if (LV(keep) >= 0) cand[lane] = LV(keep);
      RG_PHASE_END
      n2 += tot;
      RG_STAT(rg_stat_x[6] += tot;)
      n1 = n1 > 32 ? n1 - 32 : 0;
    }
    RG_PROFC(c, 10)
    /* stage C: narrow phase, one pair per lane, once a full warp of work is queued (or at the end) */
    if (n2 >= 32 || (n2 > 0 && k0 >= m.npair && n1 == 0)) {
      const int n = n2 < 32 ? n2 : 32;
      work += n;
      LANEVAR(int, cnt); LANEVAR(int, cpos); LANEVAR(int, keep); LANEVAR(int, conv); LANEVAR(int, vpos);
      LANEARR(float, cb, 28);
      int tot, nconv;
      float* stage = s + L.stage;              /* [32][8] results of the convex-convex pairs */
      int* clist = (int*)(s + L.stage) + 256;  /* their candidate slots */
      /* convex-convex pairs: grouped MPR */
      RG_PHASE_BEGIN
      int cv = 0;
      if (lane < n) {
        int g1, g2;
        rg_pair(m, cand2[lane], g1, g2);
#ifdef RG_NO_BOXBOX   /* A/B measurement only: box-box pairs through MPR (1 point) */
        cv = m.geom_type[g1] != RG_GEOM_PLANE;
#else
        cv = m.geom_type[g1] != RG_GEOM_PLANE && !(m.geom_type[g1] == RG_GEOM_BOX && m.geom_type[g2] == RG_GEOM_BOX);
#endif
      }
      LV(conv) = cv;
      RG_PHASE_END
      RG_WARP_SCAN(conv, vpos, nconv);
      RG_PHASE_BEGIN
      if (LV(conv)) clist[LV(vpos)] = lane;
      RG_PHASE_END
      if (nconv > 0) rg_mpr_batch(c, cand2, clist, nconv, stage);
      /* planes: one lane per pair */
      RG_PHASE_BEGIN
      int cn = 0;
      if (lane < n) {
        if (LV(conv)) {
          const float* o = stage + 8 * lane;
          cn = (int)o[0];
          for (int q = 0; q < 7; q++) LA(cb, q) = o[1 + q];
        } else {
          const int k = cand2[lane];
          int g1, g2;
          rg_pair(m, k, g1, g2);
          const float mg = fmaxf(m.geom_margin[g1], m.geom_margin[g2]);
          cn = m.geom_type[g1] == RG_GEOM_PLANE ? rg_narrow_plane(c, g1, g2, mg, &LA(cb, 0)) : rg_box_box(c, g1, g2, mg, &LA(cb, 0));
        }
      }
      LV(cnt) = cn;
      LV(keep) = (lane + 32 < n2) ? cand2[lane + 32] : -1;
      RG_PHASE_END
      RG_WARP_SCAN(cnt, cpos, tot);
      RG_PHASE_BEGIN
      if (lane < n && LV(cnt) > 0) {
        const int k = cand2[lane];
        int g1, g2;
        rg_pair(m, k, g1, g2);
        /* mixed contact parameters: max condim / friction, solmix-weighted solref / solimp */
        int condim = m.geom_condim[g1] > m.geom_condim[g2] ? m.geom_condim[g1] : m.geom_condim[g2];
        float fri[3], solref[2], solimp[5];
        const int p1 = m.geom_priority[g1], p2 = m.geom_priority[g2];
        if (p1 != p2) {
          const int gp = p1 > p2 ? g1 : g2;
          condim = m.geom_condim[gp];
          for (int i = 0; i < 3; i++) fri[i] = m.geom_friction[3 * gp + i];
          for (int i = 0; i < 2; i++) solref[i] = m.geom_solref[2 * gp + i];
          for (int i = 0; i < 5; i++) solimp[i] = m.geom_solimp[5 * gp + i];
        } else {
          for (int i = 0; i < 3; i++) fri[i] = fmaxf(m.geom_friction[3 * g1 + i], m.geom_friction[3 * g2 + i]);
          const float s1 = m.geom_solmix[g1], s2 = m.geom_solmix[g2];
          float mix;
          if (s1 >= RG_MINVAL && s2 >= RG_MINVAL) mix = s1 / (s1 + s2);
          else if (s1 < RG_MINVAL && s2 < RG_MINVAL) mix = 0.5f;
          else mix = s1 < RG_MINVAL ? 0.0f : 1.0f;
          const float* r1 = m.geom_solref + 2 * g1; const float* r2 = m.geom_solref + 2 * g2;
          if (r1[0] > 0 && r2[0] > 0) for (int i = 0; i < 2; i++) solref[i] = mix * r1[i] + (1 - mix) * r2[i];
          else for (int i = 0; i < 2; i++) solref[i] = fminf(r1[i], r2[i]);
          for (int i = 0; i < 5; i++) solimp[i] = mix * m.geom_solimp[5 * g1 + i] + (1 - mix) * m.geom_solimp[5 * g2 + i];
        }
        const float margin = fmaxf(m.geom_margin[g1], m.geom_margin[g2]), gap = fmaxf(m.geom_gap[g1], m.geom_gap[g2]);
        for (int i = 0; i < LV(cnt); i++) {
          const int idx = ncon + LV(cpos) + i;
          if (idx >= cap) break;
          float* r = s + L.con + RG_CON_STRIDE * idx;
          const float* o = &LA(cb, 7 * i);
          r[0] = o[0];
          rg_copy3(r + 1, o + 1);
          rg_copy3(r + 4, o + 4);
          rg_make_frame(r + 4, r + 7, r + 10);
          r[13] = margin - gap;
          r[14] = fri[0]; r[15] = fri[1]; r[16] = fri[2];
          r[17] = (float)condim;
          r[18] = (float)m.geom_bodyid[g1]; r[19] = (float)m.geom_bodyid[g2];
          r[20] = (float)g1; r[21] = (float)g2;
          r[22] = solref[0]; r[23] = solref[1];
          for (int q = 0; q < 5; q++) s[L.cprm + RG_CPRM * idx + q] = solimp[q];   /* consumed by rg_make_constraints */
        }
      }
      RG_PHASE_END
      if (ncon + tot > cap) { warn |= RG_WARN_CONTACT_FULL; ncon = cap; } else ncon += tot;
      RG_PHASE_BEGIN
      if (LV(keep) >= 0) cand2[lane] = LV(keep);
      RG_PHASE_END
      n2 = n2 > 32 ? n2 - 32 : 0;
      RG_PROFC(c, 12)
    }
    RG_PROFC(c, 11)
  }
  RG_PHASE_BEGIN
  RG_STAT(if (lane == 0) rg_stat_x[7] += ncon;)
  if (lane == 0) { RG_SI(c, RG_S_NCON) = ncon; RG_SI(c, RG_S_WARN) |= warn; RG_SI(c, RG_S_WORK) += work; }
  RG_PHASE_END
}
