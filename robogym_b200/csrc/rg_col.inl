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
static long long rg_stat_hist[2][64] = {{0}}; static long long rg_stat_support = 0, rg_stat_climb = 0, rg_stat_mpr = 0, rg_stat_mpr_hit = 0, rg_stat_narrow = 0, rg_stat_maxsup = 0, rg_stat_cur = 0;
#define RG_STAT(x) x
#else
#define RG_STAT(x)
#endif
struct RgGeomView {
  float pos[3], mat[9], size[3];
  int type, vadr, vnum, mid;
  int hint;              /* adjacency range of the last support vertex (hill-climb warm start), -1 = pick an extreme vertex */
  float hv[3];           /* ... and its local coordinates */
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
  v.hint = -1;
  v.hv[0] = v.hv[1] = v.hv[2] = 0.0f;
  v.halfmargin = 0.5f * margin;
  v.vadr = 0; v.vnum = 0; v.mid = 0;
  if (v.type == RG_GEOM_MESH) { const int mid = m.geom_dataid[g]; v.mid = mid; v.vadr = m.mesh_vertadr[mid]; v.vnum = m.mesh_vertnum[mid]; }
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

/* ---- hull support mapping: steepest-ascent hill climb on the convex hull's edge graph.  A vertex is known by its
 * adjacency range (first entry | degree << 20) in mesh_nbr; every entry carries the neighbour's coordinates AND the
 * neighbour's own range, so one climb step is ONE level of dependent loads, and the RG_CLIMB_W entries of a step are
 * fetched together (independent loads in flight) before any of them is compared. */
#define RG_CLIMB_W 6
struct RgClimb { int pk; float b[3], best; };

RG_DEV void rg_climb_init(const RG_MODEL_T& m, const RgGeomView& v, const float* dl, RgClimb& st) {
  if (v.hint >= 0) { st.pk = v.hint; rg_copy3(st.b, v.hv); }
  else {
    const float ax = fabsf(dl[0]), ay = fabsf(dl[1]), az = fabsf(dl[2]);
    const int axis = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    float n4[4];
    RG_LDG4(m.mesh_ext, 6 * v.mid + 2 * axis + (dl[axis] >= 0 ? 0 : 1), n4);
    rg_copy3(st.b, n4); st.pk = rg_f2i(n4[3]);
  }
  st.best = st.b[0] * dl[0] + st.b[1] * dl[1] + st.b[2] * dl[2];
}
/* entries [base, base + RG_CLIMB_W) of the current vertex (clamped to its last entry: a repeated entry never wins the
   strict comparison twice, so the visiting order -- and with it every tie-break -- is that of a plain loop) */
RG_DEV void rg_climb_load(const RG_MODEL_T& m, int pk, int base, float n[RG_CLIMB_W][4]) {
  const int a0 = pk & 0xFFFFF, last = (int)((unsigned)pk >> 20) - 1;
#pragma unroll
  for (int i = 0; i < RG_CLIMB_W; i++) { const int k = base + i < last ? base + i : last; RG_LDG4(m.mesh_nbr, a0 + k, n[i]); }
}
RG_DEV void rg_climb_eval(const float n[RG_CLIMB_W][4], const float* dl, RgClimb& st, int& nxt) {
#pragma unroll
  for (int i = 0; i < RG_CLIMB_W; i++) {
    const float dd = n[i][0] * dl[0] + n[i][1] * dl[1] + n[i][2] * dl[2];
    if (dd > st.best) { st.best = dd; nxt = rg_f2i(n[i][3]); st.b[0] = n[i][0]; st.b[1] = n[i][1]; st.b[2] = n[i][2]; }
  }
}
/* one step; returns 1 when the climb moved to a better neighbour */
RG_DEV int rg_climb_step(const RG_MODEL_T& m, const float* dl, RgClimb& st) {
  const int deg = (int)((unsigned)st.pk >> 20);
  int nxt = st.pk;
  for (int base = 0; base < deg; base += RG_CLIMB_W) {
    float n[RG_CLIMB_W][4];
    rg_climb_load(m, st.pk, base, n);
    rg_climb_eval(n, dl, st, nxt);
  }
  const int moved = nxt != st.pk;
  st.pk = nxt;
  RG_STAT(rg_stat_climb += moved;)
  return moved;
}

RG_DEV void rg_support_world(const RgGeomView& v, const float* loc, const float* dir, float* res) {
  rg_mulmat3(res, v.mat, loc);
  res[0] += v.pos[0] + dir[0] * v.halfmargin;
  res[1] += v.pos[1] + dir[1] * v.halfmargin;
  res[2] += v.pos[2] + dir[2] * v.halfmargin;
}

RG_DEV void rg_support(const RG_MODEL_T& m, RgGeomView& v, const float* dir, float* res) {
  float dl[3], loc[3];
  RG_STAT(rg_stat_support++; rg_stat_cur++;)
  rg_mulmatT3(dl, v.mat, dir);
  if (v.type == RG_GEOM_MESH) {
    RgClimb st;
    rg_climb_init(m, v, dl, st);
    for (int guard = 0; guard < v.vnum; guard++) if (!rg_climb_step(m, dl, st)) break;
    v.hint = st.pk; rg_copy3(v.hv, st.b);
    rg_copy3(loc, st.b);
  } else rg_support_prim(v, dl, loc);
  rg_support_world(v, loc, dir, res);
}

struct RgSup { float v[3], v1[3], v2[3]; };

/* support of the Minkowski difference o1 - o2.  For two hulls the two climbs advance together, so that the loads of
   both are in flight at the same time (the narrow phase is a chain of dependent L2 round trips, not arithmetic). */
RG_DEV void rg_mpr_support(const RG_MODEL_T& m, RgGeomView& o1, RgGeomView& o2, const float* dir, RgSup& sp) {
  const float nd[3] = {-dir[0], -dir[1], -dir[2]};
  if (o1.type == RG_GEOM_MESH && o2.type == RG_GEOM_MESH) {
    float d1[3], d2[3];
    RG_STAT(rg_stat_support += 2; rg_stat_cur += 2;)
    rg_mulmatT3(d1, o1.mat, dir);
    rg_mulmatT3(d2, o2.mat, nd);
    RgClimb s1, s2;
    rg_climb_init(m, o1, d1, s1);
    rg_climb_init(m, o2, d2, s2);
    int go1 = 1, go2 = 1;
    for (int guard = 0; guard < o1.vnum + o2.vnum && (go1 | go2); guard++) {
      const int deg1 = go1 ? (int)((unsigned)s1.pk >> 20) : 0, deg2 = go2 ? (int)((unsigned)s2.pk >> 20) : 0;
      int n1 = s1.pk, n2 = s2.pk;
      for (int base = 0; base < deg1 || base < deg2; base += RG_CLIMB_W) {
        float e1[RG_CLIMB_W][4], e2[RG_CLIMB_W][4];
        if (base < deg1) rg_climb_load(m, s1.pk, base, e1);
        if (base < deg2) rg_climb_load(m, s2.pk, base, e2);
        if (base < deg1) rg_climb_eval(e1, d1, s1, n1);
        if (base < deg2) rg_climb_eval(e2, d2, s2, n2);
      }
      go1 = go1 && n1 != s1.pk; go2 = go2 && n2 != s2.pk;
      RG_STAT(rg_stat_climb += go1 + go2;)
      s1.pk = n1; s2.pk = n2;
    }
    o1.hint = s1.pk; rg_copy3(o1.hv, s1.b);
    o2.hint = s2.pk; rg_copy3(o2.hv, s2.b);
    rg_support_world(o1, s1.b, dir, sp.v1);
    rg_support_world(o2, s2.b, nd, sp.v2);
  } else {
    rg_support(m, o1, dir, sp.v1);
    rg_support(m, o2, nd, sp.v2);
  }
  rg_sub3(sp.v, sp.v1, sp.v2);
}
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

/* depth >= 0 with dir,pos when the inflated geoms intersect; -1 otherwise */
/* The geom views are built HERE (not handed in by reference): a reference into the caller's frame would pin them in
   local memory, and with ~28 KB of L1 beside the shared-memory scratch every access to them is an L2 round trip. */
struct RgMprOut { float depth, dir[3], pos[3]; };
RG_DEV_NOINLINE RgMprOut rg_mpr(const RgCtx c, int g1, int g2, float margin) {
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  const float tol = m.opt_mpr_tolerance[0];
  const int maxiter = m.opt_mpr_iterations[0];
  RgGeomView o1, o2;
  rg_geom_view(c, g1, margin, o1);
  rg_geom_view(c, g2, margin, o2);
  RgMprOut out;
  float* dir_out = out.dir; float* pos = out.pos;
  dir_out[0] = dir_out[1] = dir_out[2] = 0.0f; pos[0] = pos[1] = pos[2] = 0.0f;
  enum { S_V1 = 0, S_V2 = 1, S_V3 = 2, S_REFINE = 3, S_PENETR = 4, S_DONE = 5 };
  RgSup P1, P2, P3, sp;
  float v0[3], dir[3], va[3], vb[3];
  float result = -1.0f;
  rg_sub3(v0, o1.pos, o2.pos);
  if (rg_mpr_zero(v0[0]) && rg_mpr_zero(v0[1]) && rg_mpr_zero(v0[2])) v0[0] = 1e-5f;
  rg_scl3(dir, v0, -1.0f); rg_normalize3(dir);
  int state = S_V1, pen_it = 0;
  P1 = RgSup(); P2 = RgSup(); P3 = RgSup();
  int it = 0;
  for (; it < 200 + maxiter && state != S_DONE; it++) {
    rg_mpr_support(m, o1, o2, dir, sp);
    const float dot = rg_dot3(sp.v, dir);
    if (state == S_V1) {
      P1 = sp;
      if (dot < 0 || rg_mpr_zero(dot)) { state = S_DONE; continue; }
      rg_cross(dir, v0, P1.v);
      /* fp32: the parallel test must be scale-free (libccd compares the raw squared norm with DBL_EPSILON) */
      if (rg_dot3(dir, dir) <= 1e-12f * rg_dot3(v0, v0) * rg_dot3(P1.v, P1.v)) {
        const float n = sqrtf(rg_dot3(P1.v, P1.v));
        for (int i = 0; i < 3; i++) pos[i] = 0.5f * (P1.v1[i] + P1.v2[i]);
        if (n < RG_EPS) { dir_out[0] = dir_out[1] = dir_out[2] = 0; result = 0.0f; }
        else { rg_scl3(dir_out, P1.v, 1.0f / n); result = n; }
        state = S_DONE;
        continue;
      }
      rg_normalize3(dir);
      state = S_V2;
    } else if (state == S_V2) {
      P2 = sp;
      if (dot < 0 || rg_mpr_zero(dot)) { state = S_DONE; continue; }
      rg_sub3(va, P1.v, v0); rg_sub3(vb, P2.v, v0);
      rg_cross(dir, va, vb); rg_normalize3(dir);
      if (rg_dot3(dir, v0) > 0) { const RgSup t = P1; P1 = P2; P2 = t; rg_scl3(dir, dir, -1.0f); }
      state = S_V3;
    } else if (state == S_V3) {
      P3 = sp;
      if (dot < 0 || rg_mpr_zero(dot)) { state = S_DONE; continue; }
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
        state = (d1 > 0 || rg_mpr_zero(d1)) ? S_PENETR : S_REFINE;
      }
    } else if (state == S_REFINE) {
      if (!(dot > 0 || rg_mpr_zero(dot)) || rg_reach_tol3(P1, P2, P3, sp, dir, tol)) { state = S_DONE; continue; }
      rg_expand_portal3(P1, P2, P3, v0, sp);
      rg_portal_dir3(P1, P2, P3, dir);
      const float d1 = rg_dot3(dir, P1.v);
      if (d1 > 0 || rg_mpr_zero(d1)) state = S_PENETR;
    } else { /* S_PENETR */
      if (rg_reach_tol3(P1, P2, P3, sp, dir, tol) || pen_it > maxiter) {
        float w[3];
        const float depth = sqrtf(rg_tri_closest(P1.v, P2.v, P3.v, w));
        if (depth < RG_EPS) { dir_out[0] = dir_out[1] = dir_out[2] = 0; }
        else rg_scl3(dir_out, w, 1.0f / depth);
        rg_find_pos3(v0, o1.pos, o2.pos, P1, P2, P3, pos);
        result = depth;
        state = S_DONE;
        continue;
      }
      rg_expand_portal3(P1, P2, P3, v0, sp);
      rg_portal_dir3(P1, P2, P3, dir);
      pen_it++;
    }
  }
  RG_STAT(rg_stat_hist[result >= 0][it < 63 ? it : 63]++;)
  out.depth = result;
  return out;
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

RG_DEV_NOINLINE int rg_narrow(const RgCtx c, int g1, int g2, float margin, float* out) {
  const RG_MODEL_T& m = RG_MDEREF(c.mref);
  const int t1 = m.geom_type[g1], t2 = m.geom_type[g2];
  int cnt = 0;
  if (t1 == RG_GEOM_PLANE) {
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
  RG_STAT(rg_stat_mpr++; rg_stat_cur = 0;)
  const RgMprOut r = rg_mpr(c, g1, g2, margin);
  RG_STAT(if (rg_stat_cur > rg_stat_maxsup) rg_stat_maxsup = rg_stat_cur; if (r.depth >= 0) rg_stat_mpr_hit++;)
  if (r.depth < 0 || rg_dot3(r.dir, r.dir) < 0.5f) return 0;
  out[0] = margin - r.depth;
  rg_copy3(out + 1, r.pos);
  rg_copy3(out + 4, r.dir);
  return 1;
}

/* append the first `n` survivors (flag per lane) of list `src` to list `dst`, then drop them from `src` */
RG_DEV void rg_pair(const RG_MODEL_T& m, int k, int& g1, int& g2) {
  if (RG_HAS_PAIRS(m)) { const unsigned p = m.pair_packed[k]; g1 = (int)(p & 255u); g2 = (int)(p >> 8); }
  else { g1 = RG_LDG(m.pair_geom1 + k); g2 = RG_LDG(m.pair_geom2 + k); }
}
RG_DEV_NOINLINE void rg_collision(const RgCtx c) {
  RG_LANE_DECL
  const RG_MODEL_T& m = RG_MDEREF(c.mref); const RgLayout& L = RG_CL(c); float* s = RG_SCRATCH(c);
  int* cand = (int*)(s + L.cand);
  int* cand2 = (int*)(s + L.cand2);
  int ncon = 0, n1 = 0, n2 = 0, k0 = 0, warn = 0, work = 0;
  const int cap = (m.nconmax > 0 && m.nconmax < RG_NCON) ? m.nconmax : RG_NCON;
  const int enabled = !(m.opt_disableflags[0] & (RG_DSBL_CONTACT | RG_DSBL_CONSTRAINT));
  RG_PROF_BEGIN
  while (enabled && (k0 < m.npair || n1 > 0 || n2 > 0)) {
    /* stage A: bounding spheres over the static pair list -> cand */
    RG_PROF(c, 15)
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
      k0 += 32;
    }
    RG_PROF(c, 9)
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
      n1 = n1 > 32 ? n1 - 32 : 0;
    }
    RG_PROF(c, 10)
    /* stage C: narrow phase, one pair per lane, once a full warp of work is queued (or at the end) */
    if (n2 >= 32 || (n2 > 0 && k0 >= m.npair && n1 == 0)) {
      const int n = n2 < 32 ? n2 : 32;
      work += n;
      LANEVAR(int, cnt); LANEVAR(int, cpos); LANEVAR(int, keep);
      LANEARR(float, cb, 28);
      int tot;
      RG_PHASE_BEGIN
      int cn = 0;
      if (lane < n) {
        const int k = cand2[lane];
        int g1, g2;
        rg_pair(m, k, g1, g2);
        cn = rg_narrow(c, g1, g2, fmaxf(m.geom_margin[g1], m.geom_margin[g2]), &LA(cb, 0));
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
          for (int q = 0; q < 5; q++) s[L.cprm + 8 * idx + q] = solimp[q];   /* consumed by rg_make_constraints */
        }
      }
      RG_PHASE_END
      if (ncon + tot > cap) { warn |= RG_WARN_CONTACT_FULL; ncon = cap; } else ncon += tot;
      RG_PHASE_BEGIN
      if (LV(keep) >= 0) cand2[lane] = LV(keep);
      RG_PHASE_END
      n2 = n2 > 32 ? n2 - 32 : 0;
      RG_PROF(c, 12)
    }
    RG_PROF(c, 11)
  }
  RG_PHASE_BEGIN
  if (lane == 0) { RG_SI(c, RG_S_NCON) = ncon; RG_SI(c, RG_S_WARN) |= warn; RG_SI(c, RG_S_WORK) += work; }
  RG_PHASE_END
}
