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

struct RgGeomView {
  float pos[3], mat[9], size[3];
  int type, vadr, vnum;
  int hint;
  float halfmargin;
};

RG_DEV void rg_geom_view(const RgCtx& c, int g, float margin, RgGeomView& v) {
  const RgModel& m = c.m;
  const int b = m.geom_bodyid[g];
  float q[4];
  rg_quat_mul(q, c.s + c.L.xquat + 4 * b, m.geom_quat + 4 * g);
  rg_quat_norm(q);
  rg_quat2mat(v.mat, q);
  rg_copy3(v.pos, c.s + c.L.gxpos + 3 * g);
  rg_copy3(v.size, m.geom_size + 3 * g);
  v.type = m.geom_type[g];
  v.hint = 0;
  v.halfmargin = 0.5f * margin;
  v.vadr = 0; v.vnum = 0;
  if (v.type == RG_GEOM_MESH) { const int mid = m.geom_dataid[g]; v.vadr = m.mesh_vertadr[mid]; v.vnum = m.mesh_vertnum[mid]; }
}

RG_DEV void rg_support(const RgModel& m, RgGeomView& v, const float* dir, float* res) {
  float dl[3], loc[3] = {0, 0, 0};
  rg_mulmatT3(dl, v.mat, dir);
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
    case RG_GEOM_MESH: {
      /* steepest-ascent hill climb on the convex hull's edge graph */
      const float* vert = m.mesh_vert + 3 * v.vadr;
      const int* adjadr = m.mesh_adjadr + v.vadr;
      int cur = v.hint;
      float best = RG_LDG(vert + 3 * cur) * dl[0] + RG_LDG(vert + 3 * cur + 1) * dl[1] + RG_LDG(vert + 3 * cur + 2) * dl[2];
      for (int guard = 0; guard < v.vnum; guard++) {
        const int a0 = RG_LDG(adjadr + cur), a1 = RG_LDG(adjadr + cur + 1);
        int nxt = cur;
        for (int a = a0; a < a1; a++) {
          const int nb = RG_LDG(m.mesh_adj + a);
          const float dd = RG_LDG(vert + 3 * nb) * dl[0] + RG_LDG(vert + 3 * nb + 1) * dl[1] + RG_LDG(vert + 3 * nb + 2) * dl[2];
          if (dd > best) { best = dd; nxt = nb; }
        }
        if (nxt == cur) break;
        cur = nxt;
      }
      v.hint = cur;
      loc[0] = RG_LDG(vert + 3 * cur); loc[1] = RG_LDG(vert + 3 * cur + 1); loc[2] = RG_LDG(vert + 3 * cur + 2);
    } break;
    default: break;
  }
  rg_mulmat3(res, v.mat, loc);
  res[0] += v.pos[0] + dir[0] * v.halfmargin;
  res[1] += v.pos[1] + dir[1] * v.halfmargin;
  res[2] += v.pos[2] + dir[2] * v.halfmargin;
}

struct RgSup { float v[3], v1[3], v2[3]; };

RG_DEV void rg_mpr_support(const RgModel& m, RgGeomView& o1, RgGeomView& o2, const float* dir, RgSup& sp) {
  const float nd[3] = {-dir[0], -dir[1], -dir[2]};
  rg_support(m, o1, dir, sp.v1);
  rg_support(m, o2, nd, sp.v2);
  rg_sub3(sp.v, sp.v1, sp.v2);
}
RG_DEV int rg_mpr_zero(float x) { return fabsf(x) < RG_EPS; }
RG_DEV void rg_portal_dir(const RgSup* p, float* dir) {
  float a[3], b[3];
  rg_sub3(a, p[2].v, p[1].v); rg_sub3(b, p[3].v, p[1].v);
  rg_cross(dir, a, b);
  rg_normalize3(dir);
}
RG_DEV void rg_expand_portal(RgSup* p, const RgSup& v4) {
  float c4[3];
  rg_cross(c4, v4.v, p[0].v);
  if (rg_dot3(p[1].v, c4) > 0) { if (rg_dot3(p[2].v, c4) > 0) p[1] = v4; else p[3] = v4; }
  else { if (rg_dot3(p[3].v, c4) > 0) p[2] = v4; else p[1] = v4; }
}
RG_DEV int rg_reach_tol(const RgSup* p, const RgSup& v4, const float* dir, float tol) {
  const float dv4 = rg_dot3(v4.v, dir);
  const float mn = fminf(fminf(dv4 - rg_dot3(p[1].v, dir), dv4 - rg_dot3(p[2].v, dir)), dv4 - rg_dot3(p[3].v, dir));
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
RG_DEV void rg_find_pos(const RgSup* p, float* pos) {
  float dir[3], b[4], t[3];
  rg_portal_dir(p, dir);
  rg_cross(t, p[1].v, p[2].v); b[0] = rg_dot3(t, p[3].v);
  rg_cross(t, p[3].v, p[2].v); b[1] = rg_dot3(t, p[0].v);
  rg_cross(t, p[0].v, p[1].v); b[2] = rg_dot3(t, p[3].v);
  rg_cross(t, p[2].v, p[1].v); b[3] = rg_dot3(t, p[0].v);
  float sum = b[0] + b[1] + b[2] + b[3];
  if (sum <= 1e-30f) {
    b[0] = 0;
    rg_cross(t, p[2].v, p[3].v); b[1] = rg_dot3(t, dir);
    rg_cross(t, p[3].v, p[1].v); b[2] = rg_dot3(t, dir);
    rg_cross(t, p[1].v, p[2].v); b[3] = rg_dot3(t, dir);
    sum = b[1] + b[2] + b[3];
  }
  const float inv = 0.5f / sum;
  for (int i = 0; i < 3; i++)
    pos[i] = inv * (b[0] * (p[0].v1[i] + p[0].v2[i]) + b[1] * (p[1].v1[i] + p[1].v2[i]) + b[2] * (p[2].v1[i] + p[2].v2[i]) + b[3] * (p[3].v1[i] + p[3].v2[i]));
}

/* depth >= 0 with dir,pos when the inflated geoms intersect; -1 otherwise */
RG_DEV float rg_mpr(const RgModel& m, RgGeomView& o1, RgGeomView& o2, float tol, int maxiter, float* dir_out, float* pos) {
  RgSup p[4], v4;
  float dir[3], va[3], vb[3], dot;
  rg_copy3(p[0].v1, o1.pos); rg_copy3(p[0].v2, o2.pos);
  rg_sub3(p[0].v, p[0].v1, p[0].v2);
  if (rg_mpr_zero(p[0].v[0]) && rg_mpr_zero(p[0].v[1]) && rg_mpr_zero(p[0].v[2])) p[0].v[0] = 1e-5f;
  rg_scl3(dir, p[0].v, -1.0f); rg_normalize3(dir);
  rg_mpr_support(m, o1, o2, dir, p[1]);
  dot = rg_dot3(p[1].v, dir);
  if (dot < 0 || rg_mpr_zero(dot)) return -1.0f;
  rg_cross(dir, p[0].v, p[1].v);
  /* fp32: the parallel test must be scale-free (libccd compares the raw squared norm with DBL_EPSILON) */
  if (rg_dot3(dir, dir) <= 1e-12f * rg_dot3(p[0].v, p[0].v) * rg_dot3(p[1].v, p[1].v)) {
    const float n = sqrtf(rg_dot3(p[1].v, p[1].v));
    for (int i = 0; i < 3; i++) pos[i] = 0.5f * (p[1].v1[i] + p[1].v2[i]);
    if (n < RG_EPS) { dir_out[0] = dir_out[1] = dir_out[2] = 0; return 0.0f; }
    rg_scl3(dir_out, p[1].v, 1.0f / n);
    return n;
  }
  rg_normalize3(dir);
  rg_mpr_support(m, o1, o2, dir, p[2]);
  dot = rg_dot3(p[2].v, dir);
  if (dot < 0 || rg_mpr_zero(dot)) return -1.0f;
  rg_sub3(va, p[1].v, p[0].v); rg_sub3(vb, p[2].v, p[0].v);
  rg_cross(dir, va, vb); rg_normalize3(dir);
  if (rg_dot3(dir, p[0].v) > 0) { RgSup t = p[1]; p[1] = p[2]; p[2] = t; rg_scl3(dir, dir, -1.0f); }
  for (int guard = 0;; guard++) {
    if (guard > 100) return -1.0f;
    rg_mpr_support(m, o1, o2, dir, p[3]);
    dot = rg_dot3(p[3].v, dir);
    if (dot < 0 || rg_mpr_zero(dot)) return -1.0f;
    int cont = 0;
    rg_cross(va, p[1].v, p[3].v);
    dot = rg_dot3(va, p[0].v);
    if (dot < 0) { p[2] = p[3]; cont = 1; } /* triple products are ~1e-6: sign only */
    if (!cont) {
      rg_cross(va, p[3].v, p[2].v);
      dot = rg_dot3(va, p[0].v);
      if (dot < 0) { p[1] = p[3]; cont = 1; }
    }
    if (!cont) break;
    rg_sub3(va, p[1].v, p[0].v); rg_sub3(vb, p[2].v, p[0].v);
    rg_cross(dir, va, vb); rg_normalize3(dir);
  }
  for (int guard = 0;; guard++) {
    if (guard > 1000) return -1.0f;
    rg_portal_dir(p, dir);
    dot = rg_dot3(dir, p[1].v);
    if (dot > 0 || rg_mpr_zero(dot)) break;
    rg_mpr_support(m, o1, o2, dir, v4);
    dot = rg_dot3(v4.v, dir);
    if (!(dot > 0 || rg_mpr_zero(dot))) return -1.0f;
    if (rg_reach_tol(p, v4, dir, tol)) return -1.0f;
    rg_expand_portal(p, v4);
  }
  for (int it = 0;; it++) {
    rg_portal_dir(p, dir);
    rg_mpr_support(m, o1, o2, dir, v4);
    if (rg_reach_tol(p, v4, dir, tol) || it > maxiter) {
      float w[3];
      const float depth = sqrtf(rg_tri_closest(p[1].v, p[2].v, p[3].v, w));
      if (depth < RG_EPS) { dir_out[0] = dir_out[1] = dir_out[2] = 0; }
      else rg_scl3(dir_out, w, 1.0f / depth);
      rg_find_pos(p, pos);
      return depth;
    }
    rg_expand_portal(p, v4);
  }
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
RG_DEV int rg_narrow(const RgCtx& c, int g1, int g2, float margin, float* out) {
  const RgModel& m = c.m;
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
  RgGeomView o1, o2;
  rg_geom_view(c, g1, margin, o1);
  rg_geom_view(c, g2, margin, o2);
  float dir[3], pos[3];
  const float depth = rg_mpr(m, o1, o2, m.opt_mpr_tolerance[0], m.opt_mpr_iterations[0], dir, pos);
  if (depth < 0 || rg_dot3(dir, dir) < 0.5f) return 0;
  out[0] = margin - depth;
  rg_copy3(out + 1, pos);
  rg_copy3(out + 4, dir);
  return 1;
}

RG_DEV_NOINLINE void rg_collision(RgCtx& c) {
  RG_LANE_DECL
  const RgModel& m = c.m; const RgLayout& L = c.L; float* s = c.s;
  int* cand = (int*)(s + L.cand);
  int ncon = 0, ncand = 0, k0 = 0, warn = 0;
  const int cap = (m.nconmax > 0 && m.nconmax < RG_NCON) ? m.nconmax : RG_NCON;
  const int enabled = !(m.opt_disableflags[0] & (RG_DSBL_CONTACT | RG_DSBL_CONSTRAINT));
  while (enabled && (k0 < m.npair || ncand > 0)) {
    while (ncand < 32 && k0 < m.npair) {
      LANEVAR(int, pred); LANEVAR(int, pos);
      int tot;
      RG_PHASE_BEGIN
      const int k = k0 + lane;
      int pr = 0;
      if (k < m.npair) {
        const int g1 = RG_LDG(m.pair_geom1 + k), g2 = RG_LDG(m.pair_geom2 + k);
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
      if (LV(pred)) cand[ncand + LV(pos)] = k0 + lane;
      RG_PHASE_END
      ncand += tot;
      k0 += 32;
    }
    const int n = ncand < 32 ? ncand : 32;
    LANEVAR(int, cnt); LANEVAR(int, cpos); LANEVAR(int, keep);
    LANEARR(float, cb, 28);
    int tot;
    RG_PHASE_BEGIN
    int cn = 0;
    if (lane < n) {
      const int k = cand[lane];
      const int g1 = RG_LDG(m.pair_geom1 + k), g2 = RG_LDG(m.pair_geom2 + k);
      cn = rg_narrow(c, g1, g2, fmaxf(m.geom_margin[g1], m.geom_margin[g2]), &LA(cb, 0));
    }
    LV(cnt) = cn;
    LV(keep) = (lane + 32 < ncand) ? cand[lane + 32] : -1;
    RG_PHASE_END
    RG_WARP_SCAN(cnt, cpos, tot);
    RG_PHASE_BEGIN
    if (lane < n && LV(cnt) > 0) {
      const int k = cand[lane];
      const int g1 = RG_LDG(m.pair_geom1 + k), g2 = RG_LDG(m.pair_geom2 + k);
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
        for (int q = 0; q < 5; q++) r[24 + q] = solimp[q];
      }
    }
    RG_PHASE_END
    if (ncon + tot > cap) { warn |= RG_WARN_CONTACT_FULL; ncon = cap; } else ncon += tot;
    RG_PHASE_BEGIN
    if (LV(keep) >= 0) cand[lane] = LV(keep);
    RG_PHASE_END
    ncand = ncand > 32 ? ncand - 32 : 0;
  }
  RG_PHASE_BEGIN
  if (lane == 0) { RG_SI(c, RG_S_NCON) = ncon; RG_SI(c, RG_S_WARN) |= warn; }
  RG_PHASE_END
}
