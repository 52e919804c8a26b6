/* rgo_oracle.c -- CPU ORACLE, test infrastructure only.  See rgo_oracle.h for scope and the
 * "parity unpinned" statement.  Dense, scalar, fp64; written for clarity, not speed.
 *
 * Stage map (SURVEY.md 8(a'), reference call site robogym/mujoco/simulation_interface.py:184-185):
 *   S1 kinematics ........ rgo_kinematics      S9  impedance ......... make_rows
 *   S2 spatial inertia ... rgo_inertia         S10 passive + bias .... rgo_passive, rgo_bias
 *   S3 tendons ........... rgo_tendon          S11 actuation (PID, cascaded PI) ... rgo_actuation
 *   S4 transmission ...... rgo_transmission    S12 smooth accel ...... rgo_smooth
 *   S5/6 mass matrix ..... rgo_massmatrix      S13 Newton solver ..... rgo_solve
 *   S7 collision ......... rgo_collision       S15 Euler ............. rgo_euler
 *   S8 constraint rows ... make_rows           S14 sensors (touch, jointpos, force / torque) ... rgo_sensors
 */
#include "rgo_oracle.h"

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MINVAL 1e-15
#define MAXCON 256
#define MAXEFC 2048

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_ELLIPSOID = 4, GEOM_CYLINDER = 5, GEOM_BOX = 6, GEOM_MESH = 7 };
enum { WRAP_JOINT = 1, WRAP_PULLEY = 2, WRAP_SITE = 3, WRAP_SPHERE = 4, WRAP_CYLINDER = 5 };
enum { TRN_JOINT = 0, TRN_TENDON = 3 };
enum { GAIN_FIXED = 0, GAIN_USER = 2 };
enum { BIAS_NONE = 0, BIAS_AFFINE = 1, BIAS_USER = 3 };
enum { DSBL_CONSTRAINT = 1, DSBL_EQUALITY = 2, DSBL_FRICTIONLOSS = 4, DSBL_LIMIT = 8, DSBL_CONTACT = 16,
       DSBL_PASSIVE = 32, DSBL_GRAVITY = 64, DSBL_CLAMPCTRL = 128, DSBL_WARMSTART = 256,
       DSBL_ACTUATION = 1024, DSBL_REFSAFE = 2048 };
enum { EQ_CONNECT = 0, EQ_WELD = 1, EQ_JOINT = 2 };
enum { ROW_EQUALITY = 0, ROW_FRICTION = 1, ROW_LIMIT = 2, ROW_CONTACT = 3, ROW_CONTACT_ELL = 4 /* first row of an elliptic-cone contact: dim rows follow each other */, ROW_CONTACT_ELLF = 5 /* its friction rows */ };

/* ------------------------------------------------------------------ model */
struct rgo_model {
#define RG_DIM(n) int n;
#define RG_I(n, c) int* n;
#define RG_F(n, c) double* n;
#include "../include/rg_model_fields.h"
#undef RG_DIM
#undef RG_I
#undef RG_F
  void* storage;
};

rgo_model* rgo_model_load(const void* blob, size_t len) {
  const char* p = (const char*)blob;
  if (len < 12 || memcmp(p, "RGMODEL1", 8)) return NULL;
  rgo_model* m = (rgo_model*)calloc(1, sizeof(rgo_model));
  m->storage = malloc(len);
  memcpy(m->storage, blob, len);
  char* base = (char*)m->storage;
  int ndim;
  memcpy(&ndim, base + 8, 4);
  const int* dims = (const int*)(base + 12);
  int k = 0;
#define RG_DIM(n) m->n = dims[k++];
#define RG_I(n, c)
#define RG_F(n, c)
#include "../include/rg_model_fields.h"
#undef RG_DIM
#undef RG_I
#undef RG_F
  if (k != ndim) { free(m->storage); free(m); return NULL; }
  size_t off = 12 + 4 * (size_t)ndim;
  /* bring the dims into scope for the count expressions */
#define RG_DIM(n) const int n = m->n; (void)n;
#define RG_I(n, c)
#define RG_F(n, c)
#include "../include/rg_model_fields.h"
#undef RG_DIM
#undef RG_I
#undef RG_F
#define RG_DIM(n)
#define RG_I(n, c) off = (off + 7) & ~(size_t)7; m->n = (int*)(base + off); off += 4 * (size_t)(c);
#define RG_F(n, c) off = (off + 7) & ~(size_t)7; m->n = (double*)(base + off); off += 8 * (size_t)(c);
#include "../include/rg_model_fields.h"
#undef RG_DIM
#undef RG_I
#undef RG_F
  if (off > len) { free(m->storage); free(m); return NULL; }
  return m;
}

void rgo_model_free(rgo_model* m) { if (m) { free(m->storage); free(m); } }

int rgo_model_dim(const rgo_model* m, const char* name) {
#define RG_DIM(n) if (!strcmp(name, #n)) return m->n;
#define RG_I(n, c)
#define RG_F(n, c)
#include "../include/rg_model_fields.h"
#undef RG_DIM
#undef RG_I
#undef RG_F
  return -1;
}

void* rgo_model_field(rgo_model* m, const char* name, int* count, int* is_int) {
#define RG_DIM(n) const int n = m->n; (void)n;
#define RG_I(n, c)
#define RG_F(n, c)
#include "../include/rg_model_fields.h"
#undef RG_DIM
#undef RG_I
#undef RG_F
#define RG_DIM(n)
#define RG_I(n, c) if (!strcmp(name, #n)) { *count = (c); *is_int = 1; return m->n; }
#define RG_F(n, c) if (!strcmp(name, #n)) { *count = (c); *is_int = 0; return m->n; }
#include "../include/rg_model_fields.h"
#undef RG_DIM
#undef RG_I
#undef RG_F
  return NULL;
}

/* ------------------------------------------------------------------ data */
#define DATA_FIELDS(F)                                                                                         \
  F(qpos, nq) F(qvel, nv) F(ctrl, nu) F(userdata, nuserdata) F(qacc_warmstart, nv) F(xfrc_applied, nbody * 6) \
  F(time, 1) F(xpos, nbody * 3) F(xquat, nbody * 4) F(xmat, nbody * 9) F(xipos, nbody * 3) F(ximat, nbody * 9) \
  F(geom_xpos, ngeom * 3) F(geom_xmat, ngeom * 9) F(site_xpos, nsite * 3) F(site_xmat, nsite * 9)             \
  F(dof_S, nv * 6) F(dof_Sdot, nv * 6) F(cvel, nbody * 6) F(cacc, nbody * 6) F(ten_length, ntendon)            \
  F(ten_J, ntendon * nv) F(ten_velocity, ntendon) F(actuator_length, nu) F(actuator_velocity, nu)              \
  F(actuator_moment, nu * nv) F(actuator_force, nu) F(M, nv * nv) F(Mchol, nv * nv) F(qfrc_bias, nv)           \
  F(qfrc_passive, nv) F(qfrc_actuator, nv) F(qfrc_applied_total, nv) F(qfrc_smooth, nv) F(qacc_smooth, nv)     \
  F(qacc, nv) F(qfrc_constraint, nv) F(contact, MAXCON * RGO_CON_STRIDE) F(efc_J, MAXEFC * nv)                \
  F(efc_aref, MAXEFC) F(efc_D, MAXEFC) F(efc_R, MAXEFC) F(efc_force, MAXEFC) F(efc_pos, MAXEFC)                \
  F(efc_margin, MAXEFC) F(efc_floss, MAXEFC) F(efc_vel, MAXEFC) F(efc_diagApprox, MAXEFC)                     \
  F(efc_solref, MAXEFC * 2) F(efc_solimp, MAXEFC * 5) F(efc_jar, MAXEFC) F(sensordata, nsensordata + 1)        \
  F(body_I10, nbody * 10) F(wrap_xpos, nwrap * 6 + 6) F(solver_stat, 8) F(contact_solimp, MAXCON * 5)         \
  F(mocap_pos, nmocap * 3) F(mocap_quat, nmocap * 4)

struct rgo_data {
#define F(n, c) double* n;
  DATA_FIELDS(F)
#undef F
  int ncon, nefc, solver_niter, warning;
  int* efc_type; /* MAXEFC */
  int* efc_id;   /* MAXEFC */
  /* sizes for field lookup */
  const rgo_model* m;
};

rgo_data* rgo_data_new(const rgo_model* m) {
  rgo_data* d = (rgo_data*)calloc(1, sizeof(rgo_data));
  d->m = m;
#define RG_DIM(n) const int n = m->n; (void)n;
#define RG_I(n, c)
#define RG_F(n, c)
#include "../include/rg_model_fields.h"
#undef RG_DIM
#undef RG_I
#undef RG_F
#define F(n, c) d->n = (double*)calloc((size_t)(c) + 1, sizeof(double));
  DATA_FIELDS(F)
#undef F
  d->efc_type = (int*)calloc(MAXEFC, sizeof(int));
  d->efc_id = (int*)calloc(MAXEFC, sizeof(int));
  rgo_reset(m, d);
  return d;
}

void rgo_data_free(rgo_data* d) {
  if (!d) return;
#define F(n, c) free(d->n);
  DATA_FIELDS(F)
#undef F
  free(d->efc_type);
  free(d->efc_id);
  free(d);
}

void* rgo_data_field(rgo_data* d, const char* name, int* count, int* is_int) {
  const rgo_model* m = d->m;
#define RG_DIM(n) const int n = m->n; (void)n;
#define RG_I(n, c)
#define RG_F(n, c)
#include "../include/rg_model_fields.h"
#undef RG_DIM
#undef RG_I
#undef RG_F
  *is_int = 0;
#define F(n, c) if (!strcmp(name, #n)) { *count = (c); return d->n; }
  DATA_FIELDS(F)
#undef F
  *is_int = 1;
  *count = 1;
  if (!strcmp(name, "ncon")) return &d->ncon;
  if (!strcmp(name, "nefc")) return &d->nefc;
  if (!strcmp(name, "solver_niter")) return &d->solver_niter;
  if (!strcmp(name, "warning")) return &d->warning;
  if (!strcmp(name, "efc_type")) { *count = MAXEFC; return d->efc_type; }
  if (!strcmp(name, "efc_id")) { *count = MAXEFC; return d->efc_id; }
  return NULL;
}

void rgo_reset(const rgo_model* m, rgo_data* d) {
  memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
  memset(d->qvel, 0, sizeof(double) * m->nv);
  memset(d->ctrl, 0, sizeof(double) * m->nu);
  memset(d->userdata, 0, sizeof(double) * m->nuserdata);
  memset(d->qacc_warmstart, 0, sizeof(double) * m->nv);
  memset(d->xfrc_applied, 0, sizeof(double) * m->nbody * 6);
  memset(d->qacc, 0, sizeof(double) * m->nv);
  d->time[0] = 0;
  d->ncon = d->nefc = 0;
  d->warning = 0;
  /* mocap bodies start at their model pose (mj_resetData; robogym then moves them through data.mocap_pos / mocap_quat,
     robogym/robot/control/tcp/mocap_solver.py:41-46 via gym's mocap_set_action / reset_mocap2body_xpos) */
  for (int b = 0; b < m->nbody; b++)
    if (m->body_mocapid[b] >= 0) {
      memcpy(d->mocap_pos + 3 * m->body_mocapid[b], m->body_pos + 3 * b, 3 * sizeof(double));
      memcpy(d->mocap_quat + 4 * m->body_mocapid[b], m->body_quat + 4 * b, 4 * sizeof(double));
    }
}

/* ------------------------------------------------------------------ small math */
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void sub3(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void add3(double* r, const double* a, const double* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static inline void copy3(double* r, const double* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline void scl3(double* r, const double* a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
static inline void addscl3(double* r, const double* a, double s) { r[0] += a[0] * s; r[1] += a[1] * s; r[2] += a[2] * s; }
static inline double norm3(const double* a) { return sqrt(dot3(a, a)); }
static inline double normalize3(double* a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = a[2] = 0; return 0; }
  a[0] /= n; a[1] /= n; a[2] /= n;
  return n;
}
static void quat_mul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_norm(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void quat2mat(double* m, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static inline void mulmat3(double* r, const double* m, const double* v) { /* r = m v */
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
         z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmatT3(double* r, const double* m, const double* v) { /* r = m^T v */
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
         z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void rotvec(double* r, const double* q, const double* v) {
  double m[9];
  quat2mat(m, q);
  mulmat3(r, m, v);
}
static void axisangle2quat(double* q, const double* axis, double angle) {
  double s = sin(angle * 0.5);
  q[0] = cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
static void matmul33(double* r, const double* a, const double* b) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(r, t, sizeof t);
}
static inline int dof_in_body(const rgo_model* m, int body, int dof) {
  return (((const uint32_t*)m->body_dofmask)[body * m->nmaskw + (dof >> 5)] >> (dof & 31)) & 1u;
}

/* ------------------------------------------------------------------ S1 kinematics */
static void rgo_kinematics(const rgo_model* m, rgo_data* d) {
  double* xpos = d->xpos; double* xquat = d->xquat;
  xpos[0] = xpos[1] = xpos[2] = 0; xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
  quat2mat(d->xmat, xquat);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b];
    double pos[3], quat[4], t[3];
    rotvec(t, xquat + 4 * p, m->body_pos + 3 * b);
    add3(pos, xpos + 3 * p, t);
    quat_mul(quat, xquat + 4 * p, m->body_quat + 4 * b);
    if (m->body_mocapid[b] >= 0) {   /* mocap body (child of the world, no joints): pose comes from the data */
      copy3(pos, d->mocap_pos + 3 * m->body_mocapid[b]);
      memcpy(quat, d->mocap_quat + 4 * m->body_mocapid[b], 4 * sizeof(double));
    }
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k, qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
      int type = m->jnt_type[j];
      if (type == JNT_FREE) {
        copy3(pos, d->qpos + qa);
        memcpy(quat, d->qpos + qa + 3, 4 * sizeof(double));
        quat_norm(quat);
        /* motion axes: 3 world translations, then 3 body-frame rotations about the body origin */
        double R[9];
        quat2mat(R, quat);
        for (int a = 0; a < 3; a++) {
          double* S = d->dof_S + 6 * (da + a);
          S[0] = S[1] = S[2] = 0; S[3] = S[4] = S[5] = 0; S[3 + a] = 1;
          double* Sr = d->dof_S + 6 * (da + 3 + a);
          double ax[3] = {R[a], R[3 + a], R[6 + a]};
          copy3(Sr, ax);
          cross3(Sr + 3, pos, ax);
        }
        continue;
      }
      double anchor[3], axis[3];
      rotvec(t, quat, m->jnt_pos + 3 * j);
      add3(anchor, pos, t);
      rotvec(axis, quat, m->jnt_axis + 3 * j);
      if (type == JNT_SLIDE) {
        addscl3(pos, axis, d->qpos[qa] - m->qpos0[qa]);
        double* S = d->dof_S + 6 * da;
        S[0] = S[1] = S[2] = 0; copy3(S + 3, axis);
      } else if (type == JNT_HINGE) {
        double qloc[4], qn[4];
        axisangle2quat(qloc, m->jnt_axis + 3 * j, d->qpos[qa] - m->qpos0[qa]);
        quat_mul(qn, quat, qloc);
        memcpy(quat, qn, sizeof qn);
        rotvec(t, quat, m->jnt_pos + 3 * j);
        sub3(pos, anchor, t);
        double* S = d->dof_S + 6 * da;
        copy3(S, axis); cross3(S + 3, anchor, axis);
      } else { /* ball */
        double qloc[4], qn[4];
        memcpy(qloc, d->qpos + qa, sizeof qloc);
        quat_norm(qloc);
        quat_mul(qn, quat, qloc);
        memcpy(quat, qn, sizeof qn);
        rotvec(t, quat, m->jnt_pos + 3 * j);
        sub3(pos, anchor, t);
        double R[9];
        quat2mat(R, quat);
        for (int a = 0; a < 3; a++) {
          double* S = d->dof_S + 6 * (da + a);
          double ax[3] = {R[a], R[3 + a], R[6 + a]};
          copy3(S, ax); cross3(S + 3, anchor, ax);
        }
      }
    }
    quat_norm(quat);
    copy3(xpos + 3 * b, pos);
    memcpy(xquat + 4 * b, quat, sizeof quat);
    quat2mat(d->xmat + 9 * b, quat);
    rotvec(t, quat, m->body_ipos + 3 * b);
    add3(d->xipos + 3 * b, pos, t);
    double iq[4], R[9];
    quat_mul(iq, quat, m->body_iquat + 4 * b);
    quat2mat(R, iq);
    memcpy(d->ximat + 9 * b, R, sizeof R);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    double t[3], q[4];
    rotvec(t, xquat + 4 * b, m->geom_pos + 3 * g);
    add3(d->geom_xpos + 3 * g, xpos + 3 * b, t);
    quat_mul(q, xquat + 4 * b, m->geom_quat + 4 * g);
    quat_norm(q);
    quat2mat(d->geom_xmat + 9 * g, q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s];
    double t[3], q[4];
    rotvec(t, xquat + 4 * b, m->site_pos + 3 * s);
    add3(d->site_xpos + 3 * s, xpos + 3 * b, t);
    quat_mul(q, xquat + 4 * b, m->site_quat + 4 * s);
    quat_norm(q);
    quat2mat(d->site_xmat + 9 * s, q);
  }
}

/* translational/rotational Jacobian column of dof `dof` for a point fixed to a body it moves */
static inline void jac_col(const rgo_data* d, int dof, const double* point, double* jp, double* jr) {
  const double* S = d->dof_S + 6 * dof;
  copy3(jr, S);
  double t[3];
  cross3(t, S, point); /* v(point) = v_O + w x point */
  add3(jp, S + 3, t);
}

/* ------------------------------------------------------------------ S2/S5 inertia + mass matrix */
/* spatial inertia about the world origin stored as (mass, com[3], Ic[6] = xx yy zz xy xz yz) */
static void rgo_inertia(const rgo_model* m, rgo_data* d) {
  for (int b = 0; b < m->nbody; b++) {
    double* I = d->body_I10 + 10 * b;
    const double* R = d->ximat + 9 * b;
    const double* di = m->body_inertia + 3 * b;
    I[0] = m->body_mass[b];
    copy3(I + 1, d->xipos + 3 * b);
    double Ic[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Ic[3 * i + j] = R[3 * i] * di[0] * R[3 * j] + R[3 * i + 1] * di[1] * R[3 * j + 1] + R[3 * i + 2] * di[2] * R[3 * j + 2];
    I[4] = Ic[0]; I[5] = Ic[4]; I[6] = Ic[8]; I[7] = Ic[1]; I[8] = Ic[2]; I[9] = Ic[5];
  }
}
/* F = I * V  (V = [w; vO] motion, F = [nO; f] force) */
static void inertia_mul(double* F, const double* I, const double* V) {
  double mass = I[0];
  const double* c = I + 1;
  double t[3], f[3], n[3];
  cross3(t, V, c);          /* w x c */
  add3(f, V + 3, t);        /* vO + w x c */
  scl3(f, f, mass);
  n[0] = I[4] * V[0] + I[7] * V[1] + I[8] * V[2];
  n[1] = I[7] * V[0] + I[5] * V[1] + I[9] * V[2];
  n[2] = I[8] * V[0] + I[9] * V[1] + I[6] * V[2];
  cross3(t, c, f);
  add3(F, n, t);
  copy3(F + 3, f);
}
static inline double dot6(const double* a, const double* b) { return dot3(a, b) + dot3(a + 3, b + 3); }

static void rgo_massmatrix(const rgo_model* m, rgo_data* d) {
  int nv = m->nv;
  memset(d->M, 0, sizeof(double) * nv * nv);
  for (int b = 1; b < m->nbody; b++) {
    const double* I = d->body_I10 + 10 * b;
    if (I[0] == 0 && I[4] == 0 && I[5] == 0 && I[6] == 0) continue;
    for (int j = 0; j < nv; j++) {
      if (!dof_in_body(m, b, j)) continue;
      double F[6];
      inertia_mul(F, I, d->dof_S + 6 * j);
      for (int i = 0; i < nv; i++) {
        if (!dof_in_body(m, b, i)) continue;
        d->M[i * nv + j] += dot6(d->dof_S + 6 * i, F);
      }
    }
  }
  for (int i = 0; i < nv; i++) d->M[i * nv + i] += m->dof_armature[i];
}

/* dense Cholesky A = L L^T (lower, in place into L); returns 0 on success */
static int chol_factor(double* L, const double* A, int n) {
  memcpy(L, A, sizeof(double) * n * n);
  for (int j = 0; j < n; j++) {
    double s = L[j * n + j];
    for (int k = 0; k < j; k++) s -= L[j * n + k] * L[j * n + k];
    if (s < MINVAL) s = MINVAL;
    double lj = sqrt(s);
    L[j * n + j] = lj;
    for (int i = j + 1; i < n; i++) {
      double t = L[i * n + j];
      for (int k = 0; k < j; k++) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / lj;
    }
  }
  return 0;
}
static void chol_solve(const double* L, double* x, const double* b, int n) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

/* ------------------------------------------------------------------ velocities + bias (RNE) */
static void cross_motion(double* r, const double* V, const double* S) {
  double a[3], b[3], c[3];
  cross3(a, V, S);
  cross3(b, V, S + 3);
  cross3(c, V + 3, S);
  copy3(r, a);
  add3(r + 3, b, c);
}
static void cross_force(double* r, const double* V, const double* F) {
  double a[3], b[3], c[3];
  cross3(a, V, F);
  cross3(b, V + 3, F + 3);
  cross3(c, V, F + 3);
  add3(r, a, b);
  copy3(r + 3, c);
}

static void rgo_velocity(const rgo_model* m, rgo_data* d) {
  memset(d->cvel, 0, sizeof(double) * 6);
  for (int b = 1; b < m->nbody; b++) {
    double V[6];
    memcpy(V, d->cvel + 6 * m->body_parentid[b], sizeof V);
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k, da = m->jnt_dofadr[j];
      int nd = m->jnt_type[j] == JNT_FREE ? 6 : (m->jnt_type[j] == JNT_BALL ? 3 : 1);
      /* axis derivative: the velocity accumulated before the joint's rotational dofs -- the three axes of a ball joint
         turn together, so its own rotation does not enter; a free joint's translations come first and do count
         (mj_comVel: cdof_dot = 0 for the translations, cvel += them, then the ball rule for the rotations) */
      int a0 = 0;
      if (m->jnt_type[j] == JNT_FREE) {
        for (int a = 0; a < 3; a++) cross_motion(d->dof_Sdot + 6 * (da + a), V, d->dof_S + 6 * (da + a));
        for (int a = 0; a < 3; a++)
          for (int i = 0; i < 6; i++) V[i] += d->dof_S[6 * (da + a) + i] * d->qvel[da + a];
        a0 = 3;
      }
      for (int a = a0; a < nd; a++) cross_motion(d->dof_Sdot + 6 * (da + a), V, d->dof_S + 6 * (da + a));
      for (int a = a0; a < nd; a++)
        for (int i = 0; i < 6; i++) V[i] += d->dof_S[6 * (da + a) + i] * d->qvel[da + a];
    }
    memcpy(d->cvel + 6 * b, V, sizeof V);
  }
}

/* qfrc_bias = C(q,qdot) qdot - gravity term */
static void rgo_bias(const rgo_model* m, rgo_data* d) {
  int nv = m->nv;
  double* A = d->cacc;
  memset(A, 0, sizeof(double) * 6);
  if (!(m->opt_disableflags[0] & DSBL_GRAVITY)) { A[3] = -m->opt_gravity[0]; A[4] = -m->opt_gravity[1]; A[5] = -m->opt_gravity[2]; }
  memset(d->qfrc_bias, 0, sizeof(double) * nv);
  for (int b = 1; b < m->nbody; b++) {
    double* Ab = A + 6 * b;
    memcpy(Ab, A + 6 * m->body_parentid[b], 6 * sizeof(double));
    for (int k = 0; k < m->body_dofnum[b]; k++) {
      int dof = m->body_dofadr[b] + k;
      for (int i = 0; i < 6; i++) Ab[i] += d->dof_Sdot[6 * dof + i] * d->qvel[dof];
    }
    const double* I = d->body_I10 + 10 * b;
    double F[6], H[6], G[6];
    inertia_mul(F, I, Ab);
    inertia_mul(H, I, d->cvel + 6 * b);
    cross_force(G, d->cvel + 6 * b, H);
    for (int i = 0; i < 6; i++) F[i] += G[i];
    for (int dof = 0; dof < nv; dof++)
      if (dof_in_body(m, b, dof)) d->qfrc_bias[dof] += dot6(d->dof_S + 6 * dof, F);
  }
}

/* ------------------------------------------------------------------ S3 tendons */
static int seg_intersect(const double* p1, const double* p2, const double* p3, const double* p4) {
  /* do 2D segments p1-p2 and p3-p4 cross? */
  double det = (p4[1] - p3[1]) * (p2[0] - p1[0]) - (p4[0] - p3[0]) * (p2[1] - p1[1]);
  if (fabs(det) < MINVAL) return 0;
  double a = ((p4[0] - p3[0]) * (p1[1] - p3[1]) - (p4[1] - p3[1]) * (p1[0] - p3[0])) / det;
  double b = ((p2[0] - p1[0]) * (p1[1] - p3[1]) - (p2[1] - p1[1]) * (p1[0] - p3[0])) / det;
  return (a >= 0 && a <= 1 && b >= 0 && b <= 1);
}
/* 2D: tangent points on circle radius rad (origin) for path d0 -> circle -> d1; returns arc length or -1 */
static double wrap_circle(double* pnt, const double* d0, const double* d1, const double* sd, double rad) {
  double sqlen0 = d0[0] * d0[0] + d0[1] * d0[1], sqlen1 = d1[0] * d1[0] + d1[1] * d1[1], sqrad = rad * rad;
  double dif[2] = {d1[0] - d0[0], d1[1] - d0[1]};
  double dd = dif[0] * dif[0] + dif[1] * dif[1];
  if (sqlen0 < sqrad || sqlen1 < sqrad || rad < MINVAL) return -1;
  if (dd < MINVAL) return -1;
  double a = -(dif[0] * d0[0] + dif[1] * d0[1]) / dd;
  if (a < 0) a = 0; else if (a > 1) a = 1;
  double tmp[2] = {a * dif[0] + d0[0], a * dif[1] + d0[1]};
  if (tmp[0] * tmp[0] + tmp[1] * tmp[1] > sqrad && (!sd || tmp[0] * sd[0] + tmp[1] * sd[1] >= 0)) return -1;
  double sol[2][4], good[2];
  double sqrt0 = sqrt(sqlen0 - sqrad), sqrt1 = sqrt(sqlen1 - sqrad);
  for (int i = 0; i < 2; i++) {
    double sgn = i == 0 ? 1 : -1;
    sol[i][0] = (d0[0] * sqrad + sgn * rad * d0[1] * sqrt0) / sqlen0;
    sol[i][1] = (d0[1] * sqrad - sgn * rad * d0[0] * sqrt0) / sqlen0;
    sol[i][2] = (d1[0] * sqrad - sgn * rad * d1[1] * sqrt1) / sqlen1;
    sol[i][3] = (d1[1] * sqrad + sgn * rad * d1[0] * sqrt1) / sqlen1;
    if (sd) {
      double t[2] = {sol[i][0] + sol[i][2], sol[i][1] + sol[i][3]};
      double n = sqrt(t[0] * t[0] + t[1] * t[1]);
      if (n < MINVAL) n = MINVAL;
      good[i] = (t[0] * sd[0] + t[1] * sd[1]) / n;
    } else {
      double t[2] = {sol[i][0] - sol[i][2], sol[i][1] - sol[i][3]};
      good[i] = -(t[0] * t[0] + t[1] * t[1]);
    }
    if (seg_intersect(d0, sol[i], d1, sol[i] + 2)) good[i] = -10000;
  }
  int i = good[0] > good[1] ? 0 : 1;
  memcpy(pnt, sol[i], 4 * sizeof(double));
  if (seg_intersect(d0, pnt, d1, pnt + 2)) return -1;
  double c = (pnt[0] * pnt[2] + pnt[1] * pnt[3]) / sqrad;
  if (c > 1) c = 1; else if (c < -1) c = -1;
  return rad * acos(c);
}
/* wrap x0 -> geom -> x1; wpnt gets the two tangent points (world); returns curve length or -1 */
static double wrap_geom(double* wpnt, const double* x0, const double* x1, const double* gpos, const double* gmat,
                        double rad, int type, const double* side) {
  double t[3], p0[3], p1[3], sdl[3];
  sub3(t, x0, gpos); mulmatT3(p0, gmat, t);
  sub3(t, x1, gpos); mulmatT3(p1, gmat, t);
  if (norm3(p0) < MINVAL || norm3(p1) < MINVAL) return -1;
  double axis0[3], axis1[3];
  if (type == WRAP_SPHERE) {
    copy3(axis0, p0); normalize3(axis0);
    double nrm[3];
    cross3(nrm, p0, p1);
    if (normalize3(nrm) < MINVAL) { /* collinear: any perpendicular */
      double e[3] = {1, 0, 0};
      if (fabs(axis0[0]) > 0.9) { e[0] = 0; e[1] = 1; }
      cross3(nrm, axis0, e); normalize3(nrm);
    }
    cross3(axis1, nrm, axis0); normalize3(axis1);
  } else {
    axis0[0] = 1; axis0[1] = axis0[2] = 0;
    axis1[1] = 1; axis1[0] = axis1[2] = 0;
  }
  double d0[2] = {dot3(p0, axis0), dot3(p0, axis1)}, d1[2] = {dot3(p1, axis0), dot3(p1, axis1)};
  double sd2[2];
  const double* sdp = NULL;
  if (side) {
    sub3(t, side, gpos); mulmatT3(sdl, gmat, t);
    sd2[0] = dot3(sdl, axis0); sd2[1] = dot3(sdl, axis1);
    double n = sqrt(sd2[0] * sd2[0] + sd2[1] * sd2[1]);
    if (n < rad) { fprintf(stderr, "rgo: inside wrap (sidesite inside wrap geom) is not implemented\n"); return -1; }
    sd2[0] /= n; sd2[1] /= n;
    sdp = sd2;
  }
  double pnt[4];
  double wlen = wrap_circle(pnt, d0, d1, sdp, rad);
  if (wlen < 0) return -1;
  double r0[3], r1[3];
  for (int i = 0; i < 3; i++) { r0[i] = axis0[i] * pnt[0] + axis1[i] * pnt[1]; r1[i] = axis0[i] * pnt[2] + axis1[i] * pnt[3]; }
  if (type == WRAP_CYLINDER) {
    double L0 = sqrt((d0[0] - pnt[0]) * (d0[0] - pnt[0]) + (d0[1] - pnt[1]) * (d0[1] - pnt[1]));
    double L1 = sqrt((d1[0] - pnt[2]) * (d1[0] - pnt[2]) + (d1[1] - pnt[3]) * (d1[1] - pnt[3]));
    double tot = L0 + wlen + L1;
    r0[2] = p0[2] + (p1[2] - p0[2]) * L0 / tot;
    r1[2] = p0[2] + (p1[2] - p0[2]) * (L0 + wlen) / tot;
    double h = r1[2] - r0[2];
    wlen = sqrt(wlen * wlen + h * h);
  }
  mulmat3(t, gmat, r0); add3(wpnt, t, gpos);
  mulmat3(t, gmat, r1); add3(wpnt + 3, t, gpos);
  return wlen;
}

/* add (jacp(B at pb) - jacp(A at pa))^T dir * scale to the dense row J */
static void tendon_seg_jac(const rgo_model* m, const rgo_data* d, double* J, int ba, const double* pa, int bb,
                           const double* pb, const double* dir, double scale) {
  if (ba == bb) return;
  for (int dof = 0; dof < m->nv; dof++) {
    int ina = dof_in_body(m, ba, dof), inb = dof_in_body(m, bb, dof);
    if (ina == inb) continue; /* common ancestors cancel */
    double jp[3], jr[3];
    if (inb) { jac_col(d, dof, pb, jp, jr); J[dof] += scale * dot3(jp, dir); }
    else { jac_col(d, dof, pa, jp, jr); J[dof] -= scale * dot3(jp, dir); }
  }
}

static void rgo_tendon(const rgo_model* m, rgo_data* d) {
  int nv = m->nv;
  memset(d->ten_J, 0, sizeof(double) * m->ntendon * nv);
  for (int t = 0; t < m->ntendon; t++) {
    double L = 0, divisor = 1;
    double* J = d->ten_J + t * nv;
    int adr = m->tendon_adr[t], num = m->tendon_num[t];
    if (m->wrap_type[adr] == WRAP_JOINT) {
      for (int w = adr; w < adr + num; w++) {
        int j = m->wrap_objid[w];
        L += m->wrap_prm[w] * d->qpos[m->jnt_qposadr[j]];
        J[m->jnt_dofadr[j]] += m->wrap_prm[w];
      }
      d->ten_length[t] = L;
      continue;
    }
    int w = adr;
    while (w < adr + num - 1) {
      int t0 = m->wrap_type[w], t1 = m->wrap_type[w + 1];
      if (t0 == WRAP_PULLEY) { divisor = m->wrap_prm[w]; w++; continue; }
      if (t1 == WRAP_PULLEY) { w++; continue; }
      /* t0 must be a site here */
      int s0 = m->wrap_objid[w];
      const double* x0 = d->site_xpos + 3 * s0;
      int b0 = m->site_bodyid[s0];
      if (t1 == WRAP_SITE) {
        int s1 = m->wrap_objid[w + 1];
        const double* x1 = d->site_xpos + 3 * s1;
        double dir[3];
        sub3(dir, x1, x0);
        double len = normalize3(dir);
        L += len / divisor;
        tendon_seg_jac(m, d, J, b0, x0, m->site_bodyid[s1], x1, dir, 1.0 / divisor);
        w += 1;
      } else { /* site, geom, site */
        int g = m->wrap_objid[w + 1], s1 = m->wrap_objid[w + 2];
        int sid = (int)m->wrap_prm[w + 1];
        const double* x1 = d->site_xpos + 3 * s1;
        int b1 = m->site_bodyid[s1], bg = m->geom_bodyid[g];
        double wp[6];
        double wlen = wrap_geom(wp, x0, x1, d->geom_xpos + 3 * g, d->geom_xmat + 9 * g, m->geom_size[3 * g], t1,
                                sid >= 0 ? d->site_xpos + 3 * sid : NULL);
        if (wlen < 0) {
          double dir[3];
          sub3(dir, x1, x0);
          double len = normalize3(dir);
          L += len / divisor;
          tendon_seg_jac(m, d, J, b0, x0, b1, x1, dir, 1.0 / divisor);
        } else {
          double dir[3];
          sub3(dir, wp, x0);
          double len = normalize3(dir);
          L += len / divisor;
          tendon_seg_jac(m, d, J, b0, x0, bg, wp, dir, 1.0 / divisor);
          L += wlen / divisor;
          sub3(dir, x1, wp + 3);
          len = normalize3(dir);
          L += len / divisor;
          tendon_seg_jac(m, d, J, bg, wp + 3, b1, x1, dir, 1.0 / divisor);
        }
        w += 2;
      }
    }
    d->ten_length[t] = L;
  }
}

void rgo_tendon_eval(const rgo_model* m, rgo_data* d, const double* qpos, double* length, double* J) {
  double* save = (double*)malloc(sizeof(double) * m->nq);
  memcpy(save, d->qpos, sizeof(double) * m->nq);
  memcpy(d->qpos, qpos, sizeof(double) * m->nq);
  rgo_kinematics(m, d);
  rgo_tendon(m, d);
  memcpy(length, d->ten_length, sizeof(double) * m->ntendon);
  memcpy(J, d->ten_J, sizeof(double) * m->ntendon * m->nv);
  memcpy(d->qpos, save, sizeof(double) * m->nq);
  free(save);
}

/* ------------------------------------------------------------------ S4 transmission */
static void rgo_transmission(const rgo_model* m, rgo_data* d) {
  int nv = m->nv;
  memset(d->actuator_moment, 0, sizeof(double) * m->nu * nv);
  for (int i = 0; i < m->nu; i++) {
    double gear = m->actuator_gear[6 * i];
    int id = m->actuator_trnid[i];
    if (m->actuator_trntype[i] == TRN_JOINT) {
      d->actuator_length[i] = gear * d->qpos[m->jnt_qposadr[id]];
      d->actuator_moment[i * nv + m->jnt_dofadr[id]] = gear;
    } else {
      d->actuator_length[i] = gear * d->ten_length[id];
      for (int k = 0; k < nv; k++) d->actuator_moment[i * nv + k] = gear * d->ten_J[id * nv + k];
    }
  }
}

/* ------------------------------------------------------------------ S10 passive */
static void rgo_passive(const rgo_model* m, rgo_data* d) {
  int nv = m->nv;
  memset(d->qfrc_passive, 0, sizeof(double) * nv);
  for (int t = 0; t < m->ntendon; t++) {
    double v = 0;
    for (int k = 0; k < nv; k++) v += d->ten_J[t * nv + k] * d->qvel[k];
    d->ten_velocity[t] = v;
  }
  if (m->opt_disableflags[0] & DSBL_PASSIVE) return;
  for (int j = 0; j < m->njnt; j++) {
    double k = m->jnt_stiffness[j];
    if (k == 0) continue;
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_SLIDE || m->jnt_type[j] == JNT_HINGE)
      d->qfrc_passive[da] -= k * (d->qpos[qa] - m->qpos_spring[qa]);
    /* ball/free springs are not used by the robogym configs */
  }
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] -= m->dof_damping[i] * d->qvel[i];
  for (int t = 0; t < m->ntendon; t++) {
    double f = -m->tendon_stiffness[t] * (d->ten_length[t] - m->tendon_lengthspring[t]) - m->tendon_damping[t] * d->ten_velocity[t];
    if (f != 0)
      for (int k = 0; k < nv; k++) d->qfrc_passive[k] += d->ten_J[t * nv + k] * f;
  }
}

/* ------------------------------------------------------------------ S11 actuation */
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* userdata floats per actuator: mujoco-py's PID keeps 3 (integral, last error, last derivative); a model with a cascaded-PI
 * actuator (actuator_user[0] = 1) keeps 6 for every actuator (position-loop integral, last error, last derivative,
 * velocity-loop integral, smoothed set-point, "a step has been taken" flag). */
static int rgo_casc_gravcomp = 1;
void rgo_set_casc_gravcomp(int on) { rgo_casc_gravcomp = on; }   /* experiment switch (tests/tools) */
int rgo_pid_stride(const rgo_model* m) {
  static int env_read = 0;
  if (!env_read) { const char* e = getenv("RGO_CASC_GRAVCOMP"); if (e) rgo_set_casc_gravcomp(atoi(e)); env_read = 1; }
  for (int i = 0; i < m->nu; i++)
    if (m->actuator_user0[i] == 1.0 && m->actuator_biastype[i] == BIAS_USER) return 6;
  return 3;
}

/* mujoco-py's PID bias callback (mjpid.pyx, pinned by robogym/mujoco/constants.py:34-53 and
 * SURVEY.md Appendix D).  State per actuator in userdata: integral, last error, last derivative. */
static double pid_bias(const rgo_model* m, rgo_data* d, int id, int stride) {
  const double* g = m->actuator_gainprm + 10 * id;
  double dt = m->opt_timestep[0];
  double Kp = g[0], Ti = g[1], imax = g[2], Td = g[3], smooth = g[4], deadband = g[5];
  double err = d->ctrl[id] - d->actuator_length[id];
  if (fabs(err) < deadband) err = 0;
  double* ud = d->userdata + stride * id;
  double integ = clampd(ud[0] + err * dt, -imax, imax);
  double deriv = (err - ud[1]) / dt;
  deriv = (1.0 - smooth) * ud[2] + smooth * deriv;
  double f = Kp * (err + (Ti != 0 ? integ / Ti : 0.0) + Td * deriv);
  ud[0] = integ; ud[1] = err; ud[2] = deriv;
  double lo = m->actuator_forcerange[2 * id], hi = m->actuator_forcerange[2 * id + 1];
  if (lo != 0 || hi != 0) f = clampd(f, lo, hi);
  return f;
}

/* mujoco-py's cascaded-PI bias callback (mjpid.pyx, selected by actuator_user[0] = 1: the UR16e's default joint calibration,
 * robogym/assets/xmls/robot/ur16e/jointspec/calibrations/cascaded_pi/joint_actuations.xml:4-10).  mjpid.pyx is not in the
 * reference tree: this is its published law restated, pinned by the reference's impulse-response
 * fixture robogym/envs/rearrange/tests/test_rearrange_sim.py:135-230 (tests/test_reference_suite.py).
 *   gainprm = Kp_x Ti_x clamp_x Td_x dsmooth_x | Kp_v Ti_v clamp_v | ema max_vel
 *   set-point: exponential moving average of ctrl (taken over as is until the first step has been integrated),
 *   position PID loop -> desired velocity, clamped to max_vel (Kp_x = 0: ctrl is the desired velocity),
 *   velocity PI loop -> force, plus the joint's bias force (gravity / Coriolis compensation), clamped to forcerange. */
static double cascaded_pi_bias(const rgo_model* m, rgo_data* d, int id, int stride) {
  const double* g = m->actuator_gainprm + 10 * id;
  double dt = m->opt_timestep[0];
  double* ud = d->userdata + stride * id;
  static int warm = -1;   /* experiment switch (tests/tools): RGO_CASC_EMA_WARM=0 smooths from the first evaluation on */
  if (warm < 0) { const char* e = getenv("RGO_CASC_EMA_WARM"); warm = e ? atoi(e) : 1; }
  double ema = (ud[5] != 0 || !warm) ? g[8] * ud[4] + (1.0 - g[8]) * d->ctrl[id] : d->ctrl[id];
  { static int smooth = -1; if (smooth < 0) { const char* e = getenv("RGO_CASC_EMA"); smooth = e ? atoi(e) : 1; } if (!smooth) ema = d->ctrl[id]; }   /* experiment switch: RGO_CASC_EMA=0 = no set-point smoothing */
  ud[4] = ema;
  double des_vel = d->ctrl[id];
  if (g[0] != 0) {
    double err = ema - d->actuator_length[id];
    double integ = clampd(ud[0] + err * dt, -g[2], g[2]);
    double deriv = (1.0 - g[4]) * ud[2] + g[4] * (err - ud[1]) / dt;
    des_vel = g[0] * (err + (g[1] != 0 ? integ / g[1] : 0.0) + g[3] * deriv);
    ud[0] = integ; ud[1] = err; ud[2] = deriv;
  }
  static int vclamp = -1;   /* experiment switch (tests/tools): RGO_CASC_VCLAMP=0 drops the max_vel clamp */
  if (vclamp < 0) { const char* e = getenv("RGO_CASC_VCLAMP"); vclamp = e ? atoi(e) : 1; }
  if (vclamp) des_vel = clampd(des_vel, -g[9], g[9]);
  double errv = des_vel - d->actuator_velocity[id];
  double integv = clampd(ud[3] + errv * dt, -g[7], g[7]);
  ud[3] = integv;
  double f = g[5] * (errv + (g[6] != 0 ? integv / g[6] : 0.0));
  if (rgo_casc_gravcomp && m->actuator_trntype[id] == TRN_JOINT) f += d->qfrc_bias[m->jnt_dofadr[m->actuator_trnid[id]]];
  double lo = m->actuator_forcerange[2 * id], hi = m->actuator_forcerange[2 * id + 1];
  if (lo != 0 || hi != 0) f = clampd(f, lo, hi);
  return f;
}

static void rgo_actuation(const rgo_model* m, rgo_data* d) {
  int nv = m->nv;
  memset(d->qfrc_actuator, 0, sizeof(double) * nv);
  if (m->opt_disableflags[0] & DSBL_ACTUATION) { memset(d->actuator_force, 0, sizeof(double) * m->nu); return; }
  for (int i = 0; i < m->nu; i++) {
    double v = 0;
    for (int k = 0; k < nv; k++) v += d->actuator_moment[i * nv + k] * d->qvel[k];
    d->actuator_velocity[i] = v;
    double ctrl = d->ctrl[i];
    if (m->actuator_ctrllimited[i] && !(m->opt_disableflags[0] & DSBL_CLAMPCTRL))
      ctrl = clampd(ctrl, m->actuator_ctrlrange[2 * i], m->actuator_ctrlrange[2 * i + 1]);
    double gain = 0, bias = 0;
    if (m->actuator_gaintype[i] == GAIN_FIXED) gain = m->actuator_gainprm[10 * i];
    else if (m->actuator_gaintype[i] == GAIN_USER) gain = 0; /* mujoco-py installs a zero gain callback */
    const double* bp = m->actuator_biasprm + 10 * i;
    if (m->actuator_biastype[i] == BIAS_AFFINE) bias = bp[0] + bp[1] * d->actuator_length[i] + bp[2] * v;
    else if (m->actuator_biastype[i] == BIAS_USER && m->opt_pid[0]) {
      int stride = rgo_pid_stride(m);
      if (stride * (i + 1) <= m->nuserdata) bias = m->actuator_user0[i] == 1.0 ? cascaded_pi_bias(m, d, i, stride) : pid_bias(m, d, i, stride);
    }
    double f = gain * ctrl + bias;
    if (m->actuator_forcelimited[i]) f = clampd(f, m->actuator_forcerange[2 * i], m->actuator_forcerange[2 * i + 1]);
    d->actuator_force[i] = f;
    for (int k = 0; k < nv; k++) d->qfrc_actuator[k] += d->actuator_moment[i * nv + k] * f;
  }
}

/* ------------------------------------------------------------------ S12 smooth acceleration */
static void rgo_smooth(const rgo_model* m, rgo_data* d) {
  int nv = m->nv;
  memset(d->qfrc_applied_total, 0, sizeof(double) * nv);
  for (int b = 1; b < m->nbody; b++) {
    const double* x = d->xfrc_applied + 6 * b;
    if (x[0] == 0 && x[1] == 0 && x[2] == 0 && x[3] == 0 && x[4] == 0 && x[5] == 0) continue;
    for (int dof = 0; dof < nv; dof++)
      if (dof_in_body(m, b, dof)) {
        double jp[3], jr[3];
        jac_col(d, dof, d->xipos + 3 * b, jp, jr);
        d->qfrc_applied_total[dof] += dot3(jp, x) + dot3(jr, x + 3);
      }
  }
  for (int i = 0; i < nv; i++)
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i] + d->qfrc_applied_total[i];
  chol_factor(d->Mchol, d->M, nv);
  chol_solve(d->Mchol, d->qacc_smooth, d->qfrc_smooth, nv);
}

#include "rgo_collision.inc"
#include "rgo_constraint.inc"

/* ------------------------------------------------------------------ S15 Euler */
static void quat_integrate(double* q, const double* w, double h) {
  double ang = norm3(w) * h;
  if (ang < MINVAL) return;
  double ax[3] = {w[0], w[1], w[2]};
  normalize3(ax);
  double dq[4], r[4];
  axisangle2quat(dq, ax, ang);
  quat_mul(r, q, dq);
  quat_norm(r);
  memcpy(q, r, sizeof r);
}

static void rgo_euler(const rgo_model* m, rgo_data* d) {
  int nv = m->nv;
  double h = m->opt_timestep[0];
  double* qacc = (double*)malloc(sizeof(double) * nv);
  int anydamp = 0;
  for (int i = 0; i < nv; i++) anydamp |= m->dof_damping[i] > 0;
  if (anydamp) {
    double* A = (double*)malloc(sizeof(double) * nv * nv);
    double* L = (double*)malloc(sizeof(double) * nv * nv);
    double* rhs = (double*)malloc(sizeof(double) * nv);
    memcpy(A, d->M, sizeof(double) * nv * nv);
    for (int i = 0; i < nv; i++) { A[i * nv + i] += h * m->dof_damping[i]; rhs[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
    chol_factor(L, A, nv);
    chol_solve(L, qacc, rhs, nv);
    free(A); free(L); free(rhs);
  } else memcpy(qacc, d->qacc, sizeof(double) * nv);
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    switch (m->jnt_type[j]) {
      case JNT_FREE:
        for (int a = 0; a < 3; a++) d->qpos[qa + a] += h * d->qvel[da + a];
        quat_integrate(d->qpos + qa + 3, d->qvel + da + 3, h);
        break;
      case JNT_BALL: quat_integrate(d->qpos + qa, d->qvel + da, h); break;
      default: d->qpos[qa] += h * d->qvel[da];
    }
  }
  d->time[0] += h;
  if (rgo_pid_stride(m) == 6)   /* cascaded-PI "a step has been taken" flag (mjpid.pyx: d.time > 0) */
    for (int i = 0; i < m->nu; i++) if (6 * (i + 1) <= m->nuserdata) d->userdata[6 * i + 5] = 1;
  free(qacc);
}

/* ------------------------------------------------------------------ drivers */
static int bad(double x) { return !(x == x) || fabs(x) > 1e10; }

/* ------------------------------------------------------------------ S14 sensors */
/* smallest non-negative root of a t^2 + 2 b t + c = 0 (both roots in xx), or -1 */
static double ray_quad(double a, double b, double c, double xx[2]) {
  double det = b * b - a * c;
  if (det < MINVAL || a < MINVAL) { xx[0] = xx[1] = -1; return -1; }
  det = sqrt(det);
  xx[0] = (-b - det) / a; xx[1] = (-b + det) / a;
  if (xx[0] >= 0) return xx[0];
  if (xx[1] >= 0) return xx[1];
  return -1;
}
/* distance along the ray pnt + t vec to a sphere / capsule volume at (pos, mat), -1 = miss (a point inside always hits) */
static double ray_site(const double* pos, const double* mat, const double* size, const double* pnt, const double* vec, int type) {
  double dif[3], lp[3], lv[3], xx[2], x = -1;
  sub3(dif, pnt, pos);
  mulmatT3(lp, mat, dif);
  mulmatT3(lv, mat, vec);
  if (type == GEOM_SPHERE) return ray_quad(dot3(lv, lv), dot3(lv, lp), dot3(lp, lp) - size[0] * size[0], xx);
  if (type != GEOM_CAPSULE) return -1;
  double sol = ray_quad(lv[0] * lv[0] + lv[1] * lv[1], lv[0] * lp[0] + lv[1] * lp[1], lp[0] * lp[0] + lp[1] * lp[1] - size[0] * size[0], xx);
  if (sol >= 0 && fabs(lp[2] + sol * lv[2]) <= size[1]) x = sol;
  for (int cap = 0; cap < 2; cap++) {
    double zc = cap == 0 ? size[1] : -size[1], q[3] = {lp[0], lp[1], lp[2] - zc};
    ray_quad(dot3(lv, lv), dot3(lv, q), dot3(q, q) - size[0] * size[0], xx);
    for (int i = 0; i < 2; i++) {
      if (xx[i] < 0) continue;
      double z = lp[2] + xx[i] * lv[2];
      if ((cap == 0 ? z >= size[1] : z <= -size[1]) && (x < 0 || xx[i] < x)) x = xx[i];
    }
  }
  return x;
}
/* force (normal, 2 tangential) and torque (torsional, 2 rolling) of contact c in its own frame, from the solver's rows */
static void contact_wrench(const rgo_model* m, const rgo_data* d, int c, double* F) {
  const double* con = d->contact + RGO_CON_STRIDE * c;
  for (int k = 0; k < 6; k++) F[k] = 0;
  int seen = 0;
  for (int r = 0; r < d->nefc; r++) {
    if (d->efc_id[r] != c) continue;
    if (d->efc_type[r] == ROW_CONTACT_ELL || d->efc_type[r] == ROW_CONTACT_ELLF) F[seen++] = d->efc_force[r];
    else if (d->efc_type[r] == ROW_CONTACT) {
      int dim = (int)con[19];
      if (dim == 1) { F[0] = d->efc_force[r]; continue; }
      int k = 1 + seen / 2, sgn = seen & 1;          /* pyramid edges: direction k, + then - */
      double mu = con[14 + k - 1];
      F[0] += d->efc_force[r];
      F[k] += (sgn == 0 ? mu : -mu) * d->efc_force[r];
      seen++;
    }
  }
}

/* mj_rnePostConstraint for one body: the wrench its parent transmits to the subtree rooted at `body`
   (cfrc_int = sum over the subtree of I a + v x* I v - external wrench, with xfrc_applied and the contact forces as the
   external part; a(world) = -gravity), as [torque about the world origin, force] in world axes */
static void subtree_wrench(const rgo_model* m, const rgo_data* d, int body, double* W) {
  int nb = m->nbody;
  double* A = (double*)calloc((size_t)6 * nb, sizeof(double));
  char* in = (char*)calloc(nb, 1);
  if (!(m->opt_disableflags[0] & DSBL_GRAVITY)) { A[3] = -m->opt_gravity[0]; A[4] = -m->opt_gravity[1]; A[5] = -m->opt_gravity[2]; }
  for (int k = 0; k < 6; k++) W[k] = 0;
  for (int b = 1; b < nb; b++) {
    double* Ab = A + 6 * b;
    memcpy(Ab, A + 6 * m->body_parentid[b], 6 * sizeof(double));
    for (int k = 0; k < m->body_dofnum[b]; k++) {
      int dof = m->body_dofadr[b] + k;
      for (int i = 0; i < 6; i++) Ab[i] += d->dof_Sdot[6 * dof + i] * d->qvel[dof] + d->dof_S[6 * dof + i] * d->qacc[dof];
    }
    in[b] = b == body || in[m->body_parentid[b]];
    if (!in[b]) continue;
    double F[6], H[6], G[6];
    inertia_mul(F, d->body_I10 + 10 * b, Ab);
    inertia_mul(H, d->body_I10 + 10 * b, d->cvel + 6 * b);
    cross_force(G, d->cvel + 6 * b, H);
    for (int i = 0; i < 6; i++) W[i] += F[i] + G[i];
    const double* x = d->xfrc_applied + 6 * b;   /* force, torque at the body's centre of mass */
    double t[3];
    cross3(t, d->xipos + 3 * b, x);
    for (int i = 0; i < 3; i++) { W[i] -= x[3 + i] + t[i]; W[3 + i] -= x[i]; }
  }
  for (int c = 0; c < d->ncon; c++) {
    const double* con = d->contact + RGO_CON_STRIDE * c;
    int b1 = m->geom_bodyid[(int)con[20]], b2 = m->geom_bodyid[(int)con[21]];
    if (in[b1] == in[b2]) continue;              /* outside, or internal to the subtree */
    double F[6], f[3] = {0, 0, 0}, tq[3] = {0, 0, 0}, t[3];
    contact_wrench(m, d, c, F);
    for (int k = 0; k < 3; k++)
      for (int i = 0; i < 3; i++) { f[i] += F[k] * con[4 + 3 * k + i]; tq[i] += F[3 + k] * con[4 + 3 * k + i]; }
    double sgn = in[b2] ? 1 : -1;                /* the contact force acts on geom2's body, its opposite on geom1's */
    cross3(t, con + 1, f);
    for (int i = 0; i < 3; i++) { W[i] -= sgn * (tq[i] + t[i]); W[3 + i] -= sgn * f[i]; }
  }
  free(A); free(in);
}

/* data.sensordata after mj_forward: joint positions, touch sensors (the 5 fingertip pads of the hand,
   robogym/assets/xmls/robot/shadowhand/assets.xml:135-142: normal forces of the contacts on the site's body whose point --
   or the ray from it along the contact normal -- lies in the site's volume), and the force / torque sensors of the UR16e's
   tool flange (robogym/assets/xmls/robot/ur16e/base.xml:48-49, read by robogym/robot/ur16e/mujoco/joint_controlled_arm.py:35-45):
   the wrench between the site's body and its parent, in the site's frame, the torque taken about the site. */
static void rgo_sensors(const rgo_model* m, rgo_data* d) {
  for (int i = 0; i < m->nsensor; i++) {
    int adr = m->sensor_adr[i], obj = m->sensor_objid[i];
    for (int k = 0; k < m->sensor_dim[i]; k++) d->sensordata[adr + k] = 0;
    if (m->sensor_type[i] == 8) d->sensordata[adr] = d->qpos[m->jnt_qposadr[obj]];
    if (m->sensor_type[i] == 4 || m->sensor_type[i] == 5) {
      double W[6], t[3];
      subtree_wrench(m, d, m->site_bodyid[obj], W);
      if (m->sensor_type[i] == 5) {            /* torque about the site instead of the world origin */
        cross3(t, d->site_xpos + 3 * obj, W + 3);
        for (int k = 0; k < 3; k++) W[k] -= t[k];
      }
      mulmatT3(d->sensordata + adr, d->site_xmat + 9 * obj, m->sensor_type[i] == 4 ? W + 3 : W);
    }
    if (m->sensor_type[i] != 0) continue;
    int body = m->site_bodyid[obj];
    for (int c = 0; c < d->ncon; c++) {
      const double* con = d->contact + RGO_CON_STRIDE * c;
      int b1 = m->geom_bodyid[(int)con[20]], b2 = m->geom_bodyid[(int)con[21]];
      if (body != b1 && body != b2) continue;
      double f = 0;
      int rows = 0;
      for (int r = 0; r < d->nefc; r++) {
        if (d->efc_id[r] != c) continue;
        if (d->efc_type[r] == ROW_CONTACT) { f += d->efc_force[r]; rows++; }
        else if (d->efc_type[r] == ROW_CONTACT_ELL) { f = d->efc_force[r]; rows++; }
      }
      if (!rows || f <= 0) continue;
      double ray[3] = {con[4], con[5], con[6]};
      if (body == b2) scl3(ray, ray, -1);
      if (ray_site(d->site_xpos + 3 * obj, d->site_xmat + 9 * obj, m->site_size + 3 * obj, con + 1, ray, m->site_type[obj]) >= 0)
        d->sensordata[adr] += f;
    }
  }
}

void rgo_forward(const rgo_model* m, rgo_data* d) {
  rgo_kinematics(m, d);
  rgo_inertia(m, d);
  rgo_tendon(m, d);
  rgo_transmission(m, d);
  rgo_massmatrix(m, d);
  rgo_collision(m, d);
  rgo_velocity(m, d);
  rgo_passive(m, d);
  rgo_bias(m, d);
  rgo_actuation(m, d);
  rgo_smooth(m, d);
  rgo_constraint(m, d);
  rgo_sensors(m, d);
}

void rgo_step(const rgo_model* m, rgo_data* d) {
  for (int i = 0; i < m->nq; i++) if (bad(d->qpos[i])) { d->warning |= 4; }
  for (int i = 0; i < m->nv; i++) if (bad(d->qvel[i])) { d->warning |= 4; }
  if (d->warning & 4) { int w = d->warning; rgo_reset(m, d); d->warning = w; }
  rgo_forward(m, d);
  for (int i = 0; i < m->nv; i++) if (bad(d->qacc[i])) { d->warning |= 4; }
  if (d->warning & 4) { int w = d->warning; rgo_reset(m, d); d->warning = w; rgo_forward(m, d); }
  rgo_euler(m, d);
}

void rgo_env_step(const rgo_model* m, rgo_data* d, int nsub) {
  for (int i = 0; i < nsub; i++) rgo_step(m, d);
  rgo_forward(m, d);
}
