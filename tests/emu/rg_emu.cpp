/* rg_emu.cpp -- CPU EMULATION BUILD of the CUDA engine's device code.  TEST INFRASTRUCTURE ONLY.
 *
 * Compiles robogym_b200/csrc/rg_*.inl with -DRG_EMU: every warp phase becomes a loop over 32
 * lanes (see rg_defs.h).  It lets the CPU-only test tier (`pytest -m "not gpu"`) check the
 * kernel's fp32 logic against the fp64 oracle; it is not reachable from the package and the
 * product path never falls back to it.
 */
#define RG_EMU 1
#include "../../robogym_b200/csrc/rg_step.inl"
#include "../../robogym_b200/csrc/rg_host.h"

#include <stdio.h>
#include <stdlib.h>

struct RgeHandle { RgHostModel hm; RgLayout L; std::vector<float> scratch; std::vector<int> sep; const float* mocap_pos = nullptr; const float* mocap_quat = nullptr; float* sensordata = nullptr; };

extern "C" {

void* rge_create_ex(const void* blob, size_t len, int ncon, int nel, int tile) {
  RgeHandle* h = new RgeHandle();
  std::string err;
  if (!rg_host_load(blob, len, h->hm, err)) { fprintf(stderr, "rge_create: %s\n", err.c_str()); delete h; return nullptr; }
  h->L = rg_make_layout(h->hm.view, ncon ? ncon : RG_NCON, nel ? nel : RG_NEL, tile ? tile : RG_TILE);
  h->scratch.assign(h->L.total, 0.0f);
  return h;
}
void* rge_create(const void* blob, size_t len) { return rge_create_ex(blob, len, 0, 0, 0); }
int rge_ncon(void* hv) { return ((RgeHandle*)hv)->L.ncon; }
int rge_pidw(void* hv) { return ((RgeHandle*)hv)->hm.view.pidw; }
/* the blob's name tables as parsed by the shared host loader (what rg_model_name2id serves) */
int rge_name2id(void* hv, const char* typ, const char* name) {
  RgeHandle* h = (RgeHandle*)hv;
  auto it = h->hm.names.find(typ);
  if (it == h->hm.names.end()) return -1;
  for (size_t i = 0; i < it->second.size(); i++) if (it->second[i] == name) return (int)i;
  return -1;
}
#ifdef RG_STATS
void rge_stats(long long* out) { out[0] = rg_stat_support; out[1] = rg_stat_climb; out[2] = rg_stat_mpr; out[3] = rg_stat_mpr_hit; out[4] = rg_stat_maxsup; for (int i = 0; i < 128; i++) out[8 + i] = rg_stat_hist[i / 64][i % 64]; for (int i = 0; i < 16; i++) out[136 + i] = rg_stat_x[i]; for (int i = 0; i < 16; i++) out[152 + i] = rg_stat_iterhist[i]; for (int i = 0; i < 16; i++) out[168 + i] = rg_stat_fliphist[i]; }
#endif
/* data.mocap_pos / mocap_quat rows ([nenv][nmocap*3], [nenv][nmocap*4]) used by the following rge_step calls (or null) */
void rge_set_mocap(void* hv, const float* pos, const float* quat) { ((RgeHandle*)hv)->mocap_pos = pos; ((RgeHandle*)hv)->mocap_quat = quat; }
void rge_set_sensordata(void* hv, float* out) { ((RgeHandle*)hv)->sensordata = out; }   /* [nenv][nsensordata] or null */
/* mj_setConst of the (edited) model: any output may be null */
void rge_set_const(void* hv, float* dof_invweight0, float* body_invweight0, float* tendon_invweight0, float* tendon_length0, float* body_subtreemass, float* opt_meaninertia) {
  RgeHandle* h = (RgeHandle*)hv;
  RgSetConstOut o = {dof_invweight0, body_invweight0, tendon_invweight0, tendon_length0, body_subtreemass, opt_meaninertia};
  rg_env_setconst(&h->hm.view, h->L, h->scratch.data(), 0, o);
}
void rge_destroy(void* hv) { delete (RgeHandle*)hv; }
int rge_dbg_size(void* hv) { return rg_dbg_size(((RgeHandle*)hv)->hm.view, ((RgeHandle*)hv)->L.ncon); }
int rge_scratch_floats(void* hv) { return ((RgeHandle*)hv)->L.total; }
int rge_small_bytes(void* hv) { return (int)((RgeHandle*)hv)->hm.small_bytes; }
/* writable pointer to a model array (float32/int32 copy) so tests can edit parameters */
void* rge_model_field(void* hv, const char* name, int* count) {
  RgeHandle* h = (RgeHandle*)hv;
  RgModel& m = h->hm.view;
#define RG_DIM(n) const int n = m.n; (void)n;
#define RG_I(n, c)
#define RG_F(n, c)
#include "../../include/rg_model_fields.h"
#undef RG_DIM
#undef RG_I
#undef RG_F
#define RG_DIM(n)
#define RG_I(n, c) if (!strcmp(name, #n)) { *count = (c); return (void*)m.n; }
#define RG_F(n, c) if (!strcmp(name, #n)) { *count = (c); return (void*)m.n; }
#include "../../include/rg_model_fields.h"
#undef RG_DIM
#undef RG_I
#undef RG_F
  return nullptr;
}

void rge_step(void* hv, int nenv, float* qpos, float* qvel, float* ctrl, float* pid, float* warm, float* time, const float* xfrc,
              const float* timestep, float* site_xpos, float* body_xpos, float* body_xquat, float* geom_xpos, float* act_force, float* qacc,
              float* contact, int* ncon, int* warn, float* dbg, int nsub, int final_forward) {
  RgeHandle* h = (RgeHandle*)hv;
  RgBatchIO io;
  io.nenv = nenv; io.qpos = qpos; io.qvel = qvel; io.ctrl = ctrl; io.pid = pid; io.warm = warm; io.time = time; io.xfrc = xfrc;
  io.timestep = timestep; io.site_xpos = site_xpos; io.body_xpos = body_xpos; io.body_xquat = body_xquat; io.geom_xpos = geom_xpos;
  io.act_force = act_force; io.qacc = qacc; io.contact = contact; io.ncon = ncon; io.warn = warn; io.dbg = dbg; io.cost = nullptr; io.body_xvel = nullptr; io.mocap_pos = h->mocap_pos; io.mocap_quat = h->mocap_quat; io.sensordata = h->sensordata;
  if (h->sep.size() != (size_t)nenv * RG_NSEP) h->sep.assign((size_t)nenv * RG_NSEP, 0xfff);   /* like the engine's per-batch buffer */
  io.sep = h->sep.data();
  for (int env = 0; env < nenv; env++) rg_env_step(&h->hm.view, h->L, h->scratch.data(), 0, io, env, nsub, final_forward, 1);
}
}
