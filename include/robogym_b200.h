/* robogym_b200.h -- C ABI of the B200-native batched step engine (librobogym_b200.so).
 *
 * This is the drop-in boundary for robogym's physics path.  Each entry point names the
 * reference interface it replaces (paths relative to /root/reference/):
 *
 *   rg_model_load      <- mujoco_py.load_model_from_xml(xml) + MjSim(model, nsubsteps)
 *                         (robogym/mujoco/mujoco_xml.py:249-260).  The MJCF itself is compiled on
 *                         the host by robogym_b200.mjcf into the blob of include/rg_model_fields.h.
 *   rg_model_set_field <- in-place edits of sim.model.<array> by randomizers / modifiers
 *                         (robogym/wrappers/randomizations.py:84,139,188,302,589,643,713,745;
 *                          robogym/envs/dactyl/common/mujoco_modifiers.py:95-101).
 *   rg_batch_create/bind <- the mjData that MjSim owns (qpos, qvel, ctrl, userdata = PID state,
 *                         qacc_warmstart, xfrc_applied, time), here one row per environment in
 *                         caller-owned device tensors (torch) -- robogym reads/writes them through
 *                         SimulationInterface.qpos/qvel/set_qpos/... (simulation_interface.py:127-172)
 *                         and robot code writes sim.data.ctrl (robot/shadow_hand/mujoco/mujoco_shadow_hand.py:120-137).
 *   rg_step            <- SimulationInterface.step(): sim.step() [nsubsteps x mj_step, with the
 *                         mujoco-py PID callback enabled by cymj.set_pid_control,
 *                         simulation_interface.py:86-88] followed by sim.forward()
 *                         (robogym/mujoco/simulation_interface.py:176-189; robogym/robot_env.py:837).
 *   rg_forward         <- SimulationInterface.forward() (simulation_interface.py:203-207).
 *   rg_reset           <- SimulationInterface.reset() = mj_resetData (simulation_interface.py:191-195).
 *   rg_set_const       <- SimulationInterface.set_constants() = mj_setConst (simulation_interface.py:197-201).
 *
 * Conventions: every function returns 0 on success or a negative code and sets a thread-local
 * message readable with rg_last_error(); no exceptions or callbacks cross the ABI.  All device
 * work is stream-ordered and asynchronous on the stream the caller passes (a cudaStream_t as
 * void*).  The set-up calls (rg_model_load, rg_batch_create[_ex]) allocate engine-internal buffers and may synchronise with
 * the device; the stepping calls (rg_step, rg_step_subset, rg_forward, rg_reset, rg_model_set_field_async) never do, and
 * nothing ever allocates or frees the bound tensors.
 * State layout: row-major [nenv][n] float32 (int32 for ncon/warn); one environment per row.
 */
#ifndef ROBOGYM_B200_H
#define ROBOGYM_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rg_model rg_model;
typedef struct rg_batch rg_batch;

/* bindable per-environment arrays (rg_batch_bind) */
enum rg_field {
  RG_FIELD_QPOS = 0,       /* [nenv][nq]          in/out */
  RG_FIELD_QVEL = 1,       /* [nenv][nv]          in/out */
  RG_FIELD_CTRL = 2,       /* [nenv][nu]          in     */
  RG_FIELD_PID = 3,        /* [nenv][npid]        in/out : mujoco-py controller state in userdata; npid = rg_model_dim(m, "npid") = 3*nu
                              (PID: integral, last error, last derivative) or 6*nu when the model has a cascaded-PI actuator
                              (actuator user="1": + velocity-loop integral, smoothed set-point, step-taken flag) */
  RG_FIELD_WARMSTART = 4,  /* [nenv][nv]          in/out : qacc_warmstart */
  RG_FIELD_TIME = 5,       /* [nenv]              in/out (optional) */
  RG_FIELD_XFRC = 6,       /* [nenv][nbody*6]     in     (optional) : data.xfrc_applied */
  RG_FIELD_TIMESTEP = 7,   /* [nenv]              in     (optional) : per-env opt.timestep override */
  RG_FIELD_SITE_XPOS = 8,  /* [nenv][nsite*3]     out    (optional) */
  RG_FIELD_BODY_XPOS = 9,  /* [nenv][nbody*3]     out    (optional) */
  RG_FIELD_BODY_XQUAT = 10,/* [nenv][nbody*4]     out    (optional) */
  RG_FIELD_GEOM_XPOS = 11, /* [nenv][ngeom*3]     out    (optional) */
  RG_FIELD_ACT_FORCE = 12, /* [nenv][nu]          out    (optional) */
  RG_FIELD_QACC = 13,      /* [nenv][nv]          out    (optional) */
  RG_FIELD_CONTACT = 14,   /* [nenv][contact capacity][4] out (optional): geom1, geom2, dist, condim (rg_batch_capacity) */
  RG_FIELD_NCON = 15,      /* [nenv] int32        out    (optional) */
  RG_FIELD_WARN = 16,      /* [nenv] int32        in/out (optional): bit0 contact buffer full, bit1 row buffer full, bit2 bad state -> reset,
                              bit3 MPR, bit4 tendon Jacobian too dense, bit5 a contact touched more dofs than the batch allows (dropped) */
  RG_FIELD_DBG = 17,       /* [nenv][rg_batch_dbg_size] out (optional): stage dump used by the parity tests */
  RG_FIELD_BODY_XVEL = 18, /* [nenv][nbody*6]     out    (optional): angular, linear velocity of every body frame in world axes
                              = data.get_body_xvelr / get_body_xvelp (robogym/robot/ur16e/mujoco/joint_controlled_arm.py:32,
                              robogym/envs/rearrange/simulation/base.py:465-472) */
  RG_FIELD_MOCAP_POS = 19, /* [nenv][nmocap*3]    in     (optional): data.mocap_pos (world coordinates); unbound = the mocap bodies' model pose.
                              robogym moves the UR16e tool centre point by a mocap body welded to it
                              (robogym/robot/control/tcp/mocap_solver.py:41-46, robogym/assets/xmls/robot/ur16e/tcp_mocap.xml:2) */
  RG_FIELD_MOCAP_QUAT = 20,/* [nenv][nmocap*4]    in     (optional): data.mocap_quat */
  RG_FIELD_SENSORDATA = 21,/* [nenv][nsensordata] out    (optional): data.sensordata after the launch's last forward pass -- joint positions and
                              touch sensors (robogym/assets/xmls/robot/shadowhand/assets.xml:135-142); force / torque sensors
                              (robogym/assets/xmls/robot/ur16e/base.xml:48-49): wrench between the site's body and its parent
                              (cfrc_int of mj_rnePostConstraint), site frame */
  RG_NFIELDS = 22
};
#define RG_MAX_CONTACTS 32   /* DEFAULT contact capacity of a batch (rg_batch_create); rg_batch_create_ex picks another */
#define RG_MAX_PARAM_OVERRIDES 16

int rg_model_load(const void* blob, size_t len, int device, rg_model** out);
void rg_model_destroy(rg_model* m);
/* value of a dimension of rg_model_fields.h (nq, nv, nu, nbody, ...) or -1 */
int rg_model_dim(const rg_model* m, const char* name);
/* id of a named object, as mjModel.body_name2id / joint_name2id / geom_name2id / site_name2id / actuator_name2id /
 * tendon_name2id / sensor_name2id do (robogym reaches them through sim.model); objtype is "body", "joint",
 * "geom", "site", "actuator", "tendon", "mesh", "sensor" or "equality".  -1 when the name is unknown or the blob was
 * packed without its name tables. */
int rg_model_name2id(const rg_model* m, const char* objtype, const char* name);
const char* rg_model_id2name(const rg_model* m, const char* objtype, int id);   /* NULL when out of range; "" for an unnamed object */
/* overwrite a model array (host float64 / int32 values, `count` elements) and re-upload it.  The upload is ordered on
 * `stream` (set_field: the legacy default stream): launches queued before it keep the old values.  The host buffers may be
 * reused as soon as the call returns.  Do not call it from two threads for the same model at once. */
int rg_model_set_field(rg_model* m, const char* name, const void* data, size_t count);
int rg_model_set_field_async(rg_model* m, const char* name, const void* data, size_t count, void* stream);
/* floats per environment of the RG_FIELD_DBG dump; bytes of shared memory per environment (one warp) */
int rg_dbg_size(const rg_model* m);
int rg_scratch_bytes(const rg_model* m);

int rg_batch_create(const rg_model* m, int nenv, rg_batch** out);
/* The same with explicit capacities per environment (0 = default): contacts kept (the reference compiles nconmax=100,
 * robogym/assets/xmls/robot/shadowhand/assets.xml:6), single-row constraint elements = friction-loss + limit rows
 * (reference njmax=500 counts these plus the contact rows, assets.xml:5), dofs one contact may touch (<= 32).  Larger
 * capacities mean more shared memory per environment, i.e. fewer environments resident per SM; overflow at run time sets
 * a warning bit (RG_FIELD_WARN) and drops the surplus, like mj_warning does. */
int rg_batch_create_ex(const rg_model* m, int nenv, int contact_capacity, int row_capacity, int dofs_per_contact, rg_batch** out);
int rg_batch_capacity(const rg_batch* b, int* contacts, int* rows, int* dofs_per_contact);
/* floats per environment of this batch's RG_FIELD_DBG dump; shared-memory bytes per environment */
int rg_batch_dbg_size(const rg_batch* b);
int rg_batch_scratch_bytes(const rg_batch* b);
void rg_batch_destroy(rg_batch* b);
int rg_batch_bind(rg_batch* b, int field, void* device_ptr);
/* Per-environment override of a float model array (domain randomisation, SURVEY 5.6: robogym's wrappers write
 * sim.model.geom_friction / dof_damping / actuator_gainprm / opt_gravity / ... per env and per episode):
 * device_ptr is a float32 [nenv][count(name)] tensor that replaces the shared array `name` for each environment.
 * Up to RG_MAX_PARAM_OVERRIDES arrays; device_ptr == NULL removes the override.  body_pos rows of bodies attached
 * to the world must be given relative to rg_model_origin(). */
int rg_batch_bind_param(rg_batch* b, const char* name, void* device_ptr);
int rg_model_origin(const rg_model* m, float origin[3]);
/* Work-ordered scheduling (default on; RG_BALANCE=0 in the environment turns it off at create): every launch records a
 * per-environment work estimate and the next launch groups environments of similar cost into the same CTA, which
 * shortens the waits at the per-stage CTA barriers.  Results are independent of the setting.  No reference
 * counterpart: mujoco-py steps one MjSim at a time (robogym/mujoco/simulation_interface.py:176). */
int rg_batch_set_balance(rg_batch* b, int on);
/* launch geometry actually used (for reporting): CTAs, warps per CTA, dynamic shared bytes */
int rg_batch_launch_info(const rg_batch* b, int* ctas, int* warps_per_cta, int* smem_bytes);

/* nsub x mj_step, then `final_forward` x mj_forward (0..4: SimulationInterface.step ends with one sim.forward(); the
 * observation path of RobotEnv runs more of them (robogym/robot_env.py:677, observation/mujoco.py:27) and mujoco-py's
 * PID state in userdata advances in each); derived outputs written once at the end */
int rg_step(rg_batch* b, int nsub, int final_forward, void* stream);
int rg_forward(rg_batch* b, void* stream);
/* the same for the environments whose mask byte (device memory, [nenv]) is non-zero; the others are untouched and cost
 * nothing (the launch covers only the selected environments).  What a Python loop over MjSim objects does when it
 * calls sim.forward()/sim.step() on some environments only (goal switches, resets). */
int rg_step_subset(rg_batch* b, const uint8_t* mask_device, int nsub, int final_forward, void* stream);
/* mj_setConst per environment (SimulationInterface.set_constants, robogym/mujoco/simulation_interface.py:197-201, which the
 * reference calls after its randomisers edited masses, inertias, armatures ...): recomputes, from each selected environment's
 * own parameter view at qpos0, the constants MuJoCo derives from the model -- dof_invweight0, body_invweight0,
 * tendon_invweight0, tendon_length0, body_subtreemass, opt_meaninertia -- and writes them into that environment's row of the
 * arrays bound with rg_batch_bind_param under those names (constants that are not bound per environment are not written:
 * the shared model is immutable while launches may be in flight; at least one must be bound).  One launch, asynchronous on
 * `stream`; mask as in rg_step_subset, NULL = every environment. */
int rg_set_const(rg_batch* b, const uint8_t* mask_device, void* stream);
/* mj_resetData for the environments whose mask byte is non-zero (mask == NULL: all) */
int rg_reset(rg_batch* b, const uint8_t* mask_device, void* stream);

const char* rg_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
