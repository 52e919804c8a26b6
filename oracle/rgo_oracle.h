/* rgo_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, double-precision, single-environment restatement of the physics
 * path that robogym runs through mujoco-py:
 *     SimulationInterface.step() = sim.step() (nsubsteps x mj_step) + sim.forward()
 *     (robogym/mujoco/simulation_interface.py:176-207, called from robogym/robot_env.py:837)
 * including mujoco-py's stateful PID and cascaded-PI actuator callbacks that robogym
 * switches on with cymj.set_pid_control (robogym/mujoco/simulation_interface.py:86-88;
 * parameter layouts robogym/mujoco/constants.py:34-53 and
 * robogym/assets/xmls/robot/ur16e/jointspec/calibrations/cascaded_pi/joint_actuations.xml:4).
 *
 * PARITY: the arithmetic of this path lives in mujoco-py==2.0.2.13 / MuJoCo 2.0
 * (robogym setup.py:16), which is not vendored in /root/reference and is not installable
 * in the build container, so no mujoco-py trajectory exists to compare against: PARITY
 * UNPINNED for everything except what follows.  This file restates MuJoCo's published
 * pipeline (kinematics -> tendons -> CRB mass matrix -> collision -> constraint rows with
 * solref/solimp impedance -> Newton solver on the pyramidal / elliptic cone convex problem
 * -> semi-implicit Euler with implicit joint damping) from its documentation.  Pinned:
 *   - the reference's recorded controller response: test_mocap_ik_impulse_response
 *     (robogym/envs/rearrange/tests/test_rearrange_sim.py:135-230: tool displacement 0.036 /
 *     0.0363 / 0.022 / 0.022 m +- 1e-3 and rise times after an impulse action, through the
 *     two-simulation MOCAP_IK loop with cascaded-PI joint controllers) passes unmodified on
 *     the mujoco_py shim with this oracle as the engine, and fails with the plain-PID
 *     calibration or without the controller's bias-force term (tests/test_reference_suite.py);
 *   - a documented real-MuJoCo output of a simulated quantity: the block heights 0.51167315 that its documentation prints for
 *     rearrange/blocks_train after env.reset() (docs/env_param_interface.md:32-38).  The
 *     unmodified reference environment, run on the mujoco_py shim with this oracle as the
 *     engine, reports exactly that number after its 4000 mj_steps of object stabilisation
 *     (tests/test_rearrange_reset_pin.py, tools/make_rearrange_reset_fixture.py);
 *   - the reference-owned fixtures listed in SURVEY.md section 8(c) (pure-numpy forward
 *     kinematics, cube mass, joint order, closed-loop PID tracking, resting behaviour, the
 *     recorded four-block stack), see tests/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 */
#ifndef RGO_ORACLE_H
#define RGO_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rgo_model rgo_model;
typedef struct rgo_data rgo_data;

rgo_model* rgo_model_load(const void* blob, size_t len);
void rgo_model_free(rgo_model* m);
/* pointer to a model array by rg_model_fields.h name; *is_int tells int32 vs double */
void* rgo_model_field(rgo_model* m, const char* name, int* count, int* is_int);
int rgo_model_dim(const rgo_model* m, const char* name);

rgo_data* rgo_data_new(const rgo_model* m);
void rgo_data_free(rgo_data* d);
/* pointer to a data array (qpos, qvel, ctrl, userdata, qacc_warmstart, xfrc_applied, time, xpos,
 * xquat, xmat, geom_xpos, geom_xmat, site_xpos, ten_length, ten_J, actuator_length,
 * actuator_force, M, qfrc_bias, qfrc_passive, qfrc_actuator, qfrc_smooth, qacc_smooth, qacc,
 * qfrc_constraint, cvel, contact, ncon, nefc, efc_J, efc_aref, efc_D, efc_force, solver_niter, warning) */
void* rgo_data_field(rgo_data* d, const char* name, int* count, int* is_int);

void rgo_reset(const rgo_model* m, rgo_data* d);      /* mj_resetData */
void rgo_forward(const rgo_model* m, rgo_data* d);    /* mj_forward (PID callback included) */
void rgo_step(const rgo_model* m, rgo_data* d);       /* mj_step */
/* SimulationInterface.step(): nsub x mj_step then mj_forward */
void rgo_env_step(const rgo_model* m, rgo_data* d, int nsub);
int rgo_pid_stride(const rgo_model* m);               /* userdata floats per actuator: 3 (PID) or 6 (model with a cascaded-PI actuator) */
void rgo_set_casc_gravcomp(int on);                   /* experiment switch for tests/tools: bias-force compensation of the cascaded-PI law */
/* spatial-tendon constants for mj_setConst: lengths and dense Jacobian at qpos */
void rgo_tendon_eval(const rgo_model* m, rgo_data* d, const double* qpos, double* length, double* J);

/* contact record layout in the "contact" field (doubles) */
#define RGO_CON_STRIDE 24
/* 0 dist, 1-3 pos, 4-12 frame (normal,t1,t2 rows), 13 includemargin, 14-18 friction[5],
 * 19 dim, 20 geom1, 21 geom2, 22 solref0 23 solref1 */

#ifdef __cplusplus
}
#endif
#endif
