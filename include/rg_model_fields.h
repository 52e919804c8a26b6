/* rg_model_fields.h -- the compiled-model blob layout, as an X-macro list.
 *
 * One list drives three consumers, so they cannot drift apart:
 *   - robogym_b200/modelblob.py  parses this file and packs/unpacks the blob,
 *   - oracle/rgo_oracle.c        (fp64 CPU restatement; test infrastructure),
 *   - robogym_b200/csrc (CUDA)   (fp32 sm_100a product path).
 *
 * The blob replaces what mujoco_py.load_model_from_xml() hands to MjSim in the
 * reference (robogym/mujoco/mujoco_xml.py:249-260): a compiled mjModel.  Field
 * names follow mjModel where a counterpart exists so that the mujoco_py shim
 * (robogym_b200/mujoco_py_shim) can expose them 1:1.
 *
 * Blob layout (little endian):
 *   char magic[8] = "RGMODEL1"; int32 ndim; int32 dims[ndim] (RG_DIM order);
 *   then every RG_I / RG_F array in list order, each starting on an 8-byte
 *   boundary; RG_I arrays are int32, RG_F arrays are float64.
 *
 * Usage: define RG_DIM(name), RG_I(name, count), RG_F(name, count) then include.
 * RG_IB / RG_FB mark the BIG arrays at the end of the list (they default to RG_I / RG_F).
 * `count` is a C expression over the dims (the includer provides them in scope).
 */

#ifndef RG_IB
#define RG_IB(n, c) RG_I(n, c)
#define RG_FB(n, c) RG_F(n, c)
#define RG_BIG_DEFAULTED
#endif

/* ---- dimensions ---- */
RG_DIM(nq)        /* generalized positions */
RG_DIM(nv)        /* degrees of freedom */
RG_DIM(nu)        /* actuators */
RG_DIM(nbody)
RG_DIM(njnt)
RG_DIM(ngeom)
RG_DIM(nsite)
RG_DIM(ntendon)
RG_DIM(nwrap)     /* tendon path elements */
RG_DIM(nmesh)
RG_DIM(nmeshvert) /* total convex-hull vertices over all meshes */
RG_DIM(nmeshadj)  /* total hull edge-adjacency entries */
RG_DIM(nmeshface) /* total hull triangles */
RG_DIM(npair)     /* statically filtered candidate geom pairs */
RG_DIM(nlevel)    /* kinematic-tree depth levels */
RG_DIM(nmaskw)    /* 32-bit words per body dof-ancestor mask */
RG_DIM(nuserdata)
RG_DIM(nconmax)
RG_DIM(njmax)
RG_DIM(neq)
RG_DIM(nmocap)
RG_DIM(nsensor)
RG_DIM(nsensordata)

/* ---- options (mjOption) ---- */
RG_F(opt_timestep, 1)
RG_F(opt_gravity, 3)
RG_F(opt_tolerance, 1)
RG_F(opt_impratio, 1)
RG_F(opt_mpr_tolerance, 1)
RG_F(opt_ls_tolerance, 1)
RG_F(opt_meaninertia, 1)     /* mjModel.stat.meaninertia: mean diag(M) at qpos0 */
RG_I(opt_iterations, 1)
RG_I(opt_ls_iterations, 1)
RG_I(opt_mpr_iterations, 1)
RG_I(opt_cone, 1)            /* 0 pyramidal, 1 elliptic */
RG_I(opt_disableflags, 1)    /* RG_DSBL_* bits */
RG_I(opt_pid, 1)             /* 1 after cymj.set_pid_control: user gain/bias = mujoco-py PID */

/* ---- bodies ---- */
RG_I(body_parentid, nbody)
RG_I(body_rootid, nbody)     /* first ancestor below world (kinematic tree root) */
RG_I(body_weldid, nbody)
RG_I(body_mocapid, nbody)
RG_I(body_jntadr, nbody)
RG_I(body_jntnum, nbody)
RG_I(body_dofadr, nbody)
RG_I(body_dofnum, nbody)
RG_I(body_geomadr, nbody)
RG_I(body_geomnum, nbody)
RG_I(body_level, nbody)
RG_I(body_order, nbody)      /* body ids sorted by (level, id) */
RG_I(level_adr, nlevel + 1)  /* ranges into body_order per level */
RG_I(body_dofmask, nbody * nmaskw) /* bit d set iff dof d moves this body */
RG_F(body_pos, nbody * 3)
RG_F(body_quat, nbody * 4)
RG_F(body_ipos, nbody * 3)
RG_F(body_iquat, nbody * 4)
RG_F(body_mass, nbody)
RG_F(body_subtreemass, nbody)
RG_F(body_inertia, nbody * 3)
RG_F(body_invweight0, nbody * 2)

/* ---- joints / dofs ---- */
RG_I(jnt_type, njnt)         /* 0 free, 1 ball, 2 slide, 3 hinge (mjtJoint) */
RG_I(jnt_qposadr, njnt)
RG_I(jnt_dofadr, njnt)
RG_I(jnt_bodyid, njnt)
RG_I(jnt_limited, njnt)
RG_F(jnt_pos, njnt * 3)
RG_F(jnt_axis, njnt * 3)
RG_F(jnt_stiffness, njnt)
RG_F(jnt_range, njnt * 2)
RG_F(jnt_margin, njnt)
RG_F(jnt_solref, njnt * 2)
RG_F(jnt_solimp, njnt * 5)
RG_I(dof_bodyid, nv)
RG_I(dof_jntid, nv)
RG_I(dof_parentid, nv)
RG_F(dof_armature, nv)
RG_F(dof_damping, nv)
RG_F(dof_frictionloss, nv)
RG_F(dof_invweight0, nv)
RG_F(dof_solref, nv * 2)
RG_F(dof_solimp, nv * 5)
RG_F(qpos0, nq)
RG_F(qpos_spring, nq)

/* ---- geoms / sites / meshes ---- */
RG_I(geom_type, ngeom)       /* mjtGeom: 0 plane 2 sphere 3 capsule 4 ellipsoid 5 cylinder 6 box 7 mesh */
RG_I(geom_bodyid, ngeom)
RG_I(geom_dataid, ngeom)     /* mesh id or -1 */
RG_I(geom_contype, ngeom)
RG_I(geom_conaffinity, ngeom)
RG_I(geom_condim, ngeom)
RG_I(geom_priority, ngeom)
RG_F(geom_size, ngeom * 3)
RG_F(geom_pos, ngeom * 3)
RG_F(geom_quat, ngeom * 4)
RG_F(geom_rbound, ngeom)
RG_F(geom_aabb, ngeom * 6)      /* geom-frame bounding box: centre[3], half extents[3] (OBB cull before the narrow phase) */
RG_F(geom_friction, ngeom * 3)
RG_F(geom_margin, ngeom)
RG_F(geom_gap, ngeom)
RG_F(geom_solmix, ngeom)
RG_F(geom_solref, ngeom * 2)
RG_F(geom_solimp, ngeom * 5)
RG_I(site_bodyid, nsite)
RG_I(site_type, nsite)       /* mjtGeom of the site's volume (touch sensors test contact points against it) */
RG_F(site_pos, nsite * 3)
RG_F(site_quat, nsite * 4)
RG_F(site_size, nsite * 3)
RG_I(mesh_vertadr, nmesh)
RG_I(mesh_vertnum, nmesh)
RG_I(mesh_faceadr, nmesh)
RG_I(mesh_facenum, nmesh)

/* ---- tendons ---- */
RG_I(tendon_adr, ntendon)
RG_I(tendon_num, ntendon)
RG_I(tendon_limited, ntendon)
RG_F(tendon_range, ntendon * 2)
RG_F(tendon_margin, ntendon)
RG_F(tendon_stiffness, ntendon)
RG_F(tendon_damping, ntendon)
RG_F(tendon_frictionloss, ntendon)
RG_F(tendon_lengthspring, ntendon)
RG_F(tendon_length0, ntendon)
RG_F(tendon_invweight0, ntendon)
RG_F(tendon_solref_lim, ntendon * 2)
RG_F(tendon_solimp_lim, ntendon * 5)
RG_I(wrap_type, nwrap)       /* mjtWrap: 1 joint 2 pulley 3 site 4 sphere 5 cylinder */
RG_I(wrap_objid, nwrap)
RG_F(wrap_prm, nwrap)        /* joint coef | pulley divisor | sidesite id or -1 */

/* ---- actuators ---- */
RG_I(actuator_trntype, nu)   /* 0 joint, 3 tendon (mjtTrn) */
RG_I(actuator_trnid, nu)
RG_I(actuator_gaintype, nu)  /* 0 fixed, 2 user (mujoco_py.const.GAIN_*) */
RG_I(actuator_biastype, nu)  /* 0 none, 1 affine, 2 user */
RG_I(actuator_ctrllimited, nu)
RG_I(actuator_forcelimited, nu)
RG_F(actuator_gainprm, nu * 10)
RG_F(actuator_biasprm, nu * 10)
RG_F(actuator_ctrlrange, nu * 2)
RG_F(actuator_forcerange, nu * 2)
RG_F(actuator_gear, nu * 6)
RG_F(actuator_user0, nu)     /* actuator_user[:,0]: 1 selects cascaded PI, else PID */

/* ---- equality constraints / mocap (rearrange rows; unused by dactyl) ---- */
RG_I(eq_type, neq)
RG_I(eq_obj1id, neq)
RG_I(eq_obj2id, neq)
RG_I(eq_active, neq)
RG_F(eq_data, neq * 7)
RG_F(eq_solref, neq * 2)
RG_F(eq_solimp, neq * 5)

/* ---- sensors (mjtSensor: 0 touch, 4 force, 5 torque, 8 jointpos) ---- */
RG_I(sensor_type, nsensor)
RG_I(sensor_objid, nsensor)
RG_I(sensor_adr, nsensor)
RG_I(sensor_dim, nsensor)

/* ---- BIG arrays (kept in global memory by the CUDA engine; everything above is small enough
 *      to be staged into shared memory with one bulk copy).  Keep these LAST. ---- */
RG_FB(mesh_vert, nmeshvert * 3)       /* hull vertices, centred on the hull's volume centroid */
RG_IB(mesh_adjadr, nmeshvert + 1)     /* CSR into mesh_adj, global vertex ids */
RG_IB(mesh_adj, nmeshadj)             /* neighbour vertex ids, LOCAL to the mesh */
RG_IB(mesh_face, nmeshface * 3)       /* hull triangles, LOCAL vertex ids (rendering/inertia only) */
RG_IB(pair_geom1, npair)
RG_IB(pair_geom2, npair)

#ifdef RG_BIG_DEFAULTED
#undef RG_IB
#undef RG_FB
#undef RG_BIG_DEFAULTED
#endif
