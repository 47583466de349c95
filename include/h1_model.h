/*
 * h1_model.h — plain-C description of the Unitree H1 kinematic tree as the iLQR hot path needs it.
 *
 * Two instances of this struct exist at run time, mirroring the reference's two model files:
 *   - the "dynamics" model  (reference: robots/h1_description/mjcf/h1.xml, loaded by
 *     RobotUtils::loadModel, src/common/robot_utils.cpp:19-55) — used by the one-step map f_D,
 *     its finite-difference linearization, the limit penalties and the CoM inside the
 *     line-search cost (src/ilqr/ilqr.cpp:409);
 *   - the "cost" model      (reference: robots/h1_description/urdf/h1.urdf, loaded by
 *     symDerivatives::symDerivatives, src/common/derivatives.cpp:26-39) — used by the
 *     CoM / end-effector / balance cost derivatives.
 * They share the tree topology but differ in a few inertial parameters.
 *
 * Body 0 is the floating base (pelvis). Body b>=1 is attached to parent[b] by ONE hinge whose
 * anchor is the body origin and whose axis is a coordinate axis of the body frame.
 * Hinge of body b drives qpos[6+b], qvel[5+b], and is driven by ctrl[b-1].
 */
#ifndef H1_MODEL_H
#define H1_MODEL_H

#ifdef __cplusplus
extern "C" {
#endif

#define H1_NB 20 /* bodies (1 free + 19 hinged) */
#define H1_NQ 26
#define H1_NV 25
#define H1_NX 51
#define H1_NU 19
#define H1_NFOOT 2
#define H1_NCP 4 /* contact points per foot sole */

typedef struct H1Model {
  int parent[H1_NB];          /* parent body index, -1 for the base */
  int axis[H1_NB];            /* hinge axis 0/1/2 = x/y/z of the body frame; -1 for the base */
  int has_rfix[H1_NB];        /* 1 if rfix differs from identity */
  double pos[H1_NB][3];       /* body origin in the parent frame */
  double rfix[H1_NB][9];      /* fixed rotation parent->body at zero joint angle, row-major */
  double mass[H1_NB];
  double ipos[H1_NB][3];      /* centre of mass in the body frame */
  double inertia[H1_NB][6];   /* rotational inertia about the CoM in body axes: xx yy zz xy xz yz */
  double armature[H1_NV];     /* reflected rotor inertia per dof (0 on the 6 base dofs) */
  double damping[H1_NV];      /* viscous joint damping per dof */
  double jnt_range[H1_NU][2]; /* hinge limits, used by the soft limit penalty only */
  double ctrl_range[H1_NU][2];/* actuator torque limits: clamp in f_D, soft penalty in the cost */
  int foot_body[H1_NFOOT];    /* ankle bodies: left, right (reference: robot_utils.cpp:44-45) */
  double foot_pts[H1_NFOOT][H1_NCP][3]; /* sole contact points in the ankle frame */
  double gravity[3];          /* world gravity vector (config.yaml mpc.gravity) */
  double timestep;            /* h (config.yaml mpc.physics_dt) */
  double contact_kn;          /* normal stiffness per point [N/m] */
  double contact_bn;          /* normal damping per point [N s/m] */
  double contact_bt;          /* tangential viscous friction per point [N s/m] */
  double contact_eps;         /* smoothing length of the penetration ramp [m] */
  double total_mass;
} H1Model;

/* Built-in models generated from the reference's h1.xml / h1.urdf by tools/gen_h1_model.py. */
const H1Model* h1_default_dynamics_model(void);
const H1Model* h1_default_cost_model(void);

#ifdef __cplusplus
}
#endif
#endif
