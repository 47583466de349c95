// Shim of the few MuJoCo symbols the reference's main/humanoid_mpc.cpp touches directly
// (main:155-157: mj_forward(robot.model(), robot.data()); robot.data()->qfrc_bias[i + 6]).
// MuJoCo is not installed here; the plant behind RobotUtils is the GPU dynamics map f_D.
#pragma once
typedef double mjtNum;
struct mjOption { double timestep; double gravity[3]; double impratio; };
struct mjModel { int nq, nv, nu; mjOption opt; void* owner; };
struct mjData { mjtNum* qpos; mjtNum* qvel; mjtNum* ctrl; mjtNum* qfrc_bias; void* owner; };
// Recomputes d->qfrc_bias (Coriolis + gravity) for the state in d->qpos/qvel on the GPU.
void mj_forward(const mjModel* m, mjData* d);
