// The two callers on either side of the solver, batched, for the wiring the reference ships (7 DoF,
// ops/mgqp.ops:233-240, :256, :289-292): one thread per scenario.
//
//   cascade_assemble_kernel  = updateHook's QP assembly (src/mgqp.cpp:886-1143) for level 0: the translational
//       task rows [J 0] with goal -(Kp (pd - pc) + Kd (vd - vc) - Jdot qdot + ad - ac) (:1037-1057), the dynamics
//       rows [M -I] (:1138-1143), the log-barrier shaped acceleration limits (:1119-1123); written straight into
//       the solver's layout (CE n x p, one constraint per column, float32 values widened to fp64 as
//       solveNextStep's COPYMAT does, :700-706).  G = I, g0 = 0 and limitsMatrix = [-I; +I] (:1111-1112) are the
//       same for every scenario: one instance each (stride 0), the inequality block in column-compressed form.
//   cascade_project_kernel   = the part of solveNextHierarchy between two levels (:828-862): y = u (Z = I),
//       Z = I - V diag(sigma >= 1e-16) V' from a one-sided Jacobi SVD of the level's condition matrix (float32),
//       then level 2's QP data: A = conditions Z = row j of Z, a = goal - y_j, B = limitsMatrix Z = [-Z; Z].
//   cascade_finish_kernel    = res = y + Z u (:828), the failed-level fallbacks (:820-824), torques = tau + h (:1152).
//
// float32 arithmetic in ascending index order with one rounding per operation (__fmul_rn / __fadd_rn are never
// contracted).  The reference's own code is Eigen float32 (JacobiSVD, Eigen's product kernels) and Eigen is not in
// its tree: parity for these kernels is pinned to oracle/cascade_oracle.py, not to Eigen (see that file).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace mgqp {

constexpr int kDof = 7, kN = 14, kP0 = 10, kM = 28;

struct CascadeParams {
  long long batch;
  const float *q, *qd, *M, *h, *J, *Jd, *pos_des, *pos_cur, *vel_des, *vel_cur, *acc_des, *acc_cur, *q1_des;
  float kp_task, kd_task, kp_joint, kd_joint;
  float acc_p[kDof], acc_n[kDof], tau_p[kDof], tau_n[kDof], q_sup[kDof], q_inf[kDof];
  int joint_task;
  // level-0 QP (solver layout, fp64): CE0 [B][14][10], ce00 [B][10], ci0 [B][28]
  double *CE0, *ce00, *ci0;
  float* goal2;  // [B]
  // level-0 answer / level-2 QP / level-2 answer
  const double* x0; const double* f0; const int* st0;
  float* Z;      // [B][14][14]
  float* y;      // [B][14]
  double *CE2, *ce02, *CI2;  // [B][14][1], [B][1], [B][14][28]
  const double* x2; const double* f2; const int* st2;
  float* tracking; float* torques; int* status0; int* status2;
};

__device__ __forceinline__ float f_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f_sub(float a, float b) { return __fsub_rn(a, b); }

__global__ void cascade_assemble_kernel(const __grid_constant__ CascadeParams P) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.batch) return;
  const float* J = P.J + b * 3 * kDof;
  const float* Jd = P.Jd + b * 3 * kDof;
  const float* Mq = P.M + b * kDof * kDof;
  const float* q = P.q + b * kDof;
  const float* qd = P.qd + b * kDof;
  double* CE = P.CE0 + b * kN * kP0;
  // conditions [10][14] -> CE[i][c] = conditions[c][i]
  for (int i = 0; i < kN; i++)
    for (int c = 0; c < kP0; c++) {
      float v = 0.0f;
      if (c < 3) v = (i < kDof) ? J[c * kDof + i] : 0.0f;                        // [J 0]
      else if (i < kDof) v = Mq[(c - 3) * kDof + i];                             // [M ...
      else v = (i - kDof == c - 3) ? -1.0f : 0.0f;                               //  ... -I]
      CE[i * kP0 + c] = (double)v;
    }
  double* ce0 = P.ce00 + b * kP0;
  for (int c = 0; c < 3; c++) {
    float jdqd = 0.0f;
    for (int k = 0; k < kDof; k++) jdqd = f_add(jdqd, f_mul(Jd[c * kDof + k], qd[k]));
    float t = f_add(f_mul(P.kp_task, f_sub(P.pos_des[b * 3 + c], P.pos_cur[b * 3 + c])),
                    f_mul(P.kd_task, f_sub(P.vel_des[b * 3 + c], P.vel_cur[b * 3 + c])));
    t = f_sub(t, jdqd);
    t = f_add(t, P.acc_des[b * 3 + c]);
    t = f_sub(t, P.acc_cur[b * 3 + c]);
    ce0[c] = (double)(-t);
  }
  for (int c = 3; c < kP0; c++) ce0[c] = 0.0;
  const int j = P.joint_task;
  const float dterm = f_mul(P.kd_joint, f_sub(qd[j], qd[j]));
  P.goal2[b] = -f_add(f_mul(P.kp_joint, f_sub(P.q1_des[b], q[j])), dterm);
  // limits = [accP', tauP, -accN', -tauN]; std::min / std::max keep their first argument when the second is NaN
  double* ci0 = P.ci0 + b * kM;
  for (int i = 0; i < kDof; i++) {
    const double lp = log((double)f_sub(P.q_sup[i], q[i]));
    const double ln = -log((double)f_sub(q[i], P.q_inf[i]));
    const double ap = (double)P.acc_p[i], an = (double)P.acc_n[i];
    const float accp = (float)((lp < ap) ? lp : ap);
    const float accn = (float)((an < ln) ? ln : an);
    ci0[i] = (double)accp;
    ci0[kDof + i] = (double)P.tau_p[i];
    ci0[2 * kDof + i] = (double)(-accn);
    ci0[3 * kDof + i] = (double)(-P.tau_n[i]);
  }
}

// verdict of solveNextStep as the caller sees it (src/mgqp.cpp:717-742): false = the level failed
__device__ __forceinline__ bool cascade_step_ok(double f, int st) {
  const int s = st & 15;
  return !(f != f || f == __longlong_as_double(0x7ff0000000000000LL) || s == 2 || s == 3);
}

// One thread per scenario; the 14 x 10 working matrix lives in shared memory, element-major
// (W[e * blockDim + tid]: conflict-free).
__global__ void cascade_project_kernel(const __grid_constant__ CascadeParams P) {
  extern __shared__ float Wsm[];
  const int tid = threadIdx.x, T = blockDim.x;
  const long long b = (long long)blockIdx.x * T + tid;
  if (b >= P.batch) return;
  float* W = Wsm + tid;
#define WM(i, c) W[((i) * kP0 + (c)) * T]
  // W = conditions' (14 x 10): the columns are the level's condition rows, float32 as the reference holds them
  const double* CE = P.CE0 + b * kN * kP0;
  for (int e = 0; e < kN * kP0; e++) W[e * T] = (float)CE[e];
  // one-sided Jacobi: rotate column pairs until they are mutually orthogonal; then column k = sigma_k v_k
  for (int sweep = 0; sweep < 30; sweep++) {
    bool rotated = false;
    for (int p = 0; p < kP0 - 1; p++)
      for (int q = p + 1; q < kP0; q++) {
        float al = 0.0f, be = 0.0f, ga = 0.0f;
        for (int i = 0; i < kN; i++) {
          const float a = WM(i, p), c = WM(i, q);
          al = fmaf(a, a, al);
          be = fmaf(c, c, be);
          ga = fmaf(a, c, ga);
        }
        if (ga == 0.0f || fabsf(ga) <= 1e-7f * sqrtf(al * be)) continue;
        rotated = true;
        const float zeta = (be - al) / (2.0f * ga);
        const float t = copysignf(1.0f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.0f)));
        const float cs = rsqrtf(fmaf(t, t, 1.0f)), sn = cs * t;
        for (int i = 0; i < kN; i++) {
          const float a = WM(i, p), c = WM(i, q);
          WM(i, p) = cs * a - sn * c;
          WM(i, q) = sn * a + cs * c;
        }
      }
    if (!rotated) break;
  }
  // normalise the columns with sigma >= 1e-16 (src/mgqp.cpp:849-858); the others drop out of the projector
  for (int c = 0; c < kP0; c++) {
    float s2 = 0.0f;
    for (int i = 0; i < kN; i++) s2 = fmaf(WM(i, c), WM(i, c), s2);
    const float sig = sqrtf(s2);
    const float inv = (sig >= 1e-16f) ? 1.0f / sig : 0.0f;
    for (int i = 0; i < kN; i++) WM(i, c) *= inv;
  }
  // Z = I - V V'
  float* Z = P.Z + b * kN * kN;
  for (int i = 0; i < kN; i++)
    for (int jx = 0; jx < kN; jx++) {
      float acc = 0.0f;
      for (int c = 0; c < kP0; c++) acc = fmaf(WM(i, c), WM(jx, c), acc);
      Z[i * kN + jx] = ((i == jx) ? 1.0f : 0.0f) - acc;
    }
#undef WM
  // y = 0 + I u0 (:828); a level-0 failure ends the cascade with zeros (:820-824)
  const bool ok0 = cascade_step_ok(P.f0[b], P.st0[b]);
  float* y = P.y + b * kN;
  for (int i = 0; i < kN; i++) y[i] = ok0 ? (float)P.x0[b * kN + i] : 0.0f;
  // level 2 ("pb2", :796-802): A = conditions Z = row j of Z, a = goal - y_j, B = limitsMatrix Z = [-Z; Z]
  const int j = P.joint_task;
  double* CE2 = P.CE2 + b * kN;
  for (int i = 0; i < kN; i++) CE2[i] = (double)Z[j * kN + i];
  P.ce02[b] = (double)f_sub(P.goal2[b], y[j]);
  double* CI2 = P.CI2 + b * kN * kM;
  for (int i = 0; i < kN; i++)
    for (int c = 0; c < kN; c++) {
      const float z = Z[c * kN + i];
      CI2[i * kM + c] = (double)(-z);
      CI2[i * kM + kN + c] = (double)z;
    }
}

__global__ void cascade_finish_kernel(const __grid_constant__ CascadeParams P) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.batch) return;
  const bool ok0 = cascade_step_ok(P.f0[b], P.st0[b]);
  const bool ok2 = cascade_step_ok(P.f2[b], P.st2[b]);
  const float* Z = P.Z + b * kN * kN;
  const float* y = P.y + b * kN;
  float u[kN];
  for (int k = 0; k < kN; k++) u[k] = (float)P.x2[b * kN + k];
  for (int i = 0; i < kN; i++) {
    float zu = 0.0f;
    for (int k = 0; k < kN; k++) zu = f_add(zu, f_mul(Z[i * kN + k], u[k]));
    const float res = ok0 ? (ok2 ? f_add(y[i], zu) : y[i]) : 0.0f;
    P.tracking[b * kN + i] = res;
    if (i >= kDof) P.torques[b * kDof + (i - kDof)] = f_add(res, P.h[b * kDof + (i - kDof)]);
  }
  if (P.status0) P.status0[b] = P.st0[b];
  if (P.status2) P.status2[b] = ok0 ? P.st2[b] : -1;
}

}  // namespace mgqp
