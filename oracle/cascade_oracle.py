"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the two callers on either side of the solver, for the wiring
the reference ships (ops/mgqp.ops:233-240, :256, :289-292; 7 DoF):

* ``assemble``  = the QP assembly of ``updateHook`` (src/mgqp.cpp:886-1143): level 0 gets the translational
  task of link 7, rows ``[J 0]`` with goal ``-(Kp (pd - pc) + Kd (vd - vc) - Jdot qdot + ad - ac)`` (:1037-1057),
  and the dynamics rows ``[M -I]`` (:1138-1143); level 2 gets the joint-1 position task, row ``e_1`` (:1058-1070);
  the inequality block is ``limitsMatrix = [-I; +I]`` (:1111-1112) with the log-barrier shaped acceleration limits
  (:1119-1123).  float32, sums in ascending index order, one rounding per operation.
* ``cascade``   = ``solveNextHierarchy`` (:751-869) over those levels: per level ``A = conditions Z``,
  ``a = goal - conditions y``, ``B = Bcumul Z``, one ``solveNextStep`` (:655-749: float32 -> fp64 copies, G = I,
  g0 = 0, the infeasible-retry policy :717-735), ``y += Z u``, ``Z = I - V diag(sigma >= 1e-16) V'`` (:836-862).

PARITY UNPINNED against the reference for these two callers: they are Eigen float32 code (``JacobiSVD``, Eigen's own
product kernels) and Eigen is not in the reference tree, so neither the summation order of its matrix products nor
its singular vectors can be reproduced here.  This restatement fixes an order (ascending index) and takes the
singular vectors from LAPACK (``numpy.linalg.svd`` in float32); the CUDA path is held to it bit for bit where the
arithmetic is exact or elementwise (the level-0 QP data) and within the tolerances written in
tests/test_cascade.py where an SVD is involved.  The QP solves themselves go through the pinned solver oracle.
"""
from __future__ import annotations

import numpy as np

DOF = 7
N = 2 * DOF          # variables [qdd; tau] (src/mgqp.cpp:759-764, :888)
M_INEQ = 4 * DOF     # src/mgqp.cpp:1076
F32 = np.float32

# gains (src/mgqp.cpp:104-110) and limits (ops/mgqp.ops:132-134, :187-189) as shipped
DEFAULT_CONFIG = dict(
    kp_task=F32(100.0), kd_task=F32(25.0), kp_joint=F32(200.0), kd_joint=F32(100.0),
    acc_p=np.full(DOF, 5.0, F32), acc_n=np.full(DOF, -5.0, F32),
    tau_p=np.full(DOF, 100.0, F32), tau_n=np.full(DOF, -100.0, F32),
    q_sup=np.array([0.8, 1.5, 2.5, 1.5, 3.0, 1.5, 3.0], F32), q_inf=-np.array([0.8, 1.5, 2.5, 1.5, 3.0, 1.5, 3.0], F32),
    joint_task=0,    # "in_desiredJointSpacePosition_1" (0-based joint index), priority level 2
)

STATE_FIELDS = ("q", "qd", "M", "h", "J", "Jd", "pos_des", "pos_cur", "vel_des", "vel_cur", "acc_des", "acc_cur",
                "q1_des")


def _dot_f32(A, x):
    """row-by-row float32 dot products, ascending index, one rounding per operation"""
    acc = np.zeros(A.shape[:-1], F32)
    for k in range(A.shape[-1]):
        acc = (acc + (A[..., k] * x[..., k]).astype(F32)).astype(F32)
    return acc


def assemble(s, cfg=DEFAULT_CONFIG):
    """-> dict(cond0 [B,10,14], goal0 [B,10], cond2 [B,1,14], goal2 [B,1], limits_matrix [28,14], limits [B,28])"""
    B = s["q"].shape[0]
    kp, kd = cfg["kp_task"], cfg["kd_task"]
    cond0 = np.zeros((B, 3 + DOF, N), F32)
    cond0[:, :3, :DOF] = s["J"]                                             # A = [J 0]           :1042-1047
    cond0[:, 3:, :DOF] = s["M"]                                             # [M -I]              :1138-1141
    cond0[:, 3 + np.arange(DOF), DOF + np.arange(DOF)] = F32(-1.0)
    jdqd = _dot_f32(s["Jd"], s["qd"][:, None, :])
    t = ((kp * (s["pos_des"] - s["pos_cur"]).astype(F32)).astype(F32)
         + (kd * (s["vel_des"] - s["vel_cur"]).astype(F32)).astype(F32)).astype(F32)
    t = (t - jdqd).astype(F32)
    t = (t + s["acc_des"]).astype(F32)
    t = (t - s["acc_cur"]).astype(F32)
    goal0 = np.zeros((B, 3 + DOF), F32)
    goal0[:, :3] = -t                                                       # :1049-1051
    j = cfg["joint_task"]
    cond2 = np.zeros((B, 1, N), F32)
    cond2[:, 0, j] = F32(1.0)                                               # :1064
    # desired joint velocity defaults to the current one (:1011-1014): the D term is kd * 0
    dterm = (cfg["kd_joint"] * (s["qd"][:, j] - s["qd"][:, j]).astype(F32)).astype(F32)
    goal2 = (-((cfg["kp_joint"] * (s["q1_des"] - s["q"][:, j]).astype(F32)).astype(F32) + dterm).astype(F32))[:, None]
    LM = np.zeros((M_INEQ, N), F32)
    LM[np.arange(N), np.arange(N)] = F32(-1.0)                              # :1111
    LM[N + np.arange(N), np.arange(N)] = F32(1.0)                           # :1112
    # log barrier (:1119-1123): min((double) accP, log(qsup - q)), max((double) accN, -log(q - qinf)) in double,
    # stored back into float32; std::min / std::max keep their FIRST argument when the second is NaN
    with np.errstate(invalid="ignore", divide="ignore"):
        lp = np.log((cfg["q_sup"][None] - s["q"]).astype(F32).astype(np.float64))
        ln = -np.log((s["q"] - cfg["q_inf"][None]).astype(F32).astype(np.float64))
    ap = np.broadcast_to(cfg["acc_p"].astype(np.float64), lp.shape)
    an = np.broadcast_to(cfg["acc_n"].astype(np.float64), ln.shape)
    accp = np.where(lp < ap, lp, ap).astype(F32)
    accn = np.where(an < ln, ln, an).astype(F32)
    limits = np.concatenate([accp, np.broadcast_to(cfg["tau_p"], accp.shape), -accn,
                             np.broadcast_to(-cfg["tau_n"], accp.shape)], axis=1).astype(F32)  # :1125-1128
    return dict(cond0=cond0, goal0=goal0, cond2=cond2, goal2=goal2, limits_matrix=LM, limits=limits)


def solve_next_step(restatement, A, a, Bm, b, threads=8):
    """solveNextStep (src/mgqp.cpp:655-749) for a batch: G = I, g0 = 0, CE = A', CI = B' in fp64, the answer
    truncated to float32; NaN / +inf -> second solve without inequalities on the factor-overwritten G (= I here);
    -> (u float32 [B,n], ok bool [B], status int [B] as MGQP_FLAG_RETRY_WITHOUT_CI reports it)."""
    B_, p, n = A.shape
    m = Bm.shape[1]
    G = np.broadcast_to(np.eye(n), (B_, n, n)).copy()
    g0 = np.zeros((B_, n))
    CE = np.ascontiguousarray(np.swapaxes(A.astype(np.float64), 1, 2))
    CI = np.ascontiguousarray(np.swapaxes(Bm.astype(np.float64), 1, 2))
    r = restatement.solve_batch(n, p, m, G, g0, CE, a.astype(np.float64), CI, b.astype(np.float64),
                                max_iter=100 + 10 * (p + m), threads=threads)
    x, f, st = r["x"].copy(), r["f"].copy(), r["status"].copy()
    again = np.nonzero(((f != f) | np.isposinf(f)) & (st != 2) & (st != 3))[0]
    if again.size:
        r2 = restatement.solve_batch(n, p, 0, G[again], g0[again], CE[again], a[again].astype(np.float64),
                                     np.zeros((again.size, n, 0)), np.zeros((again.size, 0)),
                                     max_iter=100 + 10 * p, threads=threads)
        f2, st2 = r2["f"], r2["status"]
        lost = ((f2 != f2) | np.isposinf(f2)) & (st2 != 2) & (st2 != 3)
        x[again] = np.where(lost[:, None], 0.0, r2["x"])
        f[again] = f2
        st[again] = np.where(lost, 1, st2) | 16
    ok = ~(((f != f) | np.isposinf(f)) | ((st & 15) == 2) | ((st & 15) == 3))
    return x.astype(F32), ok, st


def null_space_projector(Acumul):
    """Z = I - V diag(sigma >= 1e-16) V' (src/mgqp.cpp:836-862), float32, V from LAPACK instead of Eigen's JacobiSVD."""
    B_, r, n = Acumul.shape
    _, sig, Vt = np.linalg.svd(Acumul.astype(F32), full_matrices=False)
    keep = (sig >= F32(1e-16)).astype(F32)
    V = np.swapaxes(Vt, 1, 2).astype(F32)                   # [B, n, min(r, n)]
    VA = (V * keep[:, None, :]).astype(F32)
    Z = (np.eye(n, dtype=F32)[None] - np.einsum("bik,bjk->bij", VA, V).astype(F32)).astype(F32)
    return Z, sig


def cascade(restatement, s, cfg=DEFAULT_CONFIG, threads=8, Z_override=None, limits_override=None):
    """-> dict(tracking [B,14] float32, torques [B,7] float32, status0, status2, qp0 / qp2 (the QP data), Z)"""
    asm = assemble(s, cfg)
    if limits_override is not None:      # tests: the device's own log-barrier limits (device log vs libm log)
        asm["limits"] = limits_override
    B_ = s["q"].shape[0]
    LMb = np.broadcast_to(asm["limits_matrix"], (B_, M_INEQ, N))
    # level 0: Z = I, y = 0: A = conditions, a = goal, B = limitsMatrix, b = limits       (case "pb1", :791-795)
    u0, ok0, st0 = solve_next_step(restatement, asm["cond0"], asm["goal0"], LMb, asm["limits"], threads)
    res0 = u0.copy()                                                       # res = 0 + I u   (:828)
    Z = Z_override if Z_override is not None else null_space_projector(asm["cond0"])[0]
    # level 1 is empty (:785-788); level 2: "pb2" (:796-802) with Bcumul = level 0's limits
    j = cfg["joint_task"]
    A2 = Z[:, j:j + 1, :].copy()                                            # conditions Z with conditions = e_j
    a2 = (asm["goal2"][:, 0] - res0[:, j]).astype(F32)[:, None]            # goal - conditions y
    B2 = np.concatenate([-Z, Z], axis=1).astype(F32)                        # limitsMatrix Z
    u2, ok2, st2 = solve_next_step(restatement, A2, a2, B2, asm["limits"], threads)
    zu = np.zeros((B_, N), F32)
    for k in range(N):
        zu = (zu + (Z[:, :, k] * u2[:, k:k + 1]).astype(F32)).astype(F32)
    res2 = (res0 + zu).astype(F32)
    # a level that fails returns the levels solved so far (:820-824); level 0 failing returns zeros
    tracking = np.where(ok0[:, None], np.where(ok2[:, None], res2, res0), F32(0.0)).astype(F32)
    torques = (tracking[:, DOF:] + s["h"]).astype(F32)                     # :1152
    return dict(tracking=tracking, torques=torques, status0=st0, status2=np.where(ok0, st2, -1), ok0=ok0, ok2=ok2,
                qp0=dict(CE=np.swapaxes(asm["cond0"], 1, 2), ce0=asm["goal0"], ci0=asm["limits"]),
                qp2=dict(CE=np.swapaxes(A2, 1, 2), ce0=a2, CI=np.swapaxes(B2, 1, 2), ci0=asm["limits"]), Z=Z, res0=res0)
