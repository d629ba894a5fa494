"""Seeded synthetic QP batches of the shapes named in BASELINE.json / SURVEY.md section 8(d).

All arrays use the QuadProg++ orientation of the reference boundary
(include/QuadProgpp/QuadProg++.hh:15-25): ``G [B,n,n]``, ``g0 [B,n]``, ``CE [B,n,p]``, ``ce0 [B,p]``,
``CI [B,n,m]`` (one constraint per COLUMN), ``ci0 [B,m]``; the problem is
``min 1/2 x'Gx + g0'x  s.t.  CE'x + ce0 = 0,  CI'x + ci0 >= 0``.

Shapes:
  C2  n=7,  p=1,  m=42    "7-DoF mgqp QP (1 equality, 42 joint pos/vel/acc bounds)"   [headline]
  C3  n=14, p=7,  m=60    task-space QP with slacks
  C4  n=350,p=0,  m=2100  horizon-stacked MPC QP (7 DoF x 50 steps)
  C5  C2 shape, ~10 % made infeasible by a contradictory bound pair
  C5D C5 with a second equality (p=2): ~4 % of the QPs carry a duplicated equality column (about a quarter of them end up below the
      solver's dependence threshold: ~1 % status 3, the reference's runtime_error "Constraints are linearly dependent")
  S0  n=14, p=10, m=28    as-shipped level 0 (src/mgqp.cpp:1076-1143: G=I, g0=0, CI=[-I;+I])
  S2  n=14, p=1,  m=28    as-shipped level 2

The generator is numpy (PCG64) seeded with ``1234 + config id``; it is host-side data preparation,
not part of the solver.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SHAPES = {
    "C2": (7, 1, 42),
    "C3": (14, 7, 60),
    "C4": (350, 0, 2100),
    "C5": (7, 1, 42),
    "C5D": (7, 2, 42),
    "S0": (14, 10, 28),
    "S2": (14, 1, 28),
}
_CONFIG_ID = {"S0": 1, "S2": 1, "C2": 2, "C3": 3, "C4": 4, "C5": 5, "C5D": 6}


@dataclass
class QPBatch:
    """Host-side SoA batch in the reference's orientation."""

    name: str
    n: int
    p: int
    m: int
    G: np.ndarray
    g0: np.ndarray
    CE: np.ndarray
    ce0: np.ndarray
    CI: np.ndarray
    ci0: np.ndarray

    @property
    def batch(self) -> int:
        return self.G.shape[0]

    def fields(self):
        return (self.G, self.g0, self.CE, self.ce0, self.CI, self.ci0)

    def slice(self, lo: int, hi: int) -> "QPBatch":
        return QPBatch(self.name, self.n, self.p, self.m, *(a[lo:hi] for a in self.fields()))


def algorithmic_bytes(n: int, p: int, m: int) -> int:
    """Compulsory HBM traffic per QP: every input once, x, f and status once."""
    return 8 * (n * n + n + n * p + p + n * m + m) + 8 * (n + 1) + 4


def _spd(rng, B, n, scale=1.0):
    M = rng.standard_normal((B, n, n))
    return scale * np.einsum("bki,bkj->bij", M, M) / n + np.eye(n)[None]


def _box_columns(n_var, n_box, blocks):
    """[n_var, blocks*2*n_box]: per block, +e_i (lower bound) then -e_i (upper bound), i<n_box."""
    cols = np.zeros((n_var, blocks * 2 * n_box))
    for b in range(blocks):
        for i in range(n_box):
            cols[i, b * 2 * n_box + i] = 1.0
            cols[i, b * 2 * n_box + n_box + i] = -1.0
    return cols


def make_c2(B: int, seed: int | None = None, infeasible_frac: float = 0.0, name="C2") -> QPBatch:
    n, p, m = SHAPES["C2"]
    rng = np.random.default_rng(1234 + _CONFIG_ID[name] if seed is None else seed)
    G = _spd(rng, B, n)
    g0 = 3.0 * rng.standard_normal((B, n))
    CE = rng.standard_normal((B, n, p))
    ce0 = 0.3 * rng.standard_normal((B, p))
    CI = np.broadcast_to(_box_columns(n, n, 3), (B, n, m)).copy()
    block = np.repeat(np.arange(3), 2 * n)[None, :]
    ci0 = rng.uniform(0.2, 2.0, (B, m)) * (1.0 + block)
    if infeasible_frac > 0.0:
        bad = rng.random(B) < infeasible_frac
        j = rng.integers(0, n, B)
        idx = np.nonzero(bad)[0]
        # x_j >= a (col j of block 0, ci0=-a) and x_j <= b (col n+j, ci0=b) with a > b
        a = rng.uniform(1.0, 2.0, idx.size)
        ci0[idx, j[idx]] = -a
        ci0[idx, n + j[idx]] = a - rng.uniform(0.1, 0.5, idx.size)
    return QPBatch(name, n, p, m, G, g0, CE, ce0, CI, ci0)


def make_c5(B: int, seed: int | None = None) -> QPBatch:
    return make_c2(B, seed, infeasible_frac=0.10, name="C5")


def make_c5d(B: int, seed: int | None = None, dependent_frac: float = 0.04) -> QPBatch:
    """~10 % infeasible and ~1 % answered "linearly dependent" (duplicated equality columns, p = 2;
    an exactly duplicated column falls below the solver's eps * R_norm threshold only about one time in four)."""
    b = make_c2(B, 1234 + _CONFIG_ID["C5D"] if seed is None else seed, infeasible_frac=0.10, name="C5")
    rng = np.random.default_rng((1234 + 77 * _CONFIG_ID["C5D"]) if seed is None else seed + 77)
    CE = np.concatenate([b.CE, rng.standard_normal((B, b.n, 1))], axis=2)
    ce0 = np.concatenate([b.ce0, 0.3 * rng.standard_normal((B, 1))], axis=1)
    dup = rng.random(B) < dependent_frac
    CE[dup, :, 1] = CE[dup, :, 0]
    return QPBatch("C5D", b.n, 2, b.m, b.G, b.g0, CE, ce0, b.CI, b.ci0)


def make_c4_device(B: int, device, seed: int = 1234 + 4, dof: int = 7, steps: int = 50):
    """The C4 workload generated ON THE DEVICE (torch tensors in field order): at BASELINE.json's 10k QPs the
    batch is 69 GB, which no host generator fills in reasonable time.  Same construction as make_c4 (dense SPD
    G = H'H/n + I, integrator-matrix bounds with per-QP limits), torch's generator instead of numpy's."""
    import torch
    n = dof * steps
    m = 6 * n
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    G = torch.empty((B, n, n), dtype=torch.float64, device=device)
    eye = torch.eye(n, dtype=torch.float64, device=device)
    for lo in range(0, B, 256):  # chunked: the n*n temporaries stay small
        hi = min(B, lo + 256)
        H = torch.randn((hi - lo, n, n), dtype=torch.float64, device=device, generator=gen)
        G[lo:hi] = torch.einsum("bki,bkj->bij", H, H) / n + eye
    g0 = 2.0 * torch.randn((B, n), dtype=torch.float64, device=device, generator=gen)
    rows = torch.from_numpy(_c4_constraint_rows(dof, steps)).to(device)            # [m, n], row = constraint
    CI = rows.t().contiguous()[None].expand(B, n, m).contiguous()
    lim = torch.cat([torch.full((2 * n,), 1.0), torch.full((2 * n,), 0.5), torch.full((2 * n,), 0.3)]).to(device).double()
    ci0 = lim[None, :] * (0.5 + torch.rand((B, m), dtype=torch.float64, device=device, generator=gen))
    CE = torch.zeros((B, n, 0), dtype=torch.float64, device=device)
    ce0 = torch.zeros((B, 0), dtype=torch.float64, device=device)
    return [G, g0, CE, ce0, CI, ci0]


def _c4_constraint_rows(dof: int, steps: int) -> np.ndarray:
    n = dof * steps
    dt = 0.05
    I = np.eye(dof)
    S = np.zeros((n, n))  # vel_k = sum_{j<=k} dt a_j
    P = np.zeros((n, n))  # pos_k = sum_{j<=k} (k-j+0.5) dt^2 a_j
    for k in range(steps):
        for j in range(k + 1):
            S[k * dof:(k + 1) * dof, j * dof:(j + 1) * dof] = dt * I
            P[k * dof:(k + 1) * dof, j * dof:(j + 1) * dof] = (k - j + 0.5) * dt * dt * I
    return np.concatenate([np.eye(n), -np.eye(n), S, -S, P, -P], axis=0)


def make_c3(B: int, seed: int | None = None) -> QPBatch:
    n, p, m = SHAPES["C3"]
    rng = np.random.default_rng(1234 + _CONFIG_ID["C3"] if seed is None else seed)
    G = np.zeros((B, n, n))
    G[:, np.arange(7), np.arange(7)] = 1e-2
    G[:, np.arange(7, 14), np.arange(7, 14)] = 1.0
    N = rng.standard_normal((B, n, n))
    G += 1e-3 * np.einsum("bki,bkj->bij", N, N) / n
    g0 = 0.1 * rng.standard_normal((B, n))
    # equalities: 6 Cartesian rows [J_task  I6  0], one joint row [e_j  0  1]
    CE = np.zeros((B, n, p))
    CE[:, :7, :6] = 0.5 * rng.standard_normal((B, 7, 6))
    CE[:, 7 + np.arange(6), np.arange(6)] = 1.0
    jsel = rng.integers(0, 7, B)
    CE[np.arange(B), jsel, 6] = 1.0
    CE[:, 13, 6] = 1.0
    ce0 = rng.standard_normal((B, p))
    CI = np.zeros((B, n, m))
    CI[:, :, :42] = _box_columns(n, 7, 3)[None]
    block = np.repeat(np.arange(3), 14)[None, :]
    ci0 = np.empty((B, m))
    ci0[:, :42] = rng.uniform(0.5, 3.0, (B, 42)) * (1.0 + block)
    Mq = _spd(rng, B, 7)  # joint-space inertia; torque bounds |M qdd + h| <= 100
    h = 10.0 * rng.standard_normal((B, 7))
    CI[:, :7, 42:49] = np.swapaxes(Mq, 1, 2)
    CI[:, :7, 49:56] = -np.swapaxes(Mq, 1, 2)
    ci0[:, 42:49] = 100.0 + h
    ci0[:, 49:56] = 100.0 - h
    CI[:, :, 56:60] = rng.standard_normal((B, n, 4))
    ci0[:, 56:60] = rng.uniform(0.5, 2.0, (B, 4))
    return QPBatch("C3", n, p, m, G, g0, CE, ce0, CI, ci0)


def make_c4(B: int, seed: int | None = None, dof: int = 7, steps: int = 50) -> QPBatch:
    n = dof * steps
    m = 6 * n
    rng = np.random.default_rng(1234 + _CONFIG_ID["C4"] if seed is None else seed)
    G = _spd(rng, B, n)
    g0 = 2.0 * rng.standard_normal((B, n))
    rows = _c4_constraint_rows(dof, steps)  # [m, n], row = constraint
    CI = np.broadcast_to(rows.T[None], (B, n, m)).copy()
    lim = np.concatenate([np.full(2 * n, 1.0), np.full(2 * n, 0.5), np.full(2 * n, 0.3)])
    ci0 = lim[None, :] * rng.uniform(0.5, 1.5, (B, m))
    CE = np.zeros((B, n, 0))
    ce0 = np.zeros((B, 0))
    return QPBatch("C4", n, 0, m, G, g0, CE, ce0, CI, ci0)


def _f32(a):
    """Values as the reference adapter passes them: float32 widened to double (src/mgqp.cpp:700-706)."""
    return a.astype(np.float32).astype(np.float64)


def make_shipped(B: int, level: int = 0, seed: int | None = None) -> QPBatch:
    """As-shipped shapes S0 (level 0) / S2 (level 2): n=14=[qdd; tau], G=I, g0=0, CI=[-I;+I]'."""
    dof = 7
    n = 2 * dof
    p = 10 if level == 0 else 1
    m = 4 * dof
    name = "S0" if level == 0 else "S2"
    rng = np.random.default_rng((1234 + 1 + level) if seed is None else seed)
    G = np.broadcast_to(np.eye(n), (B, n, n)).copy()
    g0 = np.zeros((B, n))
    CE = np.zeros((B, n, p))
    if level == 0:
        CE[:, :dof, :3] = 0.3 * rng.standard_normal((B, dof, 3))  # [J 0] translation task rows
        Mq = _spd(rng, B, dof)
        CE[:, :dof, 3:10] = np.swapaxes(Mq, 1, 2)  # [M -I] dynamics rows (src/mgqp.cpp:1138-1143)
        CE[:, dof + np.arange(dof), 3 + np.arange(dof)] = -1.0
        ce0 = np.concatenate([rng.standard_normal((B, 3)), np.zeros((B, dof))], axis=1)
    else:
        CE[:, 0, 0] = 1.0  # joint-1 position task row e_j
        ce0 = rng.standard_normal((B, 1))
    CI = np.zeros((B, n, m))
    CI[:, np.arange(n), np.arange(n)] = -1.0
    CI[:, np.arange(n), n + np.arange(n)] = 1.0
    acc = 5.0 * np.where(rng.random((B, dof)) < 0.3, rng.uniform(0.05, 1.0, (B, dof)), 1.0)
    tau = np.full((B, dof), 100.0)
    ci0 = np.concatenate([acc, tau, acc, tau], axis=1)
    return QPBatch(name, n, p, m, _f32(G), _f32(g0), _f32(CE), _f32(ce0), _f32(CI), _f32(ci0))


# Gains (src/mgqp.cpp:104-110) and limits (ops/mgqp.ops:132-134, :187-189) of the 7-DoF deployment as shipped
SHIPPED_CONFIG = dict(
    kp_task=np.float32(100.0), kd_task=np.float32(25.0), kp_joint=np.float32(200.0), kd_joint=np.float32(100.0),
    acc_p=np.full(7, 5.0, np.float32), acc_n=np.full(7, -5.0, np.float32),
    tau_p=np.full(7, 100.0, np.float32), tau_n=np.full(7, -100.0, np.float32),
    q_sup=np.array([0.8, 1.5, 2.5, 1.5, 3.0, 1.5, 3.0], np.float32),
    q_inf=-np.array([0.8, 1.5, 2.5, 1.5, 3.0, 1.5, 3.0], np.float32),
    joint_task=0,
)
# the same wiring with the wide joint range the deployment script uses for its other robot (ops/mgqp.ops:190-195):
# with the shipped +-0.8 rad on joint 1 the log-barrier bounds always contradict each other and every level ends in
# solveNextStep's retry ("the problem seems never feasible", src/mgqp.cpp:716); here most levels are feasible
WIDE_CONFIG = dict(SHIPPED_CONFIG, q_sup=np.full(7, 3.1415927, np.float32), q_inf=np.full(7, -3.1415927, np.float32))
SCENARIO_FIELDS = ("q", "qd", "M", "h", "J", "Jd", "pos_des", "pos_cur", "vel_des", "vel_cur", "acc_des", "acc_cur",
                   "q1_des")


def make_scenarios(B: int, seed: int = 4321, config=None, scale: float = 1.0):
    """Seeded robot states around the shipped set-up (float32): joint angles inside the limits (some close to
    them, so that the log barrier bites), SPD inertia, a translation Jacobian, targets a few cm away.
    `scale` shrinks tracking errors and velocities: with small ones (0.01) and WIDE_CONFIG the demanded
    accelerations fit the box limits and the levels are feasible, otherwise most levels end in the retry."""
    rng = np.random.default_rng(seed)
    cfg = SHIPPED_CONFIG if config is None else config
    f32 = np.float32
    span = cfg["q_sup"] - cfg["q_inf"]
    u = rng.uniform(0.02, 0.98, (B, 7))
    near = rng.random((B, 7)) < 0.1
    u = np.where(near, rng.choice([0.0005, 0.9995], (B, 7)), u)
    q = (cfg["q_inf"] + span * u).astype(f32)
    H = rng.standard_normal((B, 7, 7))
    Mq = (0.3 * np.einsum("bki,bkj->bij", H, H) / 7 + np.eye(7) * rng.uniform(0.5, 3.0, (B, 1, 1))).astype(f32)
    s = dict(
        q=q, qd=(0.5 * scale * rng.standard_normal((B, 7))).astype(f32), M=Mq,
        h=(5.0 * rng.standard_normal((B, 7))).astype(f32),
        J=(0.4 * rng.standard_normal((B, 3, 7))).astype(f32), Jd=(0.2 * rng.standard_normal((B, 3, 7))).astype(f32),
        pos_des=rng.uniform(-0.5, 0.5, (B, 3)).astype(f32), vel_des=(0.1 * scale * rng.standard_normal((B, 3))).astype(f32),
        acc_des=(0.1 * scale * rng.standard_normal((B, 3))).astype(f32),
        vel_cur=(0.1 * scale * rng.standard_normal((B, 3))).astype(f32),
        acc_cur=(0.1 * scale * rng.standard_normal((B, 3))).astype(f32),
    )
    s["q1_des"] = (q[:, cfg["joint_task"]] + 0.3 * scale * rng.standard_normal(B)).astype(f32)
    s["pos_cur"] = (s["pos_des"] + 0.05 * scale * rng.standard_normal((B, 3))).astype(f32)
    return s


def make_random(B: int, n: int, p: int, m: int, seed: int = 0, box: bool = False) -> QPBatch:
    """Generic dense random QPs of any shape (shape-generic parity tests)."""
    rng = np.random.default_rng(seed)
    G = _spd(rng, B, n)
    g0 = 2.0 * rng.standard_normal((B, n))
    CE = rng.standard_normal((B, n, p))
    ce0 = 0.3 * rng.standard_normal((B, p))
    if box and m >= 2 * n:
        CI = np.zeros((B, n, m))
        CI[:, :, : 2 * n] = _box_columns(n, n, 1)[None]
        CI[:, :, 2 * n:] = rng.standard_normal((B, n, m - 2 * n))
    else:
        CI = rng.standard_normal((B, n, m))
    ci0 = rng.uniform(0.1, 1.5, (B, m))
    return QPBatch(f"R{n}_{p}_{m}", n, p, m, G, g0, CE, ce0, CI, ci0)


def make(name: str, B: int, seed: int | None = None) -> QPBatch:
    if name == "C2":
        return make_c2(B, seed)
    if name == "C3":
        return make_c3(B, seed)
    if name == "C4":
        return make_c4(B, seed)
    if name == "C5":
        return make_c5(B, seed)
    if name == "C5D":
        return make_c5d(B, seed)
    if name == "S0":
        return make_shipped(B, 0, seed)
    if name == "S2":
        return make_shipped(B, 2, seed)
    raise KeyError(name)
