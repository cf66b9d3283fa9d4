"""Design evidence (test infrastructure; imports oracle/): accuracy of three ways to solve the SQP sub-problem

    min 1/2 dx^T G dx + c^T dx   s.t.  A dx + b = 0                      (mini_opt qp.cc:679-729)

for the cart-pole shooting problem, against an extended-precision (np.longdouble) dense KKT solve:
  * `nullspace`  the oracle's restatement of QPNullSpaceSolver (column-pivoted QR + Cholesky, dense);
  * `condensed`  forward condensation to the W controls + Schur complement on the terminal rows
                 (the round-1 kernel's method: open-loop sensitivity products);
  * `riccati`    segment-wise Riccati recursion (state = node step + previous control, control block = the S knots of
                 one shooting segment) with the hard terminal rows handled by multipliers on closed-loop responses.

    python tests/prototypes/qp_riccati_proto.py [--W 160] [--S 10] [--n 8] [--lam 0.0]
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cart_pole_mpc_b200 import synthetic  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def wrap(a):
    return O.mod_pi(a)


def linearise(dyn, xs, u, x_cur, set_point, W, S, dt):
    """Segment transition blocks at (xs, u): Phi[s] (4x4), Gam[s] (4xS), e[s] (4) — optimization.cc:99-160."""
    nseg = W // S
    Phi, Gam, ev = [], [], []
    for s in range(nseg):
        x = xs[s].copy()
        P = np.eye(4)
        G = np.zeros((4, S))
        for i in range(S):
            x, A, B = O.rk4_jac(dyn, x, u[s * S + i], dt)
            P = A @ P
            G = A @ G
            G[:, i] = B
        x[1] = wrap(x[1])
        e = x - xs[s + 1]
        e[1] = wrap(e[1])
        Phi.append(P)
        Gam.append(G)
        ev.append(e)
    return Phi, Gam, ev


def dense_qp(Phi, Gam, ev, xs, u, x_cur, set_point, u_prev, wts, lam, W, S):
    """G, c, A, b exactly as BuildProblem lays them out (optimization.cc:194-301)."""
    w_u, w_r, wt = wts[0], wts[1], wts[2:]
    nseg = W // S
    NS = nseg + 1
    N = 4 * NS + W
    target = np.array([set_point, math.pi / 2, 0.0, 0.0])
    G = np.zeros((N, N))
    c = np.zeros(N)
    rows, rhs = [], []
    for s in range(nseg):
        for r in range(4):
            a = np.zeros(N)
            a[4 * s:4 * s + 4] = Phi[s][r]
            a[4 * (s + 1) + r] = -1.0
            a[4 * NS + s * S:4 * NS + (s + 1) * S] = Gam[s][r]
            rows.append(a)
            rhs.append(ev[s][r])
    for r in range(4):
        a = np.zeros(N)
        a[r] = 1.0
        d = xs[0][r] - x_cur[r]
        if r == 1:
            d = wrap(d)
        rows.append(a)
        rhs.append(d)
    for r in range(4):
        d = xs[NS - 1][r] - target[r]
        if r == 1:
            d = wrap(d)
        j = 4 * (NS - 1) + r
        if wt[r] >= 0:
            G[j, j] += wt[r] ** 2
            c[j] += wt[r] ** 2 * d
        else:
            a = np.zeros(N)
            a[j] = 1.0
            rows.append(a)
            rhs.append(d)
    o = 4 * NS
    for k in range(W - 1):
        r = (u[k] - u[k + 1]) * w_r
        G[o + k, o + k] += w_r ** 2
        G[o + k + 1, o + k + 1] += w_r ** 2
        G[o + k + 1, o + k] -= w_r ** 2
        G[o + k, o + k + 1] -= w_r ** 2
        c[o + k] += w_r * r
        c[o + k + 1] -= w_r * r
    G[o, o] += w_r ** 2
    c[o] += w_r ** 2 * (u[0] - u_prev)
    for k in range(W):
        G[o + k, o + k] += w_u ** 2
        c[o + k] += w_u ** 2 * u[k]
    G += lam * np.eye(N)
    return G, c, np.array(rows), np.array(rhs)


def kkt_longdouble(G, c, A, b):
    N, K = G.shape[0], A.shape[0]
    M = np.zeros((N + K, N + K + 1), dtype=np.longdouble)
    M[:N, :N] = G
    M[:N, N:N + K] = A.T
    M[N:, :N] = A
    M[:N, -1] = -c
    M[N:, -1] = -b
    n = N + K
    for i in range(n):
        p = i + int(np.argmax(np.abs(M[i:, i])))
        if p != i:
            M[[i, p]] = M[[p, i]]
        M[i + 1:] -= np.outer(M[i + 1:, i] / M[i, i], M[i])
    x = np.zeros(n, dtype=np.longdouble)
    for i in range(n - 1, -1, -1):
        x[i] = (M[i, -1] - M[i, i + 1:n] @ x[i + 1:]) / M[i, i]
    return x[:N].astype(np.float64)


def condensed(Phi, Gam, ev, xs, u, x_cur, set_point, u_prev, wts, lam, W, S):
    """Round-1 kernel method, lambda == 0 only: eliminate states forward, tridiagonal T, Schur on the terminal rows."""
    assert lam == 0.0
    w_u, w_r, wt = wts[0], wts[1], wts[2:]
    nseg = W // S
    NS = nseg + 1
    target = np.array([set_point, math.pi / 2, 0.0, 0.0])
    T = np.zeros((W, W))
    g = np.zeros(W)
    for k in range(W):
        T[k, k] = w_u ** 2 + w_r ** 2 * (2 if k + 1 < W else 1)
        if k + 1 < W:
            T[k, k + 1] = T[k + 1, k] = -w_r ** 2
        g[k] = w_u ** 2 * u[k]
        if k + 1 < W:
            g[k] += w_r ** 2 * (u[k] - u[k + 1])
        if k > 0:
            g[k] -= w_r ** 2 * (u[k - 1] - u[k])
        if k == 0:
            g[k] += w_r ** 2 * (u[k] - u_prev)
    d0 = xs[0] - x_cur
    d0[1] = wrap(d0[1])
    m = -d0
    for s in range(nseg):
        m = Phi[s] @ m + ev[s]
    U = np.zeros((4, W))
    for s in range(nseg):
        V = Gam[s].copy()
        for t in range(s + 1, nseg):
            V = Phi[t] @ V
        U[:, s * S:(s + 1) * S] = V
    dT = xs[NS - 1] - target
    dT[1] = wrap(dT[1])
    inv_d = np.zeros(4)
    off = np.zeros(4)
    for i in range(4):
        if wt[i] < 0:
            inv_d[i] = 0.0
            off[i] = m[i] + dT[i]
        else:
            inv_d[i] = 1.0 / wt[i] ** 2
            off[i] = m[i] + dT[i]
    Tinv_g = np.linalg.solve(T, g)
    Y = np.linalg.solve(T, U.T)
    Gm = U @ Y + np.diag(inv_d)
    sig = np.linalg.solve(Gm, off - U @ Tinv_g)
    du = -Tinv_g - Y @ sig
    dxs = [-d0]
    for s in range(nseg):
        dxs.append(Phi[s] @ dxs[-1] + Gam[s] @ du[s * S:(s + 1) * S] + ev[s])
    return np.concatenate([np.concatenate(dxs), du])


def riccati(Phi, Gam, ev, xs, u, x_cur, set_point, u_prev, wts, lam, W, S):
    """Segment-wise Riccati, all in float64, with the operation structure planned for the kernel."""
    w_u, w_r, wt = wts[0], wts[1], wts[2:]
    nseg = W // S
    NS = nseg + 1
    target = np.array([set_point, math.pi / 2, 0.0, 0.0])
    wu2, wr2 = w_u ** 2, w_r ** 2
    g = np.zeros(W)
    for k in range(W):
        g[k] = wu2 * u[k]
        if k + 1 < W:
            g[k] += wr2 * (u[k] - u[k + 1])
        if k > 0:
            g[k] -= wr2 * (u[k - 1] - u[k])
        if k == 0:
            g[k] += wr2 * (u[k] - u_prev)
    dT = xs[NS - 1] - target
    dT[1] = wrap(dT[1])
    hard = [i for i in range(4) if wt[i] < 0]
    nr = 1 + len(hard)
    # terminal value function over xi = (dx, nu)
    P = np.zeros((5, 5))
    p = np.zeros((5, nr))
    for i in range(4):
        P[i, i] = lam
        if wt[i] >= 0:
            P[i, i] += wt[i] ** 2
            p[i, 0] = wt[i] ** 2 * dT[i]
    for j, i in enumerate(hard):
        p[i, 1 + j] = 1.0
    Ls, Hx, hh = [None] * nseg, [None] * nseg, [None] * nseg
    M = np.zeros((nr, nr))
    for s in range(nseg - 1, -1, -1):
        F = np.zeros((5, 5))
        F[:4, :4] = Phi[s]
        Gb = np.zeros((5, S))
        Gb[:4] = Gam[s]
        Gb[4, S - 1] = 1.0
        f = np.zeros(5)
        f[:4] = ev[s]
        Ts = np.zeros((S, S))
        for i in range(S):
            k = s * S + i
            Ts[i, i] = wu2 + lam + wr2 * (2 if k + 1 < W else 1)
            if i + 1 < S:
                Ts[i, i + 1] = Ts[i + 1, i] = -wr2
        PG = P @ Gb
        Hvv = Ts + Gb.T @ PG
        Hvx = PG.T @ F
        if s > 0:
            Hvx[0, 4] += -wr2
        q = p.copy()
        q[:, 0] += P @ f
        hv = Gb.T @ q
        hv[:, 0] += g[s * S:(s + 1) * S]
        hx = F.T @ q
        Hxx = F.T @ P @ F
        for i in range(4):
            Hxx[i, i] += lam if s > 0 else 0.0  # dx_0 is fixed: its lambda term is a constant
        L = np.linalg.cholesky(Hvv)
        Hh = np.linalg.solve(L, Hvx)   # S x 5  (forward substitution)
        h = np.linalg.solve(L, hv)     # S x nr
        P = Hxx - Hh.T @ Hh
        P = 0.5 * (P + P.T)
        p = hx - Hh.T @ h
        M += h.T @ h
        Ls[s], Hx[s], hh[s] = L, Hh, h
    d0 = xs[0] - x_cur
    d0[1] = wrap(d0[1])
    # forward pass for every right-hand side
    traj_x = np.zeros((nr, NS, 4))
    traj_v = np.zeros((nr, W))
    for r in range(nr):
        xi = np.zeros(5)
        if r == 0:
            xi[:4] = -d0
        traj_x[r, 0] = xi[:4]
        for s in range(nseg):
            v = -np.linalg.solve(Ls[s].T, Hx[s] @ xi + hh[s][:, r])
            nxt = np.zeros(5)
            nxt[:4] = Phi[s] @ xi[:4] + Gam[s] @ v + (ev[s] if r == 0 else 0.0)
            nxt[4] = v[S - 1]
            traj_v[r, s * S:(s + 1) * S] = v
            traj_x[r, s + 1] = nxt[:4]
            xi = nxt
    if hard:
        Sm = np.array([[traj_x[1 + j, NS - 1, i] for j in range(len(hard))] for i in hard])
        rhs = -np.array([traj_x[0, NS - 1, i] + dT[i] for i in hard])
        mu = np.linalg.solve(Sm, rhs)
        coef = np.concatenate([[1.0], mu])
    else:
        coef = np.array([1.0])
    dx = np.tensordot(coef, traj_x, axes=1).reshape(-1)
    du = coef @ traj_v
    return np.concatenate([dx, du])


def riccati_knot(Phi, Gam, ev, xs, u, x_cur, set_point, u_prev, wts, lam, W, S):
    """Knot-level Riccati recursion in end-of-segment coordinates (round 2): one SCALAR control per step.

    Inside segment s the state is z = (y, p): y = Phi_s dx_s + e_s + sum_{j<i} Gam_s[:, j] v_j (the segment's end state
    as far as it is determined by the controls already applied) and p = the previous control.  One knot:
        z+ = A z + b v,  A = diag(1,1,1,1,0),  b = (Gam_s[:, i], 1),
        cost = 1/2 (wu2 + lam) v^2 + 1/2 wr2 (v - p)^2 + g_i v
    so every elimination is a scalar pivot  s = wu2 + lam + wr2 + b^T P b  and the value function over (z | rho) stays
    a (5 + nr) x (5 + nr) matrix (rho = affine coordinate + one multiplier per hard terminal row).  At a shooting node
    the value function is pulled back through y = Phi dx + e."""
    w_u, w_r, wt = wts[0], wts[1], wts[2:]
    nseg = W // S
    NS = nseg + 1
    target = np.array([set_point, math.pi / 2, 0.0, 0.0])
    wu2, wr2 = w_u ** 2, w_r ** 2
    g = np.zeros(W)
    for k in range(W):
        g[k] = wu2 * u[k]
        if k + 1 < W:
            g[k] += wr2 * (u[k] - u[k + 1])
        if k > 0:
            g[k] -= wr2 * (u[k - 1] - u[k])
        if k == 0:
            g[k] += wr2 * (u[k] - u_prev)
    dT = xs[NS - 1] - target
    dT[1] = wrap(dT[1])
    hard = [i for i in range(4) if wt[i] < 0]
    nr = 1 + len(hard)
    P = np.zeros((5, 5))
    Q = np.zeros((5, nr))
    C = np.zeros((nr, nr))
    for i in range(4):
        P[i, i] = lam
        if wt[i] >= 0:
            P[i, i] += wt[i] ** 2
            Q[i, 0] = wt[i] ** 2 * dT[i]
    for j, i in enumerate(hard):
        Q[i, 1 + j] = 1.0
    kz = np.zeros((W, 5))
    kr = np.zeros((W, nr))
    for s in range(nseg - 1, -1, -1):
        for i in range(S - 1, -1, -1):
            k = s * S + i
            b = np.array([Gam[s][0, i], Gam[s][1, i], Gam[s][2, i], Gam[s][3, i], 1.0])
            Pb = P @ b
            # the rate cost between v_k and its successor is carried by P[4,4] / Q (added when the successor was eliminated);
            # here: own cost + the rate term towards the PREVIOUS control p (also for k = 0: towards the applied control,
            # whose linear part is already in g[0])
            piv = wu2 + lam + wr2 + b @ Pb
            if not piv > 0:
                raise FloatingPointError("indefinite")
            rz = np.array([Pb[0], Pb[1], Pb[2], Pb[3], -wr2 if k > 0 else 0.0])
            rr = Q.T @ b
            rr[0] += g[k]
            inv = 1.0 / piv
            kz[k] = rz * inv
            kr[k] = rr * inv
            Pn = P.copy()
            Pn[4, :] = 0.0
            Pn[:, 4] = 0.0
            Pn[4, 4] = wr2 if k > 0 else 0.0
            P = Pn - np.outer(rz, rz) * inv
            Qn = Q.copy()
            Qn[4, :] = 0.0
            Q = Qn - np.outer(rz, rr) * inv
            C = C - np.outer(rr, rr) * inv
        # node s: y = Phi dx + e rho_0, p unchanged; lambda on dx_s (not on the fixed dx_0)
        F = np.eye(5)
        F[:4, :4] = Phi[s]
        f = np.zeros(5)
        f[:4] = ev[s]
        Pf = P @ f
        C[0, :] += f @ Q
        C[:, 0] += f @ Q
        C[0, 0] += f @ Pf
        Q[:, 0] += Pf
        Q = F.T @ Q
        P = F.T @ P @ F
        if s > 0:
            for i in range(4):
                P[i, i] += lam
    d0 = xs[0] - x_cur
    d0[1] = wrap(d0[1])
    m0 = -d0
    coef = np.zeros(nr)
    coef[0] = 1.0
    if hard:
        # stationarity in the multipliers: Q[:4, 1+j] . m0 + C[1+j, 0] + sum_q C[1+j, 1+q] mu_q + dT_j = 0
        nh = len(hard)
        Mh = -C[1:, 1:]
        rhs = np.array([Q[:4, 1 + j] @ m0 + C[1 + j, 0] + dT[hard[j]] for j in range(nh)])
        coef[1:] = np.linalg.solve(Mh, rhs)
    dx = [m0]
    du = np.zeros(W)
    p = 0.0
    x = m0
    for s in range(nseg):
        y = Phi[s] @ x + ev[s]
        for i in range(S):
            k = s * S + i
            z = np.array([y[0], y[1], y[2], y[3], p])
            v = -(kz[k] @ z + kr[k] @ coef)
            y = y + Gam[s][:, i] * v
            p = v
            du[k] = v
        x = y
        dx.append(x)
    return np.concatenate([np.concatenate(dx), du])


def iterate_point(dyn, x0, W, S, dt, n_newton, wts, set_point):
    """A realistic linearisation point: cold-start guess, then n_newton full QP steps (alpha = 0.5) with retraction."""
    nseg = W // S
    u = np.array([10.0 * math.sin(k / W * 2 * math.pi) for k in range(W)])
    xs = [x0.copy()]
    x = x0.copy()
    for k in range(W):
        x = O.rk4(dyn, x, u[k], dt)
        x[1] = wrap(x[1])
        if (k + 1) % S == 0:
            xs.append(x.copy())
    for _ in range(n_newton):
        Phi, Gam, ev = linearise(dyn, xs, u, x0, set_point, W, S, dt)
        G, c, A, b = dense_qp(Phi, Gam, ev, xs, u, x0, set_point, 0.0, wts, 0.0, W, S)
        st, dx = O.qp_nullspace(G, c, A, b)
        al = 0.5
        for s in range(nseg + 1):
            xs[s] = xs[s] + al * dx[4 * s:4 * s + 4]
            xs[s][1] = wrap(xs[s][1])
            xs[s][0] = min(max(xs[s][0], -5.0), 5.0)
        u = np.clip(u + al * dx[4 * (nseg + 1):], -300, 300)
    return xs, u


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--W", type=int, default=160)
    ap.add_argument("--S", type=int, default=10)
    ap.add_argument("--n", type=int, default=6)
    ap.add_argument("--lam", type=float, default=0.0)
    ap.add_argument("--newton", type=int, default=2)
    args = ap.parse_args()
    W, S, dt = args.W, args.S, 0.01
    states, dyn, sp, weights = synthetic.sweep_batch(args.n)
    print(f"W={W} S={S} lambda={args.lam}; relative error of dx (max-norm) vs longdouble KKT")
    for i in range(args.n):
        wts = weights[i]
        xs, u = iterate_point(dyn[i], states[i], W, S, dt, args.newton, wts, sp[i])
        Phi, Gam, ev = linearise(dyn[i], xs, u, states[i], sp[i], W, S, dt)
        G, c, A, b = dense_qp(Phi, Gam, ev, xs, u, states[i], sp[i], 0.0, wts, args.lam, W, S)
        truth = kkt_longdouble(G, c, A, b)
        scale = np.abs(truth).max()
        st, ns = O.qp_nullspace(G, c, A, b)
        out = {"nullspace": np.abs(ns - truth).max() / scale}
        if args.lam == 0.0:
            out["condensed"] = np.abs(condensed(Phi, Gam, ev, xs, u, states[i], sp[i], 0.0, wts, 0.0, W, S) - truth).max() / scale
        out["riccati"] = np.abs(riccati(Phi, Gam, ev, xs, u, states[i], sp[i], 0.0, wts, args.lam, W, S) - truth).max() / scale
        out["riccati_knot"] = np.abs(riccati_knot(Phi, Gam, ev, xs, u, states[i], sp[i], 0.0, wts, args.lam, W, S) - truth).max() / scale
        print(i, f"|dx|max {scale:.3e}", {k: f"{v:.2e}" for k, v in out.items()})


if __name__ == "__main__":
    main()
