"""Derives the cart-mounted double pendulum dynamics with sympy and emits the C++ used by the CUDA path and the oracle.

Physics spec: symbolic/dynamics_double.py:25-148 of the reference (a work-in-progress there: no generated code and no
optimiser exist for it).  Same model: point masses m_1, m_2 at the ends of massless rods l_1, l_2, angles measured from
the +x axis (p_1 = b + l_1 (cos th_1, sin th_1), p_2 = p_1 + l_2 (cos th_2, sin th_2)), V = g (m_1 y_1 + m_2 y_2),
control force u_b on the base, state x = [b_x, th_1, th_2, b_x', th_1', th_2'] (dynamics_double.py:118-129).
wrenfold is not available here; sympy does the Euler-Lagrange derivation.

Instead of differentiating the closed-form A^-1 f (what wrenfold's jacobian() would do), the Jacobians are emitted as
    d(acc)/dz = A^-1 (df/dz - (dA/dz) acc),
which is the same quantity and a few dozen flops.

Usage:  python -m cart_pole_mpc_b200.codegen.double_pendulum_gen   (rewrites the two generated headers in place)
"""
from __future__ import annotations

import os

import sympy as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def derive():
    t = sp.symbols("t", real=True)
    m_b, m_1, m_2, l_1, l_2 = sp.symbols("m_b m_1 m_2 l_1 l_2", positive=True)
    g, u = sp.symbols("g u", real=True)
    b = sp.Function("b")(t)
    th1 = sp.Function("th1")(t)
    th2 = sp.Function("th2")(t)
    q = [b, th1, th2]
    qd = [v.diff(t) for v in q]
    qdd = [v.diff(t, 2) for v in q]
    p1 = sp.Matrix([b + l_1 * sp.cos(th1), l_1 * sp.sin(th1)])
    p2 = p1 + sp.Matrix([l_2 * sp.cos(th2), l_2 * sp.sin(th2)])
    bd = sp.Matrix([b.diff(t), 0])
    T = (m_b * bd.dot(bd) + m_1 * p1.diff(t).dot(p1.diff(t)) + m_2 * p2.diff(t).dot(p2.diff(t))) / 2
    V = g * m_1 * p1[1] + g * m_2 * p2[1]
    L = T - V
    force = [u, 0, 0]
    eqs = [sp.expand(L.diff(qd[i]).diff(t) - L.diff(q[i]) - force[i]) for i in range(3)]
    A = sp.zeros(3, 3)
    f = sp.zeros(3, 1)
    for i, eq in enumerate(eqs):
        rest = eq
        for j in range(3):
            cf = eq.coeff(qdd[j])
            A[i, j] = cf
            rest = sp.expand(rest - cf * qdd[j])
        f[i] = -rest
    # plain symbols
    xb, x1, x2, vb, v1, v2 = sp.symbols("b_x th_1 th_2 b_xd th_1d th_2d", real=True)
    rep = [(qd[0], vb), (qd[1], v1), (qd[2], v2), (b, xb), (th1, x1), (th2, x2)]
    A = sp.simplify(A.subs(rep))
    f = sp.simplify(f.subs(rep))
    # only sin/cos of the two angles themselves (the reference's generator performs no trig fusion either)
    A = A.applyfunc(lambda e: sp.expand(sp.expand_trig(e)))
    f = f.applyfunc(lambda e: sp.expand(sp.expand_trig(e)))
    assert A == A.T
    return dict(A=A, f=f, params=(m_b, m_1, m_2, l_1, l_2, g), x=(xb, x1, x2, vb, v1, v2), u=u)


def build_expressions(d, hoist):
    """Returns (const_defs, assignments, outputs) as lists of (symbol, expr)."""
    A, f = d["A"], d["f"]
    m_b, m_1, m_2, l_1, l_2, g = d["params"]
    xb, x1, x2, vb, v1, v2 = d["x"]
    u = d["u"]
    s1, c1, s2, c2 = sp.symbols("s1 c1 s2 c2", real=True)
    trig = [(sp.sin(x1), s1), (sp.cos(x1), c1), (sp.sin(x2), s2), (sp.cos(x2), c2)]
    consts = []
    if hoist:
        k = sp.symbols("P.mt P.h1 P.h2 P.a11 P.a12 P.a22 P.g1 P.g2", real=True)
        exprs = [m_b + m_1 + m_2, (m_1 + m_2) * l_1, m_2 * l_2, (m_1 + m_2) * l_1**2, m_2 * l_1 * l_2, m_2 * l_2**2,
                 g * (m_1 + m_2) * l_1, g * m_2 * l_2]
        consts = list(zip(k, exprs))

    def prep(e):
        e = sp.expand(e.subs(trig))
        if hoist:
            # rewrite every monomial's parameter part through the hoisted constants
            e = sp.collect(e, [s1, c1, s2, c2, v1, v2, u], evaluate=True)
            e = _hoist(e, d, consts, (s1, c1, s2, c2, v1, v2, u))
        return e

    Ae = A.applyfunc(prep)
    fe = f.applyfunc(prep)
    z = [x1, x2, v1, v2]
    # derivatives before the trig substitution
    dA = {zz: A.diff(zz).applyfunc(lambda e: prep(sp.expand(e))) for zz in (x1, x2)}
    df = {zz: f.diff(zz).applyfunc(lambda e: prep(sp.expand(e))) for zz in z}
    return dict(A=Ae, f=fe, dA=dA, df=df, consts=consts, z=z, trig_syms=(s1, c1, s2, c2))


def _hoist(e, d, consts, gens):
    m_b, m_1, m_2, l_1, l_2, g = d["params"]
    poly_terms = sp.Add.make_args(sp.expand(e))
    groups = {}
    for term in poly_terms:
        coeff, rest = sp.S.One, sp.S.One
        for fac in sp.Mul.make_args(term):
            if fac.free_symbols & set(gens):
                rest *= fac
            else:
                coeff *= fac
        groups[rest] = groups.get(rest, 0) + coeff
    out = 0
    for rest, coeff in groups.items():
        coeff = sp.simplify(coeff)
        done = False
        for sym, ce in consts:
            r = sp.simplify(coeff / ce)
            if r.is_number:
                out += r * sym * rest
                done = True
                break
        if not done:
            assert coeff.is_number, f"parameter expression {coeff} is not covered by the hoisted constants"
            out += coeff * rest
    return out


def c_expr(e):
    from sympy.printing.c import C99CodePrinter

    class P(C99CodePrinter):
        def _print_Pow(self, expr):
            if expr.exp == 2:
                b = self._print(expr.base)
                return f"({b} * {b})"
            if expr.exp == -1:
                return f"(1.0 / {self._print(expr.base)})"
            return super()._print_Pow(expr)

        def _print_Symbol(self, expr):
            return expr.name

    return P().doprint(e)


def emit_body(ex, acc_name="acc", indent="  ", rcp="1.0 / det"):
    """Straight-line code: A, f, adjugate inverse, accelerations, Jacobian columns (when JAC)."""
    A, f = ex["A"], ex["f"]
    lines = []
    w = lines.append
    names = {(0, 0): "a00", (0, 1): "a01", (0, 2): "a02", (1, 1): "a11", (1, 2): "a12", (2, 2): "a22"}
    for (i, j), n in names.items():
        w(f"const double {n} = {c_expr(A[i, j])};")
    for i in range(3):
        w(f"const double f{i} = {c_expr(f[i])};")
    # symmetric adjugate
    w("const double k00 = a11 * a22 - a12 * a12;")
    w("const double k01 = a02 * a12 - a01 * a22;")
    w("const double k02 = a01 * a12 - a02 * a11;")
    w("const double k11 = a00 * a22 - a02 * a02;")
    w("const double k12 = a01 * a02 - a00 * a12;")
    w("const double k22 = a00 * a11 - a01 * a01;")
    w("const double det = a00 * k00 + a01 * k01 + a02 * k02;")
    w(f"const double idet = {rcp};")
    for n in ("00", "01", "02", "11", "12", "22"):
        w(f"const double i{n} = k{n} * idet;")
    w(f"{acc_name}[0] = i00 * f0 + i01 * f1 + i02 * f2;")
    w(f"{acc_name}[1] = i01 * f0 + i11 * f1 + i12 * f2;")
    w(f"{acc_name}[2] = i02 * f0 + i12 * f1 + i22 * f2;")
    jl = []
    j = jl.append
    zn = {ex["z"][0]: "th1", ex["z"][1]: "th2", ex["z"][2]: "w1", ex["z"][3]: "w2"}
    for zz in ex["z"]:
        tag = zn[zz]
        for i in range(3):
            e = ex["df"][zz][i]
            if zz in ex["dA"]:
                dA = ex["dA"][zz]
                e = e - sum(dA[i, k] * sp.Symbol(f"{acc_name}[{k}]") for k in range(3))
            j(f"const double r{i}_{tag} = {c_expr(e)};")
        j(f"J_{tag}[0] = i00 * r0_{tag} + i01 * r1_{tag} + i02 * r2_{tag};")
        j(f"J_{tag}[1] = i01 * r0_{tag} + i11 * r1_{tag} + i12 * r2_{tag};")
        j(f"J_{tag}[2] = i02 * r0_{tag} + i12 * r1_{tag} + i22 * r2_{tag};")
    j("J_u[0] = i00;")
    j("J_u[1] = i01;")
    j("J_u[2] = i02;")
    return [indent + l for l in lines], [indent + "  " + l for l in jl]


HEADER_NOTE = "// GENERATED by cart_pole_mpc_b200/codegen/double_pendulum_gen.py - do not edit by hand.\n"


def emit_device(d):
    ex = build_expressions(d, hoist=True)
    body, jac = emit_body(ex, rcp="fm::rcp(det)")
    consts = ex["consts"]
    out = [HEADER_NOTE]
    out.append("""// Cart-mounted double pendulum for sm_100a: accelerations (b_x'', th_1'', th_2'') and their derivatives.
// Model: symbolic/dynamics_double.py:25-148 of the reference (point masses on massless rods, angles from the +x axis,
// control force on the base).  State x = [b_x, th_1, th_2, b_x', th_1', th_2'].  The state Jacobian of x' = f(x, u)
// is [[0 I],[0 J_th1 J_th2 0 J_w1 J_w2]]: b_x and b_x' do not enter the accelerations (no friction or bumpers in this
// model), so only four 3-vectors plus d(acc)/du are produced.  Parameter-only products are hoisted into DpConst.
#pragma once
#include "fastmath.cuh"

namespace cpmpc {

struct DpConst {
  double mt, h1, h2, a11, a12, a22, g1, g2;
};

// The branch-free sincos of fastmath.cuh is valid for |x| < 2^20.  This undamped model blows up in finite time along
// rejected line-search trials (angles of 1e100 and beyond, which the reference-style solver survives because libm
// still returns finite values there: the trial is simply rejected), so huge arguments take CUDA's full-range sincos.
__device__ __forceinline__ void dp_sincos(const double x, double& s, double& c) {
  if (fabs(x) < 1048576.0) {
    fm::sincos_bf(x, s, c);
  } else {
    ::sincos(x, &s, &c);
  }
}

// p = DoubleCartPoleParams {m_b, m_1, m_2, l_1, l_2, g} (dynamics_double.py:13-22)
__device__ __forceinline__ DpConst make_dp_const(const double* __restrict__ p) {
  const double m_b = p[0], m_1 = p[1], m_2 = p[2], l_1 = p[3], l_2 = p[4], g = p[5];
  DpConst P;""")
    for sym, e in consts:
        out.append(f"  {sym.name} = {c_expr(e)};")
    out.append("""  return P;
}

template <bool JAC>
__device__ __forceinline__ void dp_eval(const DpConst& P, const double th_1, const double th_2, const double th_1d,
                                        const double th_2d, const double u, double acc[3], double J_th1[3],
                                        double J_th2[3], double J_w1[3], double J_w2[3], double J_u[3]) {
  double s1, c1, s2, c2;
  dp_sincos(th_1, s1, c1);
  dp_sincos(th_2, s2, c2);""")
    out += body
    out.append("  if (JAC) {")
    out += jac
    out.append("  }")
    out.append("}\n\n}  // namespace cpmpc")
    return "\n".join(out) + "\n"


def emit_oracle(d):
    ex = build_expressions(d, hoist=False)
    body, jac = emit_body(ex)
    out = [HEADER_NOTE]
    out.append("""// ORACLE - TEST INFRASTRUCTURE ONLY.  Host restatement of the double pendulum dynamics specified by
// symbolic/dynamics_double.py:25-148 (the reference ships no generated code for it).  x = [b_x, th_1, th_2, b_x',
// th_1', th_2'], p = {m_b, m_1, m_2, l_1, l_2, g}.  J_x is row-major 6x6, J_u has 6 entries; both optional.
#pragma once
#include <cmath>

namespace oracle {

inline void DoublePendulumDynamics(const double p[6], const double x[6], const double u, double x_dot[6],
                                   double* J_x, double* J_u_out) {
  const double m_b = p[0], m_1 = p[1], m_2 = p[2], l_1 = p[3], l_2 = p[4], g = p[5];
  const double th_1d = x[4], th_2d = x[5];
  const double s1 = std::sin(x[1]), c1 = std::cos(x[1]), s2 = std::sin(x[2]), c2 = std::cos(x[2]);
  double acc[3];""")
    out += body
    out.append("""  x_dot[0] = x[3];
  x_dot[1] = x[4];
  x_dot[2] = x[5];
  x_dot[3] = acc[0];
  x_dot[4] = acc[1];
  x_dot[5] = acc[2];
  if (J_x != nullptr || J_u_out != nullptr) {
    double J_th1[3], J_th2[3], J_w1[3], J_w2[3], J_u[3];""")
    out += jac
    out.append("""    if (J_x != nullptr) {
      for (int i = 0; i < 36; ++i) J_x[i] = 0.0;
      J_x[0 * 6 + 3] = 1.0;
      J_x[1 * 6 + 4] = 1.0;
      J_x[2 * 6 + 5] = 1.0;
      for (int r = 0; r < 3; ++r) {
        J_x[(3 + r) * 6 + 1] = J_th1[r];
        J_x[(3 + r) * 6 + 2] = J_th2[r];
        J_x[(3 + r) * 6 + 4] = J_w1[r];
        J_x[(3 + r) * 6 + 5] = J_w2[r];
      }
    }
    if (J_u_out != nullptr) {
      J_u_out[0] = J_u_out[1] = J_u_out[2] = 0.0;
      J_u_out[3] = J_u[0];
      J_u_out[4] = J_u[1];
      J_u_out[5] = J_u[2];
    }
  }
}

}  // namespace oracle""")
    return "\n".join(out) + "\n"


def main():
    d = derive()
    dev = emit_device(d)
    ora = emit_oracle(d)
    with open(os.path.join(ROOT, "cart_pole_mpc_b200", "csrc", "double_pendulum.cuh"), "w") as fh:
        fh.write(dev)
    with open(os.path.join(ROOT, "oracle", "double_pendulum_gen.hpp"), "w") as fh:
        fh.write(ora)
    print("A =", d["A"])
    print("f =", d["f"])


if __name__ == "__main__":
    main()
