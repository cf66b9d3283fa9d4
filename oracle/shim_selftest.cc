// ORACLE — TEST INFRASTRUCTURE ONLY.  Self-test of oracle/eigen_shim (the Eigen-interface shim that lets the
// reference's own translation units compile, see eigen_shim/Eigen/Core): every operation the reference's hot-path
// sources spell is checked against plain loops on random data.  Needs neither the reference nor a GPU:
//     g++ -std=c++17 -O1 -Ieigen_shim shim_selftest.cc -o shim_selftest && ./shim_selftest
// Prints "SHIM_SELFTEST_OK <number of checks>" or the first failing check.
#include <Eigen/Dense>

#include <cmath>
#include <cstdio>
#include <random>
#include <sstream>
#include <vector>

using Eigen::Index;
using Eigen::MatrixXd;
using Eigen::VectorXd;

static int g_checks = 0;
#define CHECK(cond)                                               \
  do {                                                            \
    ++g_checks;                                                   \
    if (!(cond)) {                                                \
      std::printf("FAILED line %d: %s\n", __LINE__, #cond);       \
      return 1;                                                   \
    }                                                             \
  } while (0)

static std::mt19937_64 rng(12345);
static double rnd() { return std::uniform_real_distribution<double>(-1.0, 1.0)(rng); }
static MatrixXd RandomMatrix(Index r, Index c) {
  MatrixXd m(r, c);
  for (Index j = 0; j < c; ++j)
    for (Index i = 0; i < r; ++i) m(i, j) = rnd();
  return m;
}
static VectorXd RandomVector(Index n) {
  VectorXd v(n);
  for (Index i = 0; i < n; ++i) v[i] = rnd();
  return v;
}
static double MaxAbsDiff(const MatrixXd& a, const MatrixXd& b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) return 1e300;
  double d = 0.0;
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) d = std::max(d, std::abs(a(i, j) - b(i, j)));
  return d;
}

int main() {
  // ---- construction, fixed sizes, comma initialiser, coefficient access ------------------------------------------
  Eigen::Vector4d v4(1.0, 2.0, 3.0, 4.0);
  CHECK(v4[0] == 1.0 && v4(3) == 4.0 && v4.rows() == 4 && v4.cols() == 1 && v4.size() == 4);
  Eigen::Vector2d v2{5.0, 6.0};
  CHECK(v2[1] == 6.0);
  Eigen::Matrix<double, 1, 1> one{7.5};
  CHECK(one[0] == 7.5 && one.size() == 1);
  const Eigen::Matrix2d A2 = (Eigen::Matrix2d() << 1.0, 2.0, 3.0, 4.0).finished();
  CHECK(A2(0, 0) == 1.0 && A2(0, 1) == 2.0 && A2(1, 0) == 3.0 && A2(1, 1) == 4.0);  // row by row
  const Eigen::Matrix2d A2i = A2.inverse();
  CHECK(std::abs(A2i(0, 0) + 2.0) < 1e-15 && std::abs(A2i(0, 1) - 1.0) < 1e-15 && std::abs(A2i(1, 0) - 1.5) < 1e-15 &&
        std::abs(A2i(1, 1) + 0.5) < 1e-15);
  const auto I4 = Eigen::Matrix<double, 4, 4>::Identity();
  CHECK(I4(2, 2) == 1.0 && I4(2, 1) == 0.0);
  VectorXd z = VectorXd::Zero(5);
  CHECK(z.rows() == 5 && z.sum() == 0.0);
  MatrixXd zz = MatrixXd::Zero(3, 2);
  CHECK(zz.rows() == 3 && zz.cols() == 2 && zz.squaredNorm() == 0.0);
  Eigen::Matrix<double, 4, Eigen::Dynamic> J;
  J.resize(4, 18);
  CHECK(J.rows() == 4 && J.cols() == 18);

  // ---- views write through: segment / head / tail / block / rows / cols / diagonal / transpose -------------------
  VectorXd x = RandomVector(12);
  VectorXd x0 = x;
  x.segment(2, 3).setZero();
  CHECK(x[1] == x0[1] && x[2] == 0.0 && x[4] == 0.0 && x[5] == x0[5]);
  x.segment<4>(4) = v4;
  CHECK(x[4] == 1.0 && x[7] == 4.0 && x[8] == x0[8]);
  x.head(2) = x0.tail(2);
  CHECK(x[0] == x0[10] && x[1] == x0[11]);
  const VectorXd& xc = x;
  Eigen::VectorBlock<const VectorXd> cview = xc.tail(3);
  CHECK(cview.rows() == 3 && cview[2] == x[11]);
  std::vector<double> copied{cview.begin(), cview.end()};
  CHECK(copied.size() == 3 && copied[0] == x[9]);
  MatrixXd M = RandomMatrix(6, 5);
  const MatrixXd M0 = M;
  M.block(1, 2, 2, 2).setConstant(9.0);
  CHECK(M(1, 2) == 9.0 && M(2, 3) == 9.0 && M(0, 2) == M0(0, 2) && M(3, 3) == M0(3, 3));
  M = M0;
  M.middleRows(2, 2).col(1).setZero();
  CHECK(M(2, 1) == 0.0 && M(3, 1) == 0.0 && M(1, 1) == M0(1, 1));
  M = M0;
  M.topRows(2).middleCols(1, 2).diagonal().setConstant(-1.0);
  CHECK(M(0, 1) == -1.0 && M(1, 2) == -1.0 && M(0, 2) == M0(0, 2));
  M = M0;
  M.bottomLeftCorner(2, 3) = RandomMatrix(2, 3) * 0.0;
  CHECK(M(4, 0) == 0.0 && M(5, 2) == 0.0 && M(5, 3) == M0(5, 3) && M(3, 0) == M0(3, 0));
  M = M0;
  M.diagonal().array() += 2.0;
  CHECK(M(3, 3) == M0(3, 3) + 2.0 && M(3, 2) == M0(3, 2));
  CHECK(M0.transpose()(1, 4) == M0(4, 1) && M0.transpose().rows() == 5);
  J.setZero();
  J.middleCols<4>(4) = -Eigen::Matrix<double, 4, 4>::Identity();
  J.leftCols<4>() = I4;
  J.middleCols<1>(9).noalias() = I4 * v4;
  CHECK(J(1, 5) == -1.0 && J(1, 1) == 1.0 && J(2, 9) == 3.0 && J(2, 8) == 0.0);
  VectorXd a = RandomVector(7), b = RandomVector(7);
  a.swap(b);
  b.resizeLike(a);

  // ---- arithmetic against plain loops ---------------------------------------------------------------------------
  const MatrixXd P = RandomMatrix(7, 5), Q = RandomMatrix(5, 6);
  const MatrixXd PQ = P * Q;
  for (Index i = 0; i < 7; ++i)
    for (Index j = 0; j < 6; ++j) {
      double s = 0.0;
      for (Index k = 0; k < 5; ++k) s += P(i, k) * Q(k, j);
      CHECK(PQ(i, j) == s);  // same order of accumulation: bit-equal
    }
  const MatrixXd PtP = P.transpose() * P;  // strided left operand
  for (Index i = 0; i < 5; ++i)
    for (Index j = 0; j < 5; ++j) {
      double s = 0.0;
      for (Index k = 0; k < 7; ++k) s += P(k, i) * P(k, j);
      CHECK(PtP(i, j) == s);
    }
  const VectorXd u = RandomVector(5), w = RandomVector(5);
  const VectorXd lin = u + (0.5 / 6.0) * (w + u * 2.0 + w * 2.0 + u);
  for (Index i = 0; i < 5; ++i) CHECK(lin[i] == u[i] + (0.5 / 6.0) * (((w[i] + u[i] * 2.0) + w[i] * 2.0) + u[i]));
  const VectorXd half = (u + w * 0.25 / 2.0).eval();
  for (Index i = 0; i < 5; ++i) CHECK(half[i] == u[i] + (w[i] * 0.25) / 2.0);
  const VectorXd neg = P * -u;
  const VectorXd pos = P * u;
  for (Index i = 0; i < 7; ++i) CHECK(neg[i] == -pos[i]);
  double dot = 0.0, l1 = 0.0, sq = 0.0, linf = 0.0;
  for (Index i = 0; i < 5; ++i) {
    dot += u[i] * w[i];
    l1 += std::abs(u[i]);
    sq += u[i] * u[i];
    linf = std::max(linf, std::abs(u[i]));
  }
  CHECK(u.dot(w) == dot && u.lpNorm<1>() == l1 && u.squaredNorm() == sq && u.norm() == std::sqrt(sq) &&
        u.lpNorm<Eigen::Infinity>() == linf && u.cwiseAbs().maxCoeff() == linf);
  CHECK(P.row(2).dot(u) == (P.transpose().col(2)).dot(u));
  CHECK((u.array() > -2.0).all() && !(u.array() > 2.0).all());
  const VectorXd quo = u.array() / w.array();
  for (Index i = 0; i < 5; ++i) CHECK(quo[i] == u[i] / w[i]);

  // ---- self-adjoint and triangular views, LLT, permutations, column-pivoted QR ---------------------------------
  const Index n = 9, k = 4;
  MatrixXd L0 = RandomMatrix(n, n);
  MatrixXd G = L0 * L0.transpose();
  for (Index i = 0; i < n; ++i) G(i, i) += 1.0;
  MatrixXd Glow = G;
  for (Index j = 0; j < n; ++j)
    for (Index i = 0; i < j; ++i) Glow(i, j) = 1e300;  // the upper triangle must never be read
  const VectorXd r = RandomVector(n);
  const VectorXd Gr = Glow.selfadjointView<Eigen::Lower>() * r;
  CHECK(MaxAbsDiff(Gr, G * r) < 1e-13);
  const auto llt = Glow.selfadjointView<Eigen::Lower>().llt();
  CHECK(llt.info() == Eigen::ComputationInfo::Success);
  VectorXd sol = r;
  llt.solveInPlace(sol);
  CHECK(MaxAbsDiff(G * sol, r) < 1e-12);
  MatrixXd neg_def = -G;
  CHECK(neg_def.selfadjointView<Eigen::Lower>().llt().info() != Eigen::ComputationInfo::Success);
  MatrixXd H = MatrixXd::Zero(n, n);
  H.topLeftCorner(n, n).triangularView<Eigen::Lower>() = Glow.triangularView<Eigen::Lower>();
  CHECK(H(3, 1) == G(3, 1) && H(1, 3) == 0.0);
  H = G.selfadjointView<Eigen::Lower>();
  CHECK(MaxAbsDiff(H, G) == 0.0);

  const MatrixXd Aeq = RandomMatrix(k, n);  // K x N, full row rank
  const auto QR = Aeq.transpose().colPivHouseholderQr();
  const MatrixXd Qm = QR.matrixQ();
  const auto& R = QR.matrixR();
  const auto& Pm = QR.colsPermutation();
  CHECK(QR.rank() == k && Qm.rows() == n && Qm.cols() == n && Pm.rows() == k);
  CHECK(MaxAbsDiff(Qm.transpose() * Qm, MatrixXd::Identity(n, n)) < 1e-14);
  // A^T P = Q R  <=>  column j of Q R equals column indices[j] of A^T
  MatrixXd Rtop = MatrixXd::Zero(n, k);
  for (Index j = 0; j < k; ++j)
    for (Index i = 0; i <= j; ++i) Rtop(i, j) = R(i, j);
  const MatrixXd QRm = Qm * Rtop;
  for (Index j = 0; j < k; ++j)
    for (Index i = 0; i < n; ++i) CHECK(std::abs(QRm(i, j) - Aeq(Pm.indices()[static_cast<std::size_t>(j)], i)) < 1e-14);
  for (Index j = 0; j + 1 < k; ++j) CHECK(std::abs(R(j, j)) >= std::abs(R(j + 1, j + 1)) - 1e-15);  // pivoting order
  // the particular solution of qp.cc:703-706: u = Q1 R1^-T P^T (-b) satisfies A u + b = 0
  const VectorXd beq = RandomVector(k);
  const VectorXd perm = Pm.transpose() * -beq;
  const auto R_upper = R.topRows(k).template triangularView<Eigen::Upper>();
  const VectorXd upart = Qm.leftCols(k) * R_upper.transpose().solve(perm);
  CHECK(MaxAbsDiff(Aeq * upart + beq, VectorXd::Zero(k)) < 1e-13);
  // the null-space basis annihilates A
  CHECK(MaxAbsDiff(Aeq * Qm.rightCols(n - k), MatrixXd::Zero(k, n - k)) < 1e-14);

  // ---- pieces only the interior-point half of qp.cc touches -----------------------------------------------------
  const Eigen::LDLT<MatrixXd, Eigen::Lower> ldlt(G);
  CHECK(ldlt.info() == Eigen::ComputationInfo::Success);
  MatrixXd Ginv = MatrixXd::Identity(n, n);
  ldlt.solveInPlace(Ginv);
  CHECK(MaxAbsDiff(G * Ginv, MatrixXd::Identity(n, n)) < 1e-11);
  const Eigen::SelfAdjointEigenSolver<MatrixXd> eig(G);
  double trace = 0.0;
  for (Index i = 0; i < n; ++i) trace += G(i, i);
  CHECK(std::abs(eig.eigenvalues().sum() - trace) < 1e-10 && eig.eigenvalues().minCoeff() > 0.99);
  std::ostringstream os;
  os << v2.transpose().format(Eigen::IOFormat(Eigen::FullPrecision, 0, ", ", ",\n", "[", "]", "[", "]"));
  CHECK(os.str() == "[[5, 6]]");

  std::printf("SHIM_SELFTEST_OK %d\n", g_checks);
  return 0;
}
