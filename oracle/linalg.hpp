// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing in the product path (cart_pole_mpc_b200/, include/) may
// include, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do.
//
// Small dense linear algebra used by the CPU restatement of mini_opt's null-space QP.
//
// The reference delegates this arithmetic to Eigen 3.4.0 (submodule commit 3147391d, NOT vendored under
// /root/reference and not installed here).  The routines below restate the published algorithms of the
// Eigen entry points the hot path calls (call sites: dependencies/mini_opt/source/qp.cc:687-727):
//   * ColPivHouseholderQR::computeInPlace / rank() / householderQ()   (qp.cc:687,689,696)
//   * TriangularView<Upper>::transpose().solve()                       (qp.cc:706)
//   * SelfAdjointView<Lower> products                                 (qp.cc:710,720; nonlinear.cc:499)
//   * LLT<Lower>::compute / solveInPlace                               (qp.cc:713-723)
// Bit-for-bit agreement with Eigen's blocked kernels is not claimed (summation order inside Eigen's GEMM is
// an implementation detail); every routine is backward stable, which is what fixes the QP step.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

namespace oracle {

using Vec = std::vector<double>;

// Row-major dense matrix with value semantics.
struct Mat {
  int rows{0};
  int cols{0};
  std::vector<double> a;

  Mat() = default;
  Mat(int r, int c) : rows(r), cols(c), a(static_cast<std::size_t>(r) * c, 0.0) {}
  void Resize(int r, int c) {
    rows = r;
    cols = c;
    a.assign(static_cast<std::size_t>(r) * c, 0.0);
  }
  void SetZero() { std::fill(a.begin(), a.end(), 0.0); }
  double& operator()(int r, int c) { return a[static_cast<std::size_t>(r) * cols + c]; }
  double operator()(int r, int c) const { return a[static_cast<std::size_t>(r) * cols + c]; }
};

inline double Dot(const Vec& x, const Vec& y) {
  double s = 0.0;
  for (std::size_t i = 0; i < x.size(); ++i) s += x[i] * y[i];
  return s;
}

inline double L1Norm(const double* x, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += std::abs(x[i]);
  return s;
}

// y = sym(G) * x where only the lower triangle of G holds valid data (residual.hpp:204-221).
inline Vec SymLowerTimes(const Mat& G, const Vec& x) {
  const int n = G.rows;
  Vec y(n, 0.0);
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j <= i; ++j) s += G(i, j) * x[j];
    for (int j = i + 1; j < n; ++j) s += G(j, i) * x[j];
    y[i] = s;
  }
  return y;
}

// Column-pivoted Householder QR of an (n x k) matrix, following the structure of Eigen 3.4
// ColPivHouseholderQR::computeInPlace (norm down-dating with recomputation when cancellation is detected).
struct ColPivQR {
  Mat qr;                 // Householder vectors below the diagonal, R on/above it
  Vec tau;                // Householder coefficients
  std::vector<int> perm;  // perm[j] = original column placed at position j
  int nonzero_pivots{0};
  double max_pivot{0.0};
  Vec wbuf;  // scratch of the reflector application

  void Compute(const Mat& m) {
    qr = m;
    const int rows = qr.rows, cols = qr.cols;
    const int size = std::min(rows, cols);
    tau.assign(size, 0.0);
    perm.resize(cols);
    for (int j = 0; j < cols; ++j) perm[j] = j;

    Vec norms_updated(cols), norms_direct(cols);
    for (int j = 0; j < cols; ++j) {
      double s = 0.0;
      for (int i = 0; i < rows; ++i) s += qr(i, j) * qr(i, j);
      norms_direct[j] = norms_updated[j] = std::sqrt(s);
    }
    const double eps = std::numeric_limits<double>::epsilon();
    const double max_norm = cols > 0 ? *std::max_element(norms_updated.begin(), norms_updated.end()) : 0.0;
    const double threshold_helper = (max_norm * eps / rows) * (max_norm * eps / rows);
    const double downdate_threshold = std::sqrt(eps);

    nonzero_pivots = size;
    max_pivot = 0.0;
    for (int k = 0; k < size; ++k) {
      // pick the remaining column of largest (tracked) norm
      int best = k;
      for (int j = k + 1; j < cols; ++j) {
        if (norms_updated[j] > norms_updated[best]) best = j;
      }
      const double best_sq = norms_updated[best] * norms_updated[best];
      if (nonzero_pivots == size && best_sq < threshold_helper * static_cast<double>(rows - k)) {
        nonzero_pivots = k;
      }
      if (best != k) {
        for (int i = 0; i < rows; ++i) std::swap(qr(i, k), qr(i, best));
        std::swap(norms_updated[k], norms_updated[best]);
        std::swap(norms_direct[k], norms_direct[best]);
        std::swap(perm[k], perm[best]);
      }

      // Householder reflector for column k, rows k..rows-1 (Eigen makeHouseholderInPlace)
      const double c0 = qr(k, k);
      double tail_sq = 0.0;
      for (int i = k + 1; i < rows; ++i) tail_sq += qr(i, k) * qr(i, k);
      double beta;
      if (tail_sq <= std::numeric_limits<double>::min()) {
        tau[k] = 0.0;
        beta = c0;
        for (int i = k + 1; i < rows; ++i) qr(i, k) = 0.0;
      } else {
        beta = std::sqrt(c0 * c0 + tail_sq);
        if (c0 >= 0.0) beta = -beta;
        const double denom = c0 - beta;
        for (int i = k + 1; i < rows; ++i) qr(i, k) /= denom;
        tau[k] = (beta - c0) / beta;
      }
      qr(k, k) = beta;
      if (std::abs(beta) > max_pivot) max_pivot = std::abs(beta);

      // apply H_k = I - tau v v^T (v = [1; essential]) to the trailing columns.  Swept row by row (the storage is
      // row-major) so that the compiler can vectorise across the columns j; every w[j] and every qr(i, j) still sees
      // exactly the same operations in the same order as the column-by-column form.
      if (k + 1 < cols) {
        wbuf.assign(static_cast<std::size_t>(cols), 0.0);
        double* w = wbuf.data();
        const double* rk = &qr(k, 0);
        for (int j = k + 1; j < cols; ++j) w[j] = rk[j];
        for (int i = k + 1; i < rows; ++i) {
          const double vi = qr(i, k);
          const double* ri = &qr(i, 0);
          for (int j = k + 1; j < cols; ++j) w[j] += vi * ri[j];
        }
        const double tk = tau[k];
        double* rkw = &qr(k, 0);
        for (int j = k + 1; j < cols; ++j) {
          w[j] *= tk;
          rkw[j] -= w[j];
        }
        for (int i = k + 1; i < rows; ++i) {
          const double vi = qr(i, k);
          double* ri = &qr(i, 0);
          for (int j = k + 1; j < cols; ++j) ri[j] -= vi * w[j];
        }
      }

      // down-date the tracked column norms
      for (int j = k + 1; j < cols; ++j) {
        if (norms_updated[j] != 0.0) {
          double t = std::abs(qr(k, j)) / norms_updated[j];
          t = (1.0 + t) * (1.0 - t);
          if (t < 0.0) t = 0.0;
          const double ratio = norms_updated[j] / norms_direct[j];
          const double t2 = t * ratio * ratio;
          if (t2 <= downdate_threshold) {
            double s = 0.0;
            for (int i = k + 1; i < rows; ++i) s += qr(i, j) * qr(i, j);
            norms_direct[j] = norms_updated[j] = std::sqrt(s);
          } else {
            norms_updated[j] *= std::sqrt(t);
          }
        }
      }
    }
  }

  // Eigen's default rank(): pivots larger than eps * diagSize * |max pivot|.
  int Rank() const {
    const int size = std::min(qr.rows, qr.cols);
    const double thr = std::abs(max_pivot) * std::numeric_limits<double>::epsilon() * static_cast<double>(size);
    int r = 0;
    for (int i = 0; i < nonzero_pivots; ++i) r += (std::abs(qr(i, i)) > thr) ? 1 : 0;
    return r;
  }

  // Explicit square Q = H_0 H_1 ... H_{size-1} (Eigen HouseholderSequence::evalTo, applied right to left).
  Mat MatrixQ() const {
    const int rows = qr.rows;
    const int size = std::min(qr.rows, qr.cols);
    Mat Q(rows, rows);
    for (int i = 0; i < rows; ++i) Q(i, i) = 1.0;
    Vec wv(static_cast<std::size_t>(rows), 0.0);
    double* w = wv.data();
    for (int k = size - 1; k >= 0; --k) {
      // same row-by-row sweep as in Compute(): identical arithmetic per entry, contiguous inner loops
      const double* qk = &Q(k, 0);
      for (int j = k; j < rows; ++j) w[j] = qk[j];
      for (int i = k + 1; i < rows; ++i) {
        const double vi = qr(i, k);
        const double* qi = &Q(i, 0);
        for (int j = k; j < rows; ++j) w[j] += vi * qi[j];
      }
      const double tk = tau[k];
      double* qkw = &Q(k, 0);
      for (int j = k; j < rows; ++j) {
        w[j] *= tk;
        qkw[j] -= w[j];
      }
      for (int i = k + 1; i < rows; ++i) {
        const double vi = qr(i, k);
        double* qi = &Q(i, 0);
        for (int j = k; j < rows; ++j) qi[j] -= vi * w[j];
      }
    }
    return Q;
  }
};

// In-place lower Cholesky, unblocked (Eigen llt_inplace<Lower>::unblocked).  Returns -1 on success or the index
// of the first non-positive pivot.
inline int CholeskyLowerInPlace(Mat& A) {
  const int n = A.rows;
  for (int k = 0; k < n; ++k) {
    double x = A(k, k);
    for (int j = 0; j < k; ++j) x -= A(k, j) * A(k, j);
    if (x <= 0.0) return k;
    x = std::sqrt(x);
    A(k, k) = x;
    for (int i = k + 1; i < n; ++i) {
      double s = A(i, k);
      for (int j = 0; j < k; ++j) s -= A(i, j) * A(k, j);
      A(i, k) = s / x;
    }
  }
  return -1;
}

// Solve L L^T y = rhs in place given the factor in the lower triangle of L.
inline void CholeskySolveInPlace(const Mat& L, Vec& y) {
  const int n = L.rows;
  for (int i = 0; i < n; ++i) {
    double s = y[i];
    for (int j = 0; j < i; ++j) s -= L(i, j) * y[j];
    y[i] = s / L(i, i);
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = y[i];
    for (int j = i + 1; j < n; ++j) s -= L(j, i) * y[j];
    y[i] = s / L(i, i);
  }
}

}  // namespace oracle
