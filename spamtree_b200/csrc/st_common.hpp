// st_common.hpp — shared host-side types of libspamtree_b200 (product code; never includes oracle/)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace st {

typedef std::vector<int64_t> ivec;
typedef std::vector<double> dvec;

struct CSR {
  ivec ptr, idx;
  int64_t size() const { return ptr.empty() ? 0 : (int64_t)ptr.size() - 1; }
  int64_t len(int64_t i) const { return ptr[i + 1] - ptr[i]; }
  const int64_t* row(int64_t i) const { return idx.data() + ptr[i]; }
  int64_t back(int64_t i) const { return idx[ptr[i + 1] - 1]; }
};

// Host random stream ("rng_mode 0"): xoshiro256++ seeded through splitmix64, normals by Box-Muller using both
// outputs, gamma by Marsaglia-Tsang.  Stands in for R's RNG, which the reference reaches through Rcpp
// (arma::randn, R::runif, R::rgamma; spamtree_model.cpp:1018,1378,1405, mh_adapt.h:30).
class HostRng {
 public:
  void seed(uint64_t sd) {
    for (int i = 0; i < 4; i++) s_[i] = splitmix(sd);
    have_ = false;
  }
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  uint64_t next() {
    const uint64_t r = rotl(s_[0] + s_[3], 23) + s_[0];
    const uint64_t t = s_[1] << 17;
    s_[2] ^= s_[0];
    s_[3] ^= s_[1];
    s_[1] ^= s_[2];
    s_[0] ^= s_[3];
    s_[2] ^= t;
    s_[3] = rotl(s_[3], 45);
    return r;
  }
  double unif() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  double norm() {
    if (have_) {
      have_ = false;
      return spare_;
    }
    const double u1 = unif(), u2 = unif();
    const double rad = std::sqrt(-2.0 * std::log(u1));
    const double ang = 6.283185307179586476925286766559 * u2;
    spare_ = rad * std::sin(ang);
    have_ = true;
    return rad * std::cos(ang);
  }
  double gamma(double shape, double scale) {
    const double d = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
    for (;;) {
      double x, v;
      do {
        x = norm();
        v = 1.0 + c * x;
      } while (v <= 0);
      v = v * v * v;
      const double u = unif();
      if (u < 1.0 - 0.0331 * x * x * x * x) return d * v * scale;
      if (std::log(u) < 0.5 * x * x + d * (1.0 - v + std::log(v))) return d * v * scale;
    }
  }

 private:
  static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t s_[4] = {1, 2, 3, 4};
  bool have_ = false;
  double spare_ = 0;
};

// tiny dense helpers for the p x p and npar x npar host work (column-major)
struct SmallMat {
  int n = 0;
  dvec a;
  SmallMat() {}
  explicit SmallMat(int n_) : n(n_), a((size_t)n_ * n_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i + (size_t)j * n]; }
  double operator()(int i, int j) const { return a[(size_t)i + (size_t)j * n]; }
};
// lower Cholesky in place (reads the lower triangle); false if not positive definite
inline bool small_chol(SmallMat& A) {
  const int n = A.n;
  for (int j = 0; j < n; j++) {
    double d = A(j, j);
    for (int k = 0; k < j; k++) d -= A(j, k) * A(j, k);
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A(j, j) = d;
    for (int i = j + 1; i < n; i++) {
      double s = A(i, j);
      for (int k = 0; k < j; k++) s -= A(i, k) * A(j, k);
      A(i, j) = s / d;
    }
  }
  for (int j = 0; j < n; j++)
    for (int i = 0; i < j; i++) A(i, j) = 0.0;
  return true;
}
inline SmallMat small_inv_lower(const SmallMat& L) {
  const int n = L.n;
  SmallMat X(n);
  for (int j = 0; j < n; j++) {
    X(j, j) = 1.0 / L(j, j);
    for (int i = j + 1; i < n; i++) {
      double s = 0;
      for (int k = j; k < i; k++) s += L(i, k) * X(k, j);
      X(i, j) = -s / L(i, i);
    }
  }
  return X;
}

}  // namespace st
