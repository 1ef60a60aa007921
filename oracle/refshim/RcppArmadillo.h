// oracle/refshim/RcppArmadillo.h — TEST INFRASTRUCTURE ONLY.
//
// A minimal, eager (no expression templates) stand-in for the subset of the Armadillo + Rcpp API that the reference's
// model layer uses, so that the reference's OWN, UNMODIFIED sources (/root/reference/src/spamtree_model.cpp,
// covariance_functions.cpp, tree_utils.cpp, tree_dep.cpp, mh_adapt.cpp) compile here without R, Rcpp, Armadillo or BLAS
// and can be run as `oracle/_ref/libspamtree_ref.so` to pin the CPU oracle (oracle/spamtree_oracle.cpp) against the real
// reference code.  Only semantics the reference relies on are implemented; dense kernels are plain loops (dpotrf-,
// dtrtri-like).  Nothing in the product includes this file.
#pragma once
#include <algorithm>
#include <any>
#include <map>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace arma {

typedef unsigned long long uword;
typedef long long sword;

struct arma_tag {};
template <class A> struct is_arma : std::is_base_of<arma_tag, typename std::decay<A>::type> {};

struct SizeMat {
  uword n_rows, n_cols;
};
inline std::ostream& operator<<(std::ostream& o, const SizeMat& s) { return o << s.n_rows << "x" << s.n_cols; }

template <class T> struct Mat;
template <class T> struct Col;
template <class T> struct Row;
template <class T> struct subview;
template <class T> struct subview_elem;
template <class T> struct diagview;

template <class T>
struct Mat : arma_tag {
  typedef T elem_type;
  uword n_rows = 0, n_cols = 0, n_elem = 0;
  std::vector<T> mem;
  Mat() {}
  Mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c, T(0)) {}
  Mat(const T* p, uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem(p, p + r * c) {}
  template <class V, class = typename std::enable_if<is_arma<V>::value && !std::is_base_of<Mat<T>, V>::value>::type>
  Mat(const V& v) { *this = v.eval(); }
  const Mat<T>& eval() const { return *this; }
  template <class V, class = typename std::enable_if<is_arma<V>::value && !std::is_base_of<Mat<T>, V>::value>::type>
  Mat<T>& operator=(const V& v) { Mat<T> t = v.eval(); n_rows = t.n_rows; n_cols = t.n_cols; n_elem = t.n_elem; mem.swap(t.mem); return *this; }
  void set_size(uword r, uword c) { n_rows = r; n_cols = c; n_elem = r * c; mem.assign(n_elem, T(0)); }
  void reset() { n_rows = n_cols = n_elem = 0; mem.clear(); }
  T* memptr() { return mem.data(); }
  const T* memptr() const { return mem.data(); }
  typename std::vector<T>::iterator begin() { return mem.begin(); }
  typename std::vector<T>::iterator end() { return mem.end(); }
  typename std::vector<T>::const_iterator begin() const { return mem.begin(); }
  typename std::vector<T>::const_iterator end() const { return mem.end(); }
  T& operator()(uword i) { return mem[i]; }
  const T& operator()(uword i) const { return mem[i]; }
  T& operator[](uword i) { return mem[i]; }
  const T& operator[](uword i) const { return mem[i]; }
  T& operator()(uword i, uword j) { return mem[i + j * n_rows]; }
  const T& operator()(uword i, uword j) const { return mem[i + j * n_rows]; }
  T& at(uword i, uword j) { return mem[i + j * n_rows]; }
  subview_elem<T> operator()(const Mat<uword>& r, const Mat<uword>& c);
  void fill(T v) { std::fill(mem.begin(), mem.end(), v); }
  Mat<T>& zeros() { fill(T(0)); return *this; }
  // as Armadillo's op_max::direct_max: starts from the most negative value, so NaN entries never win a comparison
  T max() const { (void)mem.at(0); T m = std::numeric_limits<T>::has_infinity ? -std::numeric_limits<T>::infinity() : std::numeric_limits<T>::lowest(); for (auto v : mem) if (v > m) m = v; return m; }
  T min() const { T m = mem.at(0); for (auto v : mem) if (v < m) m = v; return m; }
  bool is_empty() const { return n_elem == 0; }
  Mat<T> t() const {
    Mat<T> o(n_cols, n_rows);
    for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) o(j, i) = (*this)(i, j);
    return o;
  }
  // views
  subview<T> rows(uword a, uword b);
  subview<T> cols(uword a, uword b);
  subview<T> row(uword i);
  subview<T> col(uword j);
  subview<T> submat(uword r1, uword c1, uword r2, uword c2);
  subview<T> subvec(uword a, uword b);
  const subview<T> rows(uword a, uword b) const;
  const subview<T> cols(uword a, uword b) const;
  const subview<T> row(uword i) const;
  const subview<T> col(uword j) const;
  const subview<T> submat(uword r1, uword c1, uword r2, uword c2) const;
  const subview<T> subvec(uword a, uword b) const;
  subview_elem<T> rows(const Mat<uword>& ix);
  subview_elem<T> cols(const Mat<uword>& ix);
  subview_elem<T> elem(const Mat<uword>& ix);
  subview_elem<T> submat(const Mat<uword>& r, const Mat<uword>& c);
  const subview_elem<T> rows(const Mat<uword>& ix) const;
  const subview_elem<T> cols(const Mat<uword>& ix) const;
  const subview_elem<T> elem(const Mat<uword>& ix) const;
  const subview_elem<T> submat(const Mat<uword>& r, const Mat<uword>& c) const;
  diagview<T> diag();
  const diagview<T> diag() const;
  template <class V> typename std::enable_if<is_arma<V>::value, Mat<T>&>::type operator+=(const V& v) {
    const Mat<T>& b = v.eval();
    if (b.n_elem != n_elem) throw std::logic_error("operator+=: size mismatch");
    for (uword i = 0; i < n_elem; i++) mem[i] += b.mem[i];
    return *this;
  }
  template <class V> typename std::enable_if<is_arma<V>::value, Mat<T>&>::type operator-=(const V& v) {
    const Mat<T>& b = v.eval();
    if (b.n_elem != n_elem) throw std::logic_error("operator-=: size mismatch");
    for (uword i = 0; i < n_elem; i++) mem[i] -= b.mem[i];
    return *this;
  }
  Mat<T>& operator+=(T s) { for (auto& v : mem) v += s; return *this; }
  Mat<T>& operator-=(T s) { for (auto& v : mem) v -= s; return *this; }
  Mat<T>& operator*=(T s) { for (auto& v : mem) v *= s; return *this; }
  Mat<T>& operator/=(T s) { for (auto& v : mem) v /= s; return *this; }
  void print(const std::string& = "") const {}
};

template <class T>
struct Col : Mat<T> {
  Col() {}
  explicit Col(uword n) : Mat<T>(n, 1) {}
  Col(const Mat<T>& m) : Mat<T>(m) { this->n_rows = this->n_elem; this->n_cols = this->n_elem ? 1 : 0; }
  template <class V, class = typename std::enable_if<is_arma<V>::value && !std::is_base_of<Mat<T>, V>::value>::type>
  Col(const V& v) : Col(Mat<T>(v.eval())) {}
  Col<T>& operator=(const Mat<T>& m) { Mat<T>::operator=(m); this->n_rows = this->n_elem; this->n_cols = this->n_elem ? 1 : 0; return *this; }
  Col<T>& operator=(const Col<T>& m) { return operator=(static_cast<const Mat<T>&>(m)); }
  Col<T>& operator=(T s) { Mat<T> o(1, 1); o.mem[0] = s; return operator=(o); }
  template <class V, class = typename std::enable_if<is_arma<V>::value && !std::is_base_of<Mat<T>, V>::value>::type>
  Col<T>& operator=(const V& v) { return operator=(Mat<T>(v.eval())); }
  using Mat<T>::rows;
  using Mat<T>::operator();
};
template <class T>
struct Row : Mat<T> {
  Row() {}
  explicit Row(uword n) : Mat<T>(1, n) {}
  Row(const Mat<T>& m) : Mat<T>(m) { this->n_cols = this->n_elem; this->n_rows = this->n_elem ? 1 : 0; }
  template <class V, class = typename std::enable_if<is_arma<V>::value && !std::is_base_of<Mat<T>, V>::value>::type>
  Row(const V& v) : Row(Mat<T>(v.eval())) {}
  Row<T>& operator=(const Mat<T>& m) { Mat<T>::operator=(m); this->n_cols = this->n_elem; this->n_rows = this->n_elem ? 1 : 0; return *this; }
  Row<T>& operator=(const Row<T>& m) { return operator=(static_cast<const Mat<T>&>(m)); }
  template <class V, class = typename std::enable_if<is_arma<V>::value && !std::is_base_of<Mat<T>, V>::value>::type>
  Row<T>& operator=(const V& v) { return operator=(Mat<T>(v.eval())); }
};

typedef Mat<double> mat;
typedef Col<double> vec;
typedef Col<double> colvec;
typedef Row<double> rowvec;
typedef Mat<uword> umat;
typedef Col<uword> uvec;
typedef Mat<sword> imat;
typedef Col<sword> ivec;

// ---- contiguous rectangular view
template <class T>
struct subview : arma_tag {
  typedef T elem_type;
  Mat<T>* m;
  uword r0, c0, n_rows, n_cols, n_elem;
  subview(Mat<T>* m_, uword r0_, uword c0_, uword nr, uword nc) : m(m_), r0(r0_), c0(c0_), n_rows(nr), n_cols(nc), n_elem(nr * nc) {
    if (r0 + nr > m->n_rows || c0 + nc > m->n_cols) throw std::logic_error("subview: indices out of bounds");
  }
  Mat<T> eval() const {
    Mat<T> o(n_rows, n_cols);
    for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) o(i, j) = (*m)(r0 + i, c0 + j);
    return o;
  }
  T& operator()(uword i) { return (n_rows == 1) ? (*m)(r0, c0 + i) : (*m)(r0 + i % n_rows, c0 + i / n_rows); }
  T operator()(uword i) const { return (n_rows == 1) ? (*m)(r0, c0 + i) : (*m)(r0 + i % n_rows, c0 + i / n_rows); }
  T& operator()(uword i, uword j) { return (*m)(r0 + i, c0 + j); }
  template <class V> typename std::enable_if<is_arma<V>::value, subview<T>&>::type operator=(const V& v) {
    const Mat<T> b = v.eval();
    if (b.n_elem != n_elem) throw std::logic_error("subview=: size mismatch");
    for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) (*m)(r0 + i, c0 + j) = b.mem[i + j * n_rows];
    return *this;
  }
  subview<T>& operator=(const subview<T>& v) { return operator=<subview<T>>(v); }
  template <class V> typename std::enable_if<is_arma<V>::value, subview<T>&>::type operator+=(const V& v) {
    const Mat<T> b = v.eval();
    for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) (*m)(r0 + i, c0 + j) += b.mem[i + j * n_rows];
    return *this;
  }
  template <class V> typename std::enable_if<is_arma<V>::value, subview<T>&>::type operator-=(const V& v) {
    const Mat<T> b = v.eval();
    for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) (*m)(r0 + i, c0 + j) -= b.mem[i + j * n_rows];
    return *this;
  }
  subview<T>& operator*=(T s) { for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) (*m)(r0 + i, c0 + j) *= s; return *this; }
  subview<T>& operator+=(T s) { for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) (*m)(r0 + i, c0 + j) += s; return *this; }
  void fill(T v) { for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) (*m)(r0 + i, c0 + j) = v; }
  Mat<T> t() const { return eval().t(); }
  Mat<T> subvec(uword a, uword b) const { Mat<T> e = eval(); Mat<T> o = (e.n_rows == 1) ? Mat<T>(1, b - a + 1) : Mat<T>(b - a + 1, 1); for (uword i = a; i <= b; i++) o.mem[i - a] = e.mem[i]; return o; }
  Mat<T> cols(const Mat<uword>& ix) const { Mat<T> e = eval(); return Mat<T>(e.cols(ix)); }
  Mat<T> rows(const Mat<uword>& ix) const { Mat<T> e = eval(); return Mat<T>(e.rows(ix)); }
};

// ---- gather / scatter view: rows(uvec), cols(uvec), elem(uvec), submat(uvec, uvec)
template <class T>
struct subview_elem : arma_tag {
  typedef T elem_type;
  Mat<T>* m;
  int kind;  // 0 rows, 1 cols, 2 elem, 3 submat
  std::vector<uword> r, c;
  uword n_rows = 0, n_cols = 0, n_elem = 0;
  subview_elem(Mat<T>* m_, int kind_, const Mat<uword>& a, const Mat<uword>* b = nullptr) : m(m_), kind(kind_), r(a.mem) {
    if (b) c = b->mem;
    if (kind == 0) { n_rows = r.size(); n_cols = m->n_cols; for (auto v : r) if (v >= m->n_rows) throw std::logic_error("rows(): index out of bounds"); }
    if (kind == 1) { n_rows = m->n_rows; n_cols = r.size(); for (auto v : r) if (v >= m->n_cols) throw std::logic_error("cols(): index out of bounds"); }
    if (kind == 2) { n_rows = r.size(); n_cols = 1; for (auto v : r) if (v >= m->n_elem) throw std::logic_error("elem(): index out of bounds"); }
    if (kind == 3) { n_rows = r.size(); n_cols = c.size(); }
    n_elem = n_rows * n_cols;
  }
  inline T& ref(uword i, uword j) const {
    switch (kind) {
      case 0: return (*m)(r[i], j);
      case 1: return (*m)(i, r[j]);
      case 2: return m->mem[r[i]];
      default: return (*m)(r[i], c[j]);
    }
  }
  Mat<T> eval() const {
    Mat<T> o(n_rows, n_cols);
    for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) o(i, j) = ref(i, j);
    return o;
  }
  template <class V> typename std::enable_if<is_arma<V>::value, subview_elem<T>&>::type operator=(const V& v) {
    const Mat<T> b = v.eval();
    if (b.n_elem != n_elem) throw std::logic_error("subview_elem=: size mismatch");
    for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) ref(i, j) = b.mem[i + j * n_rows];
    return *this;
  }
  subview_elem<T>& operator=(const subview_elem<T>& v) { return operator=<subview_elem<T>>(v); }
  template <class V> typename std::enable_if<is_arma<V>::value, subview_elem<T>&>::type operator+=(const V& v) {
    const Mat<T> b = v.eval();
    for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) ref(i, j) += b.mem[i + j * n_rows];
    return *this;
  }
  void fill(T v) { for (uword j = 0; j < n_cols; j++) for (uword i = 0; i < n_rows; i++) ref(i, j) = v; }
  Mat<T> t() const { return eval().t(); }
};

template <class T>
struct diagview : arma_tag {
  typedef T elem_type;
  Mat<T>* m;
  uword n_rows, n_cols = 1, n_elem;
  explicit diagview(Mat<T>* m_) : m(m_), n_rows(std::min(m_->n_rows, m_->n_cols)), n_elem(n_rows) {}
  Mat<T> eval() const { Mat<T> o(n_elem, 1); for (uword i = 0; i < n_elem; i++) o.mem[i] = (*m)(i, i); return o; }
  template <class V> typename std::enable_if<is_arma<V>::value, diagview<T>&>::type operator+=(const V& v) {
    const Mat<T> b = v.eval();
    if (b.n_elem != n_elem) throw std::logic_error("diag+=: size mismatch");
    for (uword i = 0; i < n_elem; i++) (*m)(i, i) += b.mem[i];
    return *this;
  }
  diagview<T>& operator+=(T s) { for (uword i = 0; i < n_elem; i++) (*m)(i, i) += s; return *this; }
  template <class V> typename std::enable_if<is_arma<V>::value, diagview<T>&>::type operator=(const V& v) {
    const Mat<T> b = v.eval();
    for (uword i = 0; i < n_elem; i++) (*m)(i, i) = b.mem[i];
    return *this;
  }
};

#define ARMA_CM(T) const_cast<Mat<T>*>(this)
template <class T> subview<T> Mat<T>::rows(uword a, uword b) { return subview<T>(this, a, 0, b - a + 1, n_cols); }
template <class T> subview<T> Mat<T>::cols(uword a, uword b) { return subview<T>(this, 0, a, n_rows, b - a + 1); }
template <class T> subview<T> Mat<T>::row(uword i) { return subview<T>(this, i, 0, 1, n_cols); }
template <class T> subview<T> Mat<T>::col(uword j) { return subview<T>(this, 0, j, n_rows, 1); }
template <class T> subview<T> Mat<T>::submat(uword r1, uword c1, uword r2, uword c2) { return subview<T>(this, r1, c1, r2 - r1 + 1, c2 - c1 + 1); }
template <class T> subview<T> Mat<T>::subvec(uword a, uword b) { return (n_rows == 1 && n_cols > 1) ? subview<T>(this, 0, a, 1, b - a + 1) : subview<T>(this, a, 0, b - a + 1, 1); }
template <class T> const subview<T> Mat<T>::rows(uword a, uword b) const { return ARMA_CM(T)->rows(a, b); }
template <class T> const subview<T> Mat<T>::cols(uword a, uword b) const { return ARMA_CM(T)->cols(a, b); }
template <class T> const subview<T> Mat<T>::row(uword i) const { return ARMA_CM(T)->row(i); }
template <class T> const subview<T> Mat<T>::col(uword j) const { return ARMA_CM(T)->col(j); }
template <class T> const subview<T> Mat<T>::submat(uword r1, uword c1, uword r2, uword c2) const { return ARMA_CM(T)->submat(r1, c1, r2, c2); }
template <class T> const subview<T> Mat<T>::subvec(uword a, uword b) const { return ARMA_CM(T)->subvec(a, b); }
template <class T> subview_elem<T> Mat<T>::rows(const Mat<uword>& ix) { return subview_elem<T>(this, 0, ix); }
template <class T> subview_elem<T> Mat<T>::cols(const Mat<uword>& ix) { return subview_elem<T>(this, 1, ix); }
template <class T> subview_elem<T> Mat<T>::elem(const Mat<uword>& ix) { return subview_elem<T>(this, 2, ix); }
template <class T> subview_elem<T> Mat<T>::submat(const Mat<uword>& r, const Mat<uword>& c) { return subview_elem<T>(this, 3, r, &c); }
template <class T> subview_elem<T> Mat<T>::operator()(const Mat<uword>& r, const Mat<uword>& c) { return subview_elem<T>(this, 3, r, &c); }
template <class T> const subview_elem<T> Mat<T>::rows(const Mat<uword>& ix) const { return ARMA_CM(T)->rows(ix); }
template <class T> const subview_elem<T> Mat<T>::cols(const Mat<uword>& ix) const { return ARMA_CM(T)->cols(ix); }
template <class T> const subview_elem<T> Mat<T>::elem(const Mat<uword>& ix) const { return ARMA_CM(T)->elem(ix); }
template <class T> const subview_elem<T> Mat<T>::submat(const Mat<uword>& r, const Mat<uword>& c) const { return ARMA_CM(T)->submat(r, c); }
template <class T> diagview<T> Mat<T>::diag() { return diagview<T>(this); }
template <class T> const diagview<T> Mat<T>::diag() const { return diagview<T>(ARMA_CM(T)); }

// ---- cube
template <class T>
struct Cube : arma_tag {
  uword n_rows = 0, n_cols = 0, n_slices = 0, n_elem = 0;
  std::vector<Mat<T>> sl;
  Cube() {}
  Cube(uword r, uword c, uword s) : n_rows(r), n_cols(c), n_slices(s), n_elem(r * c * s), sl(s, Mat<T>(r, c)) {}
  Mat<T>& slice(uword s) { return sl.at(s); }
  const Mat<T>& slice(uword s) const { return sl.at(s); }
  Mat<T> row(uword i) const {  // n_cols x n_slices (as arma: a cube row becomes a matrix)
    Mat<T> o(n_cols, n_slices);
    for (uword s = 0; s < n_slices; s++) for (uword j = 0; j < n_cols; j++) o(j, s) = sl[s](i, j);
    return o;
  }
  struct subcube_view {
    Cube<T>* c; uword r, s;
    template <class V> subcube_view& operator=(const V& v) { const Mat<T> b = v.eval(); for (uword j = 0; j < c->n_cols; j++) c->sl[s](r, j) = b.mem[j]; return *this; }
  };
  subcube_view subcube(uword r1, uword, uword s1, uword, uword, uword) { return subcube_view{this, r1, s1}; }
  struct col_view {  // cube.col(j): an n_rows x n_slices matrix view (as arma)
    Cube<T>* c; uword j;
    template <class V> col_view& operator=(const V& v) {
      const Mat<T> b = v.eval();
      if (b.n_rows != c->n_rows || b.n_cols != c->n_slices) throw std::logic_error("cube.col(): size mismatch");
      for (uword s = 0; s < c->n_slices; s++) for (uword i = 0; i < c->n_rows; i++) c->sl[s](i, j) = b(i, s);
      return *this;
    }
  };
  col_view col(uword j) { return col_view{this, j}; }
  T& operator()(uword i, uword j, uword s) { return sl[s](i, j); }
};
typedef Cube<double> cube;

// ---- field
template <class T>
struct field {
  uword n_rows = 0, n_cols = 0, n_elem = 0;
  std::vector<T> mem;
  field() {}
  explicit field(uword n) : n_rows(n), n_cols(1), n_elem(n), mem(n) {}
  field(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c) {}
  T& operator()(uword i) { return mem.at(i); }
  const T& operator()(uword i) const { return mem.at(i); }
  T& operator()(uword i, uword j) { return mem.at(i + j * n_rows); }
  const T& operator()(uword i, uword j) const { return mem.at(i + j * n_rows); }
};

// ---- sparse placeholder (the live path never touches it)
struct sp_mat : arma_tag {
  typedef double elem_type;
  uword n_rows = 0, n_cols = 0;
  mat dense;
  sp_mat() {}
  mat eval() const { return dense; }
  subview<double> col(uword j) { return dense.col(j); }
  sp_mat t() const { sp_mat o; o.dense = dense.t(); o.n_rows = n_cols; o.n_cols = n_rows; return o; }
};

// ---- generators
inline mat zeros(uword n) { return mat(n, 1); }
inline mat zeros(uword r, uword c) { return mat(r, c); }
inline cube zeros(uword r, uword c, uword s) { return cube(r, c, s); }
inline mat zeros(const SizeMat& s) { return mat(s.n_rows, s.n_cols); }
template <class V> V zeros(uword n) { V v; static_cast<Mat<typename V::elem_type>&>(v) = Mat<typename V::elem_type>(n, 1); return V(static_cast<Mat<typename V::elem_type>&>(v)); }
template <class V> V zeros(uword r, uword c) { return V(Mat<typename V::elem_type>(r, c)); }
inline mat ones(uword n) { mat m(n, 1); m.fill(1.0); return m; }
inline mat ones(uword r, uword c) { mat m(r, c); m.fill(1.0); return m; }
template <class V> V ones(uword n) { Mat<typename V::elem_type> m(n, 1); m.fill(1); return V(m); }
template <class V> V ones(uword r, uword c) { Mat<typename V::elem_type> m(r, c); m.fill(1); return V(m); }
inline mat eye(uword r, uword c) { mat m(r, c); for (uword i = 0; i < std::min(r, c); i++) m(i, i) = 1.0; return m; }
template <class V> V regspace(sword a, sword b) {
  Mat<typename V::elem_type> m(b >= a ? (uword)(b - a + 1) : 0, 1);
  for (uword i = 0; i < m.n_elem; i++) m.mem[i] = (typename V::elem_type)(a + (sword)i);
  return V(m);
}
template <class A> SizeMat size(const A& a) { return SizeMat{a.n_rows, a.n_cols}; }

// randn / randu: the driver can inject the next normal vector (so that the reference's internal arma::randn calls see
// exactly the numbers the oracle was given); otherwise a host stream supplied by the driver is used
struct RngHooks {
  std::function<double()> norm, unif;
  std::function<double(double, double)> gamma;
  std::vector<double> injected;  // consumed by the next randn(n) with n == injected.size()
};
inline RngHooks& rng_hooks() { static RngHooks h; return h; }
inline mat randn(uword n) {
  RngHooks& h = rng_hooks();
  mat m(n, 1);
  if (h.injected.size() == n && n > 0) { m.mem = h.injected; h.injected.clear(); return m; }
  for (uword i = 0; i < n; i++) m.mem[i] = h.norm ? h.norm() : 0.0;
  return m;
}
inline mat randn(uword r, uword c) { mat m = randn(r * c); m.n_rows = r; m.n_cols = c; return m; }
inline double randu() { RngHooks& h = rng_hooks(); return h.unif ? h.unif() : 0.5; }

// ---- element-wise and scalar operators (result is always a Mat)
#define ARMA_BINOP(OP, NAME)                                                                                              \
  template <class A, class B>                                                                                              \
  typename std::enable_if<is_arma<A>::value && is_arma<B>::value, Mat<typename A::elem_type>>::type operator OP(const A& a, const B& b) { \
    const Mat<typename A::elem_type> x = a.eval();                                                                         \
    const Mat<typename A::elem_type> y = b.eval();                                                                         \
    if (x.n_rows != y.n_rows || x.n_cols != y.n_cols) {                                                                    \
      if (x.n_elem != y.n_elem) throw std::logic_error(NAME ": size mismatch");                                            \
    }                                                                                                                      \
    Mat<typename A::elem_type> o(x.n_rows, x.n_cols);                                                                      \
    for (uword i = 0; i < x.n_elem; i++) o.mem[i] = x.mem[i] OP y.mem[i];                                                  \
    return o;                                                                                                              \
  }
ARMA_BINOP(+, "addition")
ARMA_BINOP(-, "subtraction")
ARMA_BINOP(/, "element-wise division")
template <class A, class B>
typename std::enable_if<is_arma<A>::value && is_arma<B>::value, Mat<typename A::elem_type>>::type operator%(const A& a, const B& b) {
  const Mat<typename A::elem_type> x = a.eval();
  const Mat<typename A::elem_type> y = b.eval();
  if (x.n_elem != y.n_elem) throw std::logic_error("element-wise multiplication: size mismatch");
  Mat<typename A::elem_type> o(x.n_rows, x.n_cols);
  for (uword i = 0; i < x.n_elem; i++) o.mem[i] = x.mem[i] * y.mem[i];
  return o;
}
// matrix product
template <class A, class B>
typename std::enable_if<is_arma<A>::value && is_arma<B>::value, Mat<typename A::elem_type>>::type operator*(const A& a, const B& b) {
  typedef typename A::elem_type T;
  const Mat<T> x = a.eval();
  const Mat<T> y = b.eval();
  if (x.n_cols != y.n_rows) {
    if (x.n_elem == 1) { Mat<T> o = y; for (auto& v : o.mem) v *= x.mem[0]; return o; }
    if (y.n_elem == 1) { Mat<T> o = x; for (auto& v : o.mem) v *= y.mem[0]; return o; }
    throw std::logic_error("matrix multiplication: incompatible dimensions");
  }
  Mat<T> o(x.n_rows, y.n_cols);
  for (uword j = 0; j < y.n_cols; j++)
    for (uword k = 0; k < x.n_cols; k++) {
      const T bkj = y(k, j);
      const T* xp = &x.mem[k * x.n_rows];
      T* op = &o.mem[j * o.n_rows];
      for (uword i = 0; i < x.n_rows; i++) op[i] += xp[i] * bkj;
    }
  return o;
}
#define ARMA_SCALAR_OPS(OP)                                                                                               \
  template <class A, class S>                                                                                              \
  typename std::enable_if<is_arma<A>::value && std::is_arithmetic<S>::value, Mat<typename A::elem_type>>::type operator OP(const A& a, S s) { \
    Mat<typename A::elem_type> o = a.eval();                                                                               \
    for (auto& v : o.mem) v = v OP (typename A::elem_type)s;                                                               \
    return o;                                                                                                              \
  }                                                                                                                        \
  template <class A, class S>                                                                                              \
  typename std::enable_if<is_arma<A>::value && std::is_arithmetic<S>::value, Mat<typename A::elem_type>>::type operator OP(S s, const A& a) { \
    Mat<typename A::elem_type> o = a.eval();                                                                               \
    for (auto& v : o.mem) v = (typename A::elem_type)s OP v;                                                               \
    return o;                                                                                                              \
  }
ARMA_SCALAR_OPS(+)
ARMA_SCALAR_OPS(-)
ARMA_SCALAR_OPS(*)
ARMA_SCALAR_OPS(/)
template <class A> typename std::enable_if<is_arma<A>::value, Mat<typename A::elem_type>>::type operator-(const A& a) {
  Mat<typename A::elem_type> o = a.eval();
  for (auto& v : o.mem) v = -v;
  return o;
}
#define ARMA_CMP(OP)                                                                                                       \
  template <class A, class S>                                                                                              \
  typename std::enable_if<is_arma<A>::value && std::is_arithmetic<S>::value, umat>::type operator OP(const A& a, S s) {    \
    const Mat<typename A::elem_type> x = a.eval();                                                                         \
    umat o(x.n_rows, x.n_cols);                                                                                            \
    for (uword i = 0; i < x.n_elem; i++) o.mem[i] = (x.mem[i] OP (typename A::elem_type)s) ? 1 : 0;                        \
    return o;                                                                                                              \
  }
ARMA_CMP(==)
ARMA_CMP(!=)
ARMA_CMP(<)
ARMA_CMP(>)
ARMA_CMP(<=)
ARMA_CMP(>=)

#define ARMA_MAPFN(NAME, EXPR)                                                                                             \
  template <class A> typename std::enable_if<is_arma<A>::value, mat>::type NAME(const A& a) {                              \
    mat o = a.eval();                                                                                                      \
    for (auto& x : o.mem) x = EXPR;                                                                                        \
    return o;                                                                                                              \
  }
ARMA_MAPFN(exp, std::exp(x))
ARMA_MAPFN(log, std::log(x))
ARMA_MAPFN(sqrt, std::sqrt(x))
ARMA_MAPFN(abs, std::fabs(x))
template <class A> typename std::enable_if<is_arma<A>::value, mat>::type pow(const A& a, double p) {
  mat o = a.eval();
  for (auto& x : o.mem) x = std::pow(x, p);
  return o;
}

// ---- reductions, searches, reshapes
template <class A> typename std::enable_if<is_arma<A>::value, typename A::elem_type>::type accu(const A& a) {
  const Mat<typename A::elem_type> x = a.eval();
  typename A::elem_type s = 0;
  for (auto v : x.mem) s += v;
  return s;
}
template <class A> typename std::enable_if<is_arma<A>::value, mat>::type sum(const A& a, int dim) {
  const mat x = a.eval();
  if (dim == 0) { mat o(1, x.n_cols); for (uword j = 0; j < x.n_cols; j++) for (uword i = 0; i < x.n_rows; i++) o.mem[j] += x(i, j); return o; }
  mat o(x.n_rows, 1);
  for (uword j = 0; j < x.n_cols; j++) for (uword i = 0; i < x.n_rows; i++) o.mem[i] += x(i, j);
  return o;
}
inline mat sum(const cube& c, int dim) {
  if (dim != 2) throw std::logic_error("sum(cube): only dim = 2 is implemented");
  mat o(c.n_rows, c.n_cols);
  for (uword s = 0; s < c.n_slices; s++) for (uword i = 0; i < o.n_elem; i++) o.mem[i] += c.sl[s].mem[i];
  return o;
}
template <class A> typename std::enable_if<is_arma<A>::value, double>::type mean(const A& a) { const mat x = a.eval(); return x.n_elem ? accu(x) / x.n_elem : 0.0; }
template <class A> typename std::enable_if<is_arma<A>::value, typename A::elem_type>::type max(const A& a) { return Mat<typename A::elem_type>(a.eval()).max(); }
template <class A> typename std::enable_if<is_arma<A>::value, double>::type norm(const A& a) {
  const mat x = a.eval();
  double s = 0;
  for (auto v : x.mem) s += v * v;
  return std::sqrt(s);
}
template <class A> typename std::enable_if<is_arma<A>::value, uvec>::type find(const A& a, uword k = 0, const char* = "first") {
  const Mat<typename A::elem_type> x = a.eval();
  std::vector<uword> ix;
  for (uword i = 0; i < x.n_elem; i++) if (x.mem[i] != 0) { ix.push_back(i); if (k && ix.size() >= k) break; }
  umat o(ix.size(), 1);
  o.mem = ix;
  return uvec(o);
}
template <class A> typename std::enable_if<is_arma<A>::value, uvec>::type find_finite(const A& a) {
  const mat x = a.eval();
  std::vector<uword> ix;
  for (uword i = 0; i < x.n_elem; i++) if (std::isfinite(x.mem[i])) ix.push_back(i);
  umat o(ix.size(), 1);
  o.mem = ix;
  return uvec(o);
}
template <class A> typename std::enable_if<is_arma<A>::value, uvec>::type find_nonfinite(const A& a) {
  const mat x = a.eval();
  std::vector<uword> ix;
  for (uword i = 0; i < x.n_elem; i++) if (!std::isfinite(x.mem[i])) ix.push_back(i);
  umat o(ix.size(), 1);
  o.mem = ix;
  return uvec(o);
}
inline bool is_finite(double x) { return std::isfinite(x); }
template <class A> typename std::enable_if<is_arma<A>::value, Mat<typename A::elem_type>>::type unique(const A& a) {
  Mat<typename A::elem_type> x = a.eval();
  std::sort(x.mem.begin(), x.mem.end());
  x.mem.erase(std::unique(x.mem.begin(), x.mem.end()), x.mem.end());
  x.n_rows = x.n_elem = x.mem.size();
  x.n_cols = x.n_elem ? 1 : 0;
  return x;
}
template <class A, class B> typename std::enable_if<is_arma<A>::value && is_arma<B>::value, Mat<typename A::elem_type>>::type intersect(const A& a, const B& b) {
  Mat<typename A::elem_type> x = unique(a), y = unique(b), o;
  std::set_intersection(x.mem.begin(), x.mem.end(), y.mem.begin(), y.mem.end(), std::back_inserter(o.mem));
  o.n_rows = o.n_elem = o.mem.size();
  o.n_cols = o.n_elem ? 1 : 0;
  return o;
}
template <class A> typename std::enable_if<is_arma<A>::value, Mat<typename A::elem_type>>::type cumsum(const A& a) {
  Mat<typename A::elem_type> x = a.eval();
  for (uword i = 1; i < x.n_elem; i++) x.mem[i] += x.mem[i - 1];
  return x;
}
template <class A> typename std::enable_if<is_arma<A>::value, Mat<typename A::elem_type>>::type trans(const A& a) { return Mat<typename A::elem_type>(a.eval()).t(); }
template <class A> typename std::enable_if<is_arma<A>::value, Mat<typename A::elem_type>>::type vectorise(const A& a) {
  Mat<typename A::elem_type> x = a.eval();
  x.n_rows = x.n_elem;
  x.n_cols = x.n_elem ? 1 : 0;
  return x;
}
template <class A, class B> typename std::enable_if<is_arma<A>::value && is_arma<B>::value, Mat<typename A::elem_type>>::type join_vert(const A& a, const B& b) {
  const Mat<typename A::elem_type> x = a.eval();
  const Mat<typename A::elem_type> y = b.eval();
  if (x.n_elem == 0) return y;
  if (y.n_elem == 0) return x;
  if (x.n_cols != y.n_cols) throw std::logic_error("join_vert: column mismatch");
  Mat<typename A::elem_type> o(x.n_rows + y.n_rows, x.n_cols);
  for (uword j = 0; j < x.n_cols; j++) {
    for (uword i = 0; i < x.n_rows; i++) o(i, j) = x(i, j);
    for (uword i = 0; i < y.n_rows; i++) o(x.n_rows + i, j) = y(i, j);
  }
  return o;
}
template <class A> typename std::enable_if<is_arma<A>::value, mat>::type repmat(const A& a, uword r, uword c) {
  const mat x = a.eval();
  mat o(x.n_rows * r, x.n_cols * c);
  for (uword j = 0; j < o.n_cols; j++) for (uword i = 0; i < o.n_rows; i++) o(i, j) = x(i % x.n_rows, j % x.n_cols);
  return o;
}
template <class A> typename std::enable_if<is_arma<A>::value, mat>::type diagmat(const A& a) {
  const mat x = a.eval();
  mat o(x.n_elem, x.n_elem);
  for (uword i = 0; i < x.n_elem; i++) o(i, i) = x.mem[i];
  return o;
}
template <class A> typename std::enable_if<is_arma<A>::value, mat>::type symmatu(const A& a) {
  mat x = a.eval();
  for (uword j = 0; j < x.n_cols; j++) for (uword i = j + 1; i < x.n_rows; i++) x(i, j) = x(j, i);
  return x;
}
template <class A> typename std::enable_if<is_arma<A>::value, mat>::type symmatl(const A& a) {
  mat x = a.eval();
  for (uword j = 0; j < x.n_cols; j++) for (uword i = 0; i < j; i++) x(i, j) = x(j, i);
  return x;
}

// ---- factorisations (dpotrf / dtrtri / dpotri semantics)
template <class A> typename std::enable_if<is_arma<A>::value, mat>::type chol(const A& a, const char* layout = "upper") {
  mat x = a.eval();
  const uword n = x.n_rows;
  if (n != x.n_cols) throw std::runtime_error("chol(): given matrix must be square sized");
  const bool lower = layout[0] == 'l';
  if (!lower) x = x.t();  // factor the lower form of the transposed matrix, transpose back
  for (uword j = 0; j < n; j++) {
    double ajj = x(j, j);
    for (uword k = 0; k < j; k++) ajj -= x(j, k) * x(j, k);
    if (!(ajj > 0.0) || !std::isfinite(ajj)) throw std::runtime_error("chol(): decomposition failed");
    ajj = std::sqrt(ajj);
    x(j, j) = ajj;
    for (uword i = j + 1; i < n; i++) {
      double s = x(i, j);
      for (uword k = 0; k < j; k++) s -= x(i, k) * x(j, k);
      x(i, j) = s / ajj;
    }
  }
  for (uword j = 0; j < n; j++) for (uword i = 0; i < j; i++) x(i, j) = 0.0;
  return lower ? x : x.t();
}
struct trimatl_proxy : arma_tag { typedef double elem_type; mat m; mat eval() const { return m; } };
template <class A> typename std::enable_if<is_arma<A>::value, trimatl_proxy>::type trimatl(const A& a) {
  trimatl_proxy p;
  p.m = a.eval();
  for (uword j = 0; j < p.m.n_cols; j++) for (uword i = 0; i < j && i < p.m.n_rows; i++) p.m(i, j) = 0.0;
  return p;
}
inline mat inv(const trimatl_proxy& p) {
  const mat& L = p.m;
  const uword n = L.n_rows;
  mat X(n, n);
  for (uword j = 0; j < n; j++) {
    if (L(j, j) == 0.0) throw std::runtime_error("inv(): matrix is singular");
    X(j, j) = 1.0 / L(j, j);
    for (uword i = j + 1; i < n; i++) {
      double s = 0;
      for (uword k = j; k < i; k++) s += L(i, k) * X(k, j);
      X(i, j) = -s / L(i, i);
    }
  }
  return X;
}
template <class A> typename std::enable_if<is_arma<A>::value && !std::is_same<A, trimatl_proxy>::value, mat>::type inv_sympd(const A& a) {
  mat Li = inv(trimatl(chol(a, "lower")));
  return Li.t() * Li;
}

template <class T> struct conv_to;
template <> struct conv_to<double> { template <class A> static double from(const A& a) { const mat x = a.eval(); if (x.n_elem != 1) throw std::logic_error("conv_to<double>: not 1x1"); return x.mem[0]; } };
template <> struct conv_to<int> { template <class A> static int from(const A& a) { const Mat<typename A::elem_type> x = a.eval(); if (x.n_elem != 1) throw std::logic_error("conv_to<int>: not 1x1"); return (int)x.mem[0]; } };
template <class T> struct conv_to<Col<T>> { template <class A> static Col<T> from(const A& a) { const Mat<typename A::elem_type> x = a.eval(); Mat<T> o(x.n_elem, 1); for (uword i = 0; i < x.n_elem; i++) o.mem[i] = (T)x.mem[i]; return Col<T>(o); } };
template <class T> struct conv_to<Mat<T>> { template <class A> static Mat<T> from(const A& a) { const Mat<typename A::elem_type> x = a.eval(); Mat<T> o(x.n_rows, x.n_cols); for (uword i = 0; i < x.n_elem; i++) o.mem[i] = (T)x.mem[i]; return o; } };
template <> struct conv_to<sp_mat> { template <class A> static sp_mat from(const A& a) { sp_mat s; s.dense = a.eval(); s.n_rows = s.dense.n_rows; s.n_cols = s.dense.n_cols; return s; } };

template <class T> std::ostream& operator<<(std::ostream& o, const Mat<T>& m) {
  for (uword i = 0; i < m.n_rows; i++) { for (uword j = 0; j < m.n_cols; j++) o << m(i, j) << " "; o << "\n"; }
  return o;
}
template <class T> std::ostream& operator<<(std::ostream& o, const subview<T>& m) { return o << m.eval(); }

}  // namespace arma

// ------------------------------------------------------------------------------------------------ Rcpp / R stand-ins
namespace Rcpp {
struct NullStream {
  template <class T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
static NullStream Rcout;
struct RNGScope {};
inline void stop(const std::string& msg) { throw std::runtime_error(msg); }
inline void checkUserInterrupt() {}
struct StringMatrix {
  int r, c;
  std::vector<std::string> s;
  StringMatrix(int r_, int c_) : r(r_), c(c_), s((size_t)r_ * c_) {}
  std::string& operator()(int i, int j) { return s[i + (size_t)j * r]; }
};
struct RObject {};
struct NamedValue {
  std::string name;
  std::any value;  // a copy of what was assigned (arma::field<...>, arma::mat, double ...)
  template <class T> NamedValue& operator=(const T& v) { value = v; return *this; }
};
inline NamedValue Named(const std::string& n) { return NamedValue{n, std::any()}; }
struct List {  // retains its fields so that the driver can read what the reference returned
  std::map<std::string, std::any> fields;
  template <class... A> static List create(const A&... a) { List l; (l.fields.emplace(a.name, a.value), ...); return l; }
  template <class T> const T& get(const std::string& n) const { return std::any_cast<const T&>(fields.at(n)); }
  bool has(const std::string& n) const { return fields.count(n) != 0; }
};
}  // namespace Rcpp
namespace R {
inline double runif(double a, double b) { auto& h = arma::rng_hooks(); return a + (b - a) * (h.unif ? h.unif() : 0.5); }
inline double rgamma(double shape, double scale) { auto& h = arma::rng_hooks(); return h.gamma ? h.gamma(shape, scale) : shape * scale; }
}  // namespace R
inline void Rprintf(const char*, ...) {}
// interrupt_handler.h: never interrupted
#ifndef FALSE
#define FALSE 0
#endif
inline void R_CheckUserInterrupt() {}
inline int R_ToplevelExec(void (*fn)(void*), void* d) { fn(d); return 1; }
