// shim_eigen.h — hand-written stand-in for the handful of Eigen 3.3 types/members the reference's hot-path
// sources use.  TEST INFRASTRUCTURE (oracle/_ref build only).  Eigen itself is not installed in this image
// and is not vendored by the reference; nothing here is copied from Eigen.
//
// Arithmetic pins (these are the assumptions DESIGN.md "Floating-point pins" documents; the shim makes
// them explicit instead of inheriting them from an Eigen build):
//   squaredNorm(): sequential left-to-right sum of squares == (x*x + y*y) + z*z for 3-vectors, which is what
//                  Eigen 3.3 evaluates with SSE2 packets (predux of the first packet, then the scalar tail)
//   norm()       : sqrt(squaredNorm())
//   normalized() : z = squaredNorm(); z > 0 ? v / sqrt(z) : v      (component-wise true division)
//   v * s, v + w : one rounding per component, no fused multiply-add (build uses -ffp-contract=off)
#pragma once
#include <cmath>
#include <cstddef>
#include <iostream>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {
constexpr int Dynamic = -1;

template <class T, int R, int C> class Matrix;
template <class T> class Quaternion;

template <class T, int R, int C>
struct CommaInit {
  Matrix<T, R, C>& m;
  int k;
  CommaInit& operator,(T v) { m.setRowMajor(k++, v); return *this; }
  // column blocks: `m << e1, e2, e3;` with R x 1 vectors places them side by side
  CommaInit& operator,(const Matrix<T, R, 1>& col) { for (int i = 0; i < R; ++i) m(i, k) = col(i); ++k; return *this; }
};

template <class T, int R, int C, int BR, int BC>
struct BlockRef {
  Matrix<T, R, C>& m;
  int i0, j0;
  BlockRef& operator=(const Matrix<T, BR, BC>& o) {
    for (int i = 0; i < BR; ++i)
      for (int j = 0; j < BC; ++j) m(i0 + i, j0 + j) = o(i, j);
    return *this;
  }
  operator Matrix<T, BR, BC>() const {
    Matrix<T, BR, BC> o;
    for (int i = 0; i < BR; ++i)
      for (int j = 0; j < BC; ++j) o(i, j) = m(i0 + i, j0 + j);
    return o;
  }
};

template <class T, int R, int C>
struct ArrayWrap {
  Matrix<T, R, C> v;
  ArrayWrap floor() const {
    ArrayWrap o;
    for (int i = 0; i < R * C; ++i) o.v.d_[i] = std::floor(v.d_[i]);
    return o;
  }
  template <class U>
  Matrix<U, R, C> cast() const {
    Matrix<U, R, C> o;
    for (int i = 0; i < R * C; ++i) o.d_[i] = (U)v.d_[i];
    return o;
  }
};

template <class T, int R, int C>
class Matrix {
 public:
  typedef T Scalar;
  T d_[R * C];  // column-major, like Eigen's default

  Matrix() {}
  Matrix(T a, T b) { static_assert(R * C == 2, "size"); d_[0] = a; d_[1] = b; }
  Matrix(T a, T b, T c) { static_assert(R * C == 3, "size"); d_[0] = a; d_[1] = b; d_[2] = c; }
  Matrix(T a, T b, T c, T d) { static_assert(R * C == 4, "size"); d_[0] = a; d_[1] = b; d_[2] = c; d_[3] = d; }
  Matrix& operator=(const Quaternion<T>& q);

  static Matrix Constant(T v) { Matrix m; for (int i = 0; i < R * C; ++i) m.d_[i] = v; return m; }
  static Matrix Zero() { return Constant(T(0)); }
  static Matrix Ones() { return Constant(T(1)); }
  static Matrix Identity() { Matrix m = Zero(); for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }

  void setRowMajor(int k, T v) { d_[(k % C) * R + (k / C)] = v; }
  CommaInit<T, R, C> operator<<(T v) { setRowMajor(0, v); return CommaInit<T, R, C>{*this, 1}; }
  template <int R2 = R, int C2 = C, class = typename std::enable_if<(C2 > 1)>::type>
  CommaInit<T, R, C> operator<<(const Matrix<T, R2, 1>& col) { for (int i = 0; i < R; ++i) (*this)(i, 0) = col(i); return CommaInit<T, R, C>{*this, 1}; }
  void normalize() { *this = normalized(); }

  T& operator()(int i) { return d_[i]; }
  const T& operator()(int i) const { return d_[i]; }
  T& operator[](int i) { return d_[i]; }
  const T& operator[](int i) const { return d_[i]; }
  T& operator()(int i, int j) { return d_[j * R + i]; }
  const T& operator()(int i, int j) const { return d_[j * R + i]; }
  T& x() { return d_[0]; } const T& x() const { return d_[0]; }
  T& y() { return d_[1]; } const T& y() const { return d_[1]; }
  T& z() { return d_[2]; } const T& z() const { return d_[2]; }
  T& w() { return d_[3]; } const T& w() const { return d_[3]; }
  int size() const { return R * C; }
  int rows() const { return R; }
  int cols() const { return C; }
  T* data() { return d_; }
  const T* data() const { return d_; }

  Matrix operator+(const Matrix& o) const { Matrix r; for (int i = 0; i < R * C; ++i) r.d_[i] = d_[i] + o.d_[i]; return r; }
  Matrix operator-(const Matrix& o) const { Matrix r; for (int i = 0; i < R * C; ++i) r.d_[i] = d_[i] - o.d_[i]; return r; }
  Matrix operator-() const { Matrix r; for (int i = 0; i < R * C; ++i) r.d_[i] = -d_[i]; return r; }
  Matrix operator*(T s) const { Matrix r; for (int i = 0; i < R * C; ++i) r.d_[i] = d_[i] * s; return r; }
  Matrix operator/(T s) const { Matrix r; for (int i = 0; i < R * C; ++i) r.d_[i] = d_[i] / s; return r; }
  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d_[i] += o.d_[i]; return *this; }
  Matrix& operator-=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d_[i] -= o.d_[i]; return *this; }
  Matrix& operator*=(T s) { for (int i = 0; i < R * C; ++i) d_[i] *= s; return *this; }
  bool operator==(const Matrix& o) const { for (int i = 0; i < R * C; ++i) if (!(d_[i] == o.d_[i])) return false; return true; }
  bool operator!=(const Matrix& o) const { return !(*this == o); }

  template <int K>
  Matrix<T, R, K> operator*(const Matrix<T, C, K>& o) const {
    Matrix<T, R, K> r;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < K; ++j) {
        T acc = (*this)(i, 0) * o(0, j);
        for (int k = 1; k < C; ++k) acc = acc + (*this)(i, k) * o(k, j);
        r(i, j) = acc;
      }
    return r;
  }

  T squaredNorm() const { T acc = d_[0] * d_[0]; for (int i = 1; i < R * C; ++i) acc = acc + d_[i] * d_[i]; return acc; }
  T norm() const { return std::sqrt(squaredNorm()); }
  Matrix normalized() const {
    T z = squaredNorm();
    if (z > T(0)) return *this / std::sqrt(z);
    return *this;
  }
  T dot(const Matrix& o) const { T acc = d_[0] * o.d_[0]; for (int i = 1; i < R * C; ++i) acc = acc + d_[i] * o.d_[i]; return acc; }
  Matrix cross(const Matrix& o) const {
    static_assert(R * C == 3, "cross");
    return Matrix(d_[1] * o.d_[2] - d_[2] * o.d_[1], d_[2] * o.d_[0] - d_[0] * o.d_[2], d_[0] * o.d_[1] - d_[1] * o.d_[0]);
  }
  Matrix<T, C, R> transpose() const { Matrix<T, C, R> r; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) r(j, i) = (*this)(i, j); return r; }
  Matrix inverse() const {
    static_assert(R == 3 && C == 3, "inverse: 3x3 only");
    const Matrix& a = *this;
    T det = a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
            a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
    Matrix r;
    r(0, 0) = (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) / det; r(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) / det;
    r(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) / det; r(1, 0) = (a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2)) / det;
    r(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) / det; r(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) / det;
    r(2, 0) = (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0)) / det; r(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) / det;
    r(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) / det;
    return r;
  }
  ArrayWrap<T, R, C> array() const { return ArrayWrap<T, R, C>{*this}; }
  template <int BR, int BC>
  BlockRef<T, R, C, BR, BC> block(int i, int j) { return BlockRef<T, R, C, BR, BC>{*this, i, j}; }
};

template <class T, int R, int C>
Matrix<T, R, C> operator*(T s, const Matrix<T, R, C>& m) { return m * s; }
template <class T, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& m) {
  for (int i = 0; i < R; ++i) { for (int j = 0; j < C; ++j) os << m(i, j) << (j + 1 < C ? " " : ""); if (i + 1 < R) os << "\n"; }
  return os;
}

// dynamic-size matrices only appear as never-touched data members (AlgorithmBase.hpp:84) and inside the
// reference's dead getCircunferenceRadius (geometry_utils.cpp:97-150): row-major storage is enough
template <class T>
class Matrix<T, Dynamic, Dynamic> {
 public:
  std::vector<T> d_;
  int r_ = 0, c_ = 0;
  Matrix() {}
  Matrix(int r, int c) : d_((size_t)r * c), r_(r), c_(c) {}
  T& operator()(int i, int j) { return d_[(size_t)i * c_ + j]; }
  const T& operator()(int i, int j) const { return d_[(size_t)i * c_ + j]; }
  struct Comma { Matrix& m; size_t k; Comma& operator,(T v) { m.d_[k++] = v; return *this; } };
  Comma operator<<(T v) { d_[0] = v; return Comma{*this, 1}; }
};
template <class M>
class FullPivLU {  // only instantiated by dead code; solves the 3x3 system by explicit inverse
 public:
  Matrix<double, 3, 3> a_;
  template <class X> explicit FullPivLU(const X& x) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a_(i, j) = x(i, j); }
  void setThreshold(double) {}
  Matrix<double, 3, 1> solve(const Matrix<double, 3, 1>& b) const { return a_.inverse() * b; }
};

template <class T>
class Quaternion {
 public:
  T w_, x_, y_, z_;
  Quaternion() {}
  Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}
  // matrix -> quaternion in the operation order of Eigen 3.3 / 3.4's quaternionbase_assign_impl<Other, 3, 3> (the published
  // algorithm: trace branch with t = sqrt(trace + 1), w = t / 2, t = 0.5 / t, differences TIMES t; else the largest
  // diagonal element i, j = (i + 1) % 3, k = (j + 1) % 3) — used by depthOdomCallback (grid_map.cpp:982)
  explicit Quaternion(const Matrix<T, 3, 3>& m) {
    T t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > T(0)) {
      t = std::sqrt(t + T(1.0));
      w_ = T(0.5) * t;
      t = T(0.5) / t;
      x_ = (m(2, 1) - m(1, 2)) * t;
      y_ = (m(0, 2) - m(2, 0)) * t;
      z_ = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + T(1.0));
      T v[3];
      v[i] = T(0.5) * t;
      t = T(0.5) / t;
      w_ = (m(k, j) - m(j, k)) * t;
      v[j] = (m(j, i) + m(i, j)) * t;
      v[k] = (m(k, i) + m(i, k)) * t;
      x_ = v[0]; y_ = v[1]; z_ = v[2];
    }
  }
  template <int R, int C>
  explicit Quaternion(const BlockRef<T, R, C, 3, 3>& b) : Quaternion((Matrix<T, 3, 3>)b) {}
  // QuaternionBase::toRotationMatrix in Eigen's operation order (tx = 2x, twx = tx * w, ...); bit-equal to the textbook
  // form 1 - 2 (yy + zz), 2 (xy - zw), ... because scaling by two is exact
  Matrix<T, 3, 3> toRotationMatrix() const {
    Matrix<T, 3, 3> m;
    const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
    const T twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const T txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const T tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    m(0, 0) = T(1) - (tyy + tzz); m(0, 1) = txy - twz; m(0, 2) = txz + twy;
    m(1, 0) = txy + twz; m(1, 1) = T(1) - (txx + tzz); m(1, 2) = tyz - twx;
    m(2, 0) = txz - twy; m(2, 1) = tyz + twx; m(2, 2) = T(1) - (txx + tyy);
    return m;
  }
  Quaternion inverse() const { T n = w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_; return Quaternion(w_ / n, -x_ / n, -y_ / n, -z_ / n); }
};
template <class T, int R, int C>
Matrix<T, R, C>& Matrix<T, R, C>::operator=(const Quaternion<T>& q) {
  static_assert(R == 3 && C == 3, "quaternion -> 3x3");
  *this = q.toRotationMatrix();
  return *this;
}

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Quaternion<double> Quaterniond;
template <class T> struct aligned_allocator : std::allocator<T> {};
}  // namespace Eigen
