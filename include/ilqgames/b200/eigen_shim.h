// eigen_shim.h -- the sliver of Eigen's dense API that the ilqgames host classes and the
// in-scope example sources use (VectorXf, MatrixXf, Vector2f), for builds without Eigen3.
// Column-major like Eigen, float only, no expression templates.  With ILQGAMES_B200_USE_EIGEN
// defined the real Eigen is used instead and this file is a no-op.
#ifndef ILQGAMES_B200_EIGEN_SHIM_H
#define ILQGAMES_B200_EIGEN_SHIM_H

#ifdef ILQGAMES_B200_USE_EIGEN
#include <Eigen/Dense>
#else

#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <vector>

namespace Eigen {

class VectorXf {
 public:
  VectorXf() {}
  explicit VectorXf(long n) : v_((size_t)n, 0.f) {}
  static VectorXf Zero(long n) { return VectorXf(n); }
  static VectorXf Constant(long n, float c) { VectorXf out(n); for (auto& e : out.v_) e = c; return out; }
  long size() const { return (long)v_.size(); }
  long rows() const { return size(); }
  long cols() const { return 1; }
  void resize(long n) { v_.assign((size_t)n, 0.f); }
  void setZero() { for (auto& e : v_) e = 0.f; }
  float& operator()(long i) { return v_[(size_t)i]; }
  float operator()(long i) const { return v_[(size_t)i]; }
  float& operator[](long i) { return v_[(size_t)i]; }
  float operator[](long i) const { return v_[(size_t)i]; }
  float* data() { return v_.data(); }
  const float* data() const { return v_.data(); }
  VectorXf segment(long start, long n) const {
    VectorXf out(n);
    for (long i = 0; i < n; i++) out(i) = v_[(size_t)(start + i)];
    return out;
  }
  VectorXf head(long n) const { return segment(0, n); }
  VectorXf tail(long n) const { return segment(size() - n, n); }
  float squaredNorm() const { float s = 0; for (float e : v_) s += e * e; return s; }
  float norm() const { return std::sqrt(squaredNorm()); }
  float dot(const VectorXf& o) const { float s = 0; for (long i = 0; i < size(); i++) s += v_[(size_t)i] * o(i); return s; }
  VectorXf& operator+=(const VectorXf& o) { for (long i = 0; i < size(); i++) v_[(size_t)i] += o(i); return *this; }
  VectorXf& operator-=(const VectorXf& o) { for (long i = 0; i < size(); i++) v_[(size_t)i] -= o(i); return *this; }
  VectorXf& operator*=(float c) { for (auto& e : v_) e *= c; return *this; }
  bool isApprox(const VectorXf& o, float prec = 1e-5f) const {
    float d = 0, a = 0, b = 0;
    for (long i = 0; i < size(); i++) { const float t = v_[(size_t)i] - o(i); d += t * t; a += v_[(size_t)i] * v_[(size_t)i]; b += o(i) * o(i); }
    return d <= prec * prec * (a < b ? a : b);
  }

 private:
  std::vector<float> v_;
};

inline VectorXf operator+(VectorXf a, const VectorXf& b) { a += b; return a; }
inline VectorXf operator-(VectorXf a, const VectorXf& b) { a -= b; return a; }
inline VectorXf operator*(VectorXf a, float c) { a *= c; return a; }
inline VectorXf operator*(float c, VectorXf a) { a *= c; return a; }
inline VectorXf operator-(VectorXf a) { a *= -1.f; return a; }

class MatrixXf {
 public:
  MatrixXf() : r_(0), c_(0) {}
  MatrixXf(long r, long c) : r_(r), c_(c), v_((size_t)(r * c), 0.f) {}
  static MatrixXf Zero(long r, long c) { return MatrixXf(r, c); }
  static MatrixXf Identity(long r, long c) {
    MatrixXf m(r, c);
    for (long i = 0; i < (r < c ? r : c); i++) m(i, i) = 1.f;
    return m;
  }
  long rows() const { return r_; }
  long cols() const { return c_; }
  long size() const { return r_ * c_; }
  void setZero() { for (auto& e : v_) e = 0.f; }
  float& operator()(long i, long j) { return v_[(size_t)(j * r_ + i)]; }
  float operator()(long i, long j) const { return v_[(size_t)(j * r_ + i)]; }
  float* data() { return v_.data(); }
  const float* data() const { return v_.data(); }
  MatrixXf transpose() const {
    MatrixXf t(c_, r_);
    for (long i = 0; i < r_; i++) for (long j = 0; j < c_; j++) t(j, i) = (*this)(i, j);
    return t;
  }
  MatrixXf& operator+=(const MatrixXf& o) { for (size_t i = 0; i < v_.size(); i++) v_[i] += o.v_[i]; return *this; }
  MatrixXf& operator-=(const MatrixXf& o) { for (size_t i = 0; i < v_.size(); i++) v_[i] -= o.v_[i]; return *this; }
  MatrixXf& operator*=(float c) { for (auto& e : v_) e *= c; return *this; }
  float cwiseAbsMax() const { float m = 0; for (float e : v_) m = std::fabs(e) > m ? std::fabs(e) : m; return m; }

 private:
  long r_, c_;
  std::vector<float> v_;
};

inline MatrixXf operator+(MatrixXf a, const MatrixXf& b) { a += b; return a; }
inline MatrixXf operator-(MatrixXf a, const MatrixXf& b) { a -= b; return a; }
inline MatrixXf operator*(MatrixXf a, float c) { a *= c; return a; }
inline MatrixXf operator*(float c, MatrixXf a) { a *= c; return a; }
inline MatrixXf operator*(const MatrixXf& a, const MatrixXf& b) {
  MatrixXf out(a.rows(), b.cols());
  for (long j = 0; j < b.cols(); j++)
    for (long k = 0; k < a.cols(); k++) {
      const float bkj = b(k, j);
      for (long i = 0; i < a.rows(); i++) out(i, j) += a(i, k) * bkj;
    }
  return out;
}
inline VectorXf operator*(const MatrixXf& a, const VectorXf& x) {
  VectorXf out(a.rows());
  for (long k = 0; k < a.cols(); k++)
    for (long i = 0; i < a.rows(); i++) out(i) += a(i, k) * x(k);
  return out;
}

class Vector2f {
 public:
  Vector2f() : x_(0.f), y_(0.f) {}
  Vector2f(float x, float y) : x_(x), y_(y) {}
  static Vector2f Zero() { return Vector2f(); }
  float x() const { return x_; }
  float y() const { return y_; }
  float& x() { return x_; }
  float& y() { return y_; }
  float operator()(long i) const { return i == 0 ? x_ : y_; }
  float dot(const Vector2f& o) const { return x_ * o.x_ + y_ * o.y_; }
  float squaredNorm() const { return dot(*this); }
  float norm() const { return std::sqrt(squaredNorm()); }
  bool operator==(const Vector2f& o) const { return x_ == o.x_ && y_ == o.y_; }
  bool operator!=(const Vector2f& o) const { return !(*this == o); }
  Vector2f operator+(const Vector2f& o) const { return Vector2f(x_ + o.x_, y_ + o.y_); }
  Vector2f operator-(const Vector2f& o) const { return Vector2f(x_ - o.x_, y_ - o.y_); }
  Vector2f operator*(float c) const { return Vector2f(x_ * c, y_ * c); }
  Vector2f operator/(float c) const { return Vector2f(x_ / c, y_ / c); }

 private:
  float x_, y_;
};
inline Vector2f operator*(float c, const Vector2f& v) { return v * c; }

}  // namespace Eigen

#endif  // ILQGAMES_B200_USE_EIGEN
#endif
