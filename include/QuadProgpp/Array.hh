// Minimal ArrayHH::Vector / ArrayHH::Matrix for the mgqp drop-in.
//
// The reference ships a 2.5k-line container header (include/QuadProgpp/Array.hh); its only caller
// on the solver path, solveNextStep (src/mgqp.cpp:695-725), uses the default constructor, resize(),
// operator[], nrows()/ncols()/size() and the free function t().  This header provides exactly
// those, with the SAME object layout (Vector: {unsigned n; T* v}, Array.hh:57-59; Matrix:
// {unsigned n, m; T** v} with one contiguous row-major block behind v[0] and a row-pointer table,
// Array.hh:899-902, :910-919), so objects built against either header are interchangeable across
// the solve_quadprog() call.  It is written from that description, not copied.
#ifndef MGQP_B200_ARRAYHH_COMPAT
#define MGQP_B200_ARRAYHH_COMPAT

#include <cstddef>

namespace ArrayHH {

template <typename T>
class Vector {
 public:
  Vector() : n(0), v(nullptr) {}
  explicit Vector(unsigned int len) : n(len), v(new T[len]) {}
  Vector(const T& a, unsigned int len) : n(len), v(new T[len]) {
    for (unsigned int i = 0; i < n; i++) v[i] = a;
  }
  Vector(const T* a, unsigned int len) : n(len), v(new T[len]) {
    for (unsigned int i = 0; i < n; i++) v[i] = a[i];
  }
  Vector(const Vector& o) : n(o.n), v(new T[o.n]) {
    for (unsigned int i = 0; i < n; i++) v[i] = o.v[i];
  }
  ~Vector() { delete[] v; }
  Vector& operator=(const Vector& o) {
    if (this != &o) {
      resize(o.n);
      for (unsigned int i = 0; i < n; i++) v[i] = o.v[i];
    }
    return *this;
  }
  T& operator[](const unsigned int& i) { return v[i]; }
  const T& operator[](const unsigned int& i) const { return v[i]; }
  unsigned int size() const { return n; }
  void resize(const unsigned int len) {
    if (len == n) return;
    delete[] v;
    n = len;
    v = new T[len];
  }

 private:
  unsigned int n;
  T* v;
};

template <typename T>
class Matrix {
 public:
  Matrix() : n(0), m(0), v(nullptr) {}
  Matrix(unsigned int rows, unsigned int cols) : n(0), m(0), v(nullptr) { alloc(rows, cols); }
  Matrix(const T& a, unsigned int rows, unsigned int cols) : n(0), m(0), v(nullptr) {
    alloc(rows, cols);
    for (std::size_t i = 0; i < (std::size_t)n * m; i++) v[0][i] = a;
  }
  Matrix(const Matrix& o) : n(0), m(0), v(nullptr) {
    alloc(o.n, o.m);
    for (std::size_t i = 0; i < (std::size_t)n * m; i++) v[0][i] = o.v[0][i];
  }
  ~Matrix() { release(); }
  Matrix& operator=(const Matrix& o) {
    if (this != &o) {
      resize(o.n, o.m);
      for (std::size_t i = 0; i < (std::size_t)n * m; i++) v[0][i] = o.v[0][i];
    }
    return *this;
  }
  T* operator[](const unsigned int& i) { return v[i]; }
  const T* operator[](const unsigned int& i) const { return v[i]; }
  unsigned int nrows() const { return n; }
  unsigned int ncols() const { return m; }
  void resize(const unsigned int rows, const unsigned int cols) {
    if (rows == n && cols == m) return;
    release();
    alloc(rows, cols);
  }

 private:
  // Always keeps one slot in the row table so that v[0] is valid even for a 0 x m matrix (the
  // reference writes v[0] of a zero-length table there: SURVEY.md section 5, Array.hh:1060-1061).
  void alloc(unsigned int rows, unsigned int cols) {
    n = rows;
    m = cols;
    v = new T*[rows ? rows : 1];
    v[0] = new T[(std::size_t)rows * cols ? (std::size_t)rows * cols : 1];
    for (unsigned int i = 1; i < rows; i++) v[i] = v[i - 1] + cols;
  }
  void release() {
    if (v) {
      delete[] v[0];
      delete[] v;
    }
    v = nullptr;
  }
  unsigned int n;
  unsigned int m;
  T** v;
};

// transposed copy (the caller passes t(CE), t(CI): src/mgqp.cpp:708)
template <typename T>
Matrix<T> t(const Matrix<T>& a) {
  Matrix<T> r(a.ncols(), a.nrows());
  for (unsigned int i = 0; i < a.nrows(); i++)
    for (unsigned int j = 0; j < a.ncols(); j++) r[j][i] = a[i][j];
  return r;
}

}  // namespace ArrayHH
#endif
