// TEST INFRASTRUCTURE.  Minimal stand-in for boost::numeric::ublas::matrix<T> (dense, row major) with
// the members the reference's `State` uses (common/motion_planning.h:120-126, 186-189): constructor,
// resize, operator(), prod.
#pragma once
#include <cstddef>
#include <vector>

namespace boost { namespace numeric { namespace ublas {
template <class T>
class matrix {
 public:
  matrix() : r_(0), c_(0) {}
  matrix(std::size_t r, std::size_t c) : r_(r), c_(c), v_(r * c) {}
  void resize(std::size_t r, std::size_t c) { r_ = r; c_ = c; v_.assign(r * c, T()); }
  T &operator()(std::size_t i, std::size_t j) { return v_[i * c_ + j]; }
  const T &operator()(std::size_t i, std::size_t j) const { return v_[i * c_ + j]; }
  std::size_t size1() const { return r_; }
  std::size_t size2() const { return c_; }
 private:
  std::size_t r_, c_;
  std::vector<T> v_;
};
template <class T>
matrix<T> prod(const matrix<T> &a, const matrix<T> &b) {
  matrix<T> o(a.size1(), b.size2());
  for (std::size_t i = 0; i < a.size1(); ++i)
    for (std::size_t j = 0; j < b.size2(); ++j) {
      T s = T();
      for (std::size_t k = 0; k < a.size2(); ++k) s += a(i, k) * b(k, j);
      o(i, j) = s;
    }
  return o;
}
}}}  // namespace boost::numeric::ublas
