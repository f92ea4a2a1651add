// TEST INFRASTRUCTURE.  Minimal stand-in for <boost/functional/hash.hpp> so that the reference's
// common/motion_planning.h compiles offline (Boost is not installed here).  Only boost::hash_combine is
// used by the reference (std::hash<Location>, std::hash<State>).  It is restated from Boost's documented
// 64-bit behaviour before 1.81 (hash<double> of a normal number = its bit pattern, +-0 -> 0; combine =
// the MurmurHash2-style mix), the versions Ubuntu 20.04/22.04 ship: the bucket order of
// std::unordered_set<Location> -- which decides isPointCollision's "first hit" (corridor.cc:32-52) --
// then matches such a build.  oracle/_ref exports the iteration order it actually used, and the parity
// tests feed that order to the restatement and the CUDA path.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace boost {
namespace stub_detail {
inline std::size_t hash_value(double v) {
  if (v == 0) return 0;
  if (std::isinf(v)) return (std::size_t)(v > 0 ? -1 : -2);
  if (std::isnan(v)) return (std::size_t)(-3);
  std::uint64_t b;
  std::memcpy(&b, &v, sizeof b);
  return (std::size_t)b;
}
inline std::size_t hash_value(float v) { return hash_value((double)v) ; }
inline std::size_t hash_value(int v) { return (std::size_t)v; }
inline std::size_t hash_value(std::size_t v) { return v; }
}  // namespace stub_detail

template <class T>
inline void hash_combine(std::size_t &seed, const T &v) {
  const std::uint64_t m = 0xc6a4a7935bd1e995ull;
  const int r = 47;
  std::uint64_t k = (std::uint64_t)stub_detail::hash_value(v);
  k *= m; k ^= k >> r; k *= m;
  seed ^= k; seed *= m;
  seed += 0xe6546b64;
}
}  // namespace boost
