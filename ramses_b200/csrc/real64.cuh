// real64.cuh -- a double whose `/` and sqrt() are the branch-free, correctly rounded sequences of hydro_device.cuh
// (rcp_rn + Markstein division, RSQ seed + FMA refinement): the same bits as the IEEE operations on the operand range
// documented there (rgpu_selftest_div checks 2^28 pairs), but without the range check and slow-path call nvcc puts
// behind every `/` and sqrt(), so that independent chains interleave and equal divisors share one reciprocal (the
// compiler merges identical rcp_rn sequences).  +, -, * are the plain IEEE operations; evaluation order is untouched
// because the operators keep the built-in precedence and associativity.  Used by the MHD solvers, whose formulas are
// written exactly like the reference's (mhd/godunov_utils.f90).
#pragma once
#include "hydro_device.cuh"

namespace rgpu {

struct R64 {
  double v;
  __device__ __forceinline__ R64() {}
  __device__ __forceinline__ R64(double x) : v(x) {}
  __device__ __forceinline__ explicit operator double() const { return v; }
};
__device__ __forceinline__ R64 operator+(R64 a, R64 b) { return R64(a.v + b.v); }
__device__ __forceinline__ R64 operator-(R64 a, R64 b) { return R64(a.v - b.v); }
__device__ __forceinline__ R64 operator*(R64 a, R64 b) { return R64(a.v * b.v); }
__device__ __forceinline__ R64 operator-(R64 a) { return R64(-a.v); }
__device__ __forceinline__ R64 operator+(R64 a) { return a; }
// a / b.  Zero numerators are fine (q = 0); divisors are densities, wave-speed sums, sound speeds ... (normal numbers).
__device__ __forceinline__ R64 operator/(R64 a, R64 b) { return R64(div_rn(a.v, b.v, rcp_rn(b.v))); }
// a divisor with its correctly rounded reciprocal attached: x / RcpD(b) has the bits of x / b (Markstein), 3 FP64 ops
struct RcpD {
  double b, y;
  __device__ __forceinline__ explicit RcpD(R64 d) : b(d.v), y(rcp_rn(d.v)) {}
};
__device__ __forceinline__ R64 operator/(R64 a, const RcpD& d) { return R64(div_rn(a.v, d.b, d.y)); }
__device__ __forceinline__ bool operator<(R64 a, R64 b) { return a.v < b.v; }
__device__ __forceinline__ bool operator>(R64 a, R64 b) { return a.v > b.v; }
__device__ __forceinline__ bool operator<=(R64 a, R64 b) { return a.v <= b.v; }
__device__ __forceinline__ bool operator>=(R64 a, R64 b) { return a.v >= b.v; }
__device__ __forceinline__ bool operator==(R64 a, R64 b) { return a.v == b.v; }
__device__ __forceinline__ bool operator!=(R64 a, R64 b) { return a.v != b.v; }
// sqrt: arguments such as the slow speed squared or B_t^2 can be exactly zero (or tiny): those take the IEEE routine
__device__ __forceinline__ R64 rsqrt64(R64 x) { return R64(x.v < 1e-280 ? ::sqrt(x.v) : sqrt_rn(x.v)); }
__device__ __forceinline__ R64 rabs(R64 x) { return R64(::fabs(x.v)); }
__device__ __forceinline__ R64 rsign(R64 a, R64 b) { return R64(::copysign(a.v, b.v)); }
__device__ __forceinline__ R64 fmx(R64 a, R64 b) { return (b.v > a.v) ? b : a; }
__device__ __forceinline__ R64 fmn(R64 a, R64 b) { return (b.v < a.v) ? b : a; }

}  // namespace rgpu
