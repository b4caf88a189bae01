// The scalar type of the fused-iteration kernels is a build parameter: `real` = double (default) or, with
// -DBN_REAL32, r32 -- a float that swallows the double literals of the formulas at compile time, so the same
// kernel text compiles to pure fp32 arithmetic (no silent promotion to fp64 through a literal).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <type_traits>

#ifndef BN_DEV
#define BN_DEV __host__ __device__ __forceinline__
#endif
// the namespace of everything built on `real`: bn in the fp64 library, bn32 in the fp32 translation units, so that
// the two builds of the same templates do not collide at link time
#ifndef BN_NS
#define BN_NS bn
#endif

namespace BN_NS {

#ifdef BN_REAL32
// the fp64 overloads stay visible next to the r32 ones declared below (a declaration in this namespace would
// otherwise hide the global ones)
using ::sqrt; using ::exp; using ::log; using ::erf; using ::erfc; using ::lgamma; using ::fabs; using ::log1p;
using ::expm1; using ::floor; using ::fma; using ::fmax; using ::fmin; using ::isnan; using ::isfinite;
#ifdef __CUDACC__
using ::rsqrt; using ::__shfl_sync; using ::__shfl_up_sync; using ::__shfl_down_sync; using ::__shfl_xor_sync; using ::__ldg;
#endif
struct r32 {
    float v;
    r32() = default;
    template <class U, class = typename std::enable_if<std::is_arithmetic<U>::value>::type>
    constexpr BN_DEV r32(U x) : v((float)x) {}
    explicit constexpr BN_DEV operator float() const { return v; }
    explicit constexpr BN_DEV operator double() const { return (double)v; }
    explicit constexpr BN_DEV operator int() const { return (int)v; }
    explicit constexpr BN_DEV operator long long() const { return (long long)v; }
    BN_DEV r32& operator+=(r32 o) { v += o.v; return *this; }
    BN_DEV r32& operator-=(r32 o) { v -= o.v; return *this; }
    BN_DEV r32& operator*=(r32 o) { v *= o.v; return *this; }
    BN_DEV r32& operator/=(r32 o) { v /= o.v; return *this; }
};
constexpr BN_DEV r32 operator+(r32 a, r32 b) { return r32(a.v + b.v); }
constexpr BN_DEV r32 operator-(r32 a, r32 b) { return r32(a.v - b.v); }
constexpr BN_DEV r32 operator*(r32 a, r32 b) { return r32(a.v * b.v); }
constexpr BN_DEV r32 operator/(r32 a, r32 b) { return r32(a.v / b.v); }
constexpr BN_DEV r32 operator-(r32 a) { return r32(-a.v); }
constexpr BN_DEV r32 operator+(r32 a) { return a; }
constexpr BN_DEV bool operator<(r32 a, r32 b) { return a.v < b.v; }
constexpr BN_DEV bool operator>(r32 a, r32 b) { return a.v > b.v; }
constexpr BN_DEV bool operator<=(r32 a, r32 b) { return a.v <= b.v; }
constexpr BN_DEV bool operator>=(r32 a, r32 b) { return a.v >= b.v; }
constexpr BN_DEV bool operator==(r32 a, r32 b) { return a.v == b.v; }
constexpr BN_DEV bool operator!=(r32 a, r32 b) { return a.v != b.v; }
#define BN_R32_FN1(name, fn) BN_DEV r32 name(r32 a) { return r32(fn(a.v)); }
BN_R32_FN1(sqrt, ::sqrtf) BN_R32_FN1(exp, ::expf) BN_R32_FN1(log, ::logf) BN_R32_FN1(erf, ::erff) BN_R32_FN1(erfc, ::erfcf)
BN_R32_FN1(lgamma, ::lgammaf) BN_R32_FN1(fabs, ::fabsf) BN_R32_FN1(log1p, ::log1pf) BN_R32_FN1(expm1, ::expm1f)
BN_R32_FN1(floor, ::floorf)
#undef BN_R32_FN1
BN_DEV r32 fma(r32 a, r32 b, r32 c) { return r32(::fmaf(a.v, b.v, c.v)); }
BN_DEV r32 fmax(r32 a, r32 b) { return r32(::fmaxf(a.v, b.v)); }
BN_DEV r32 fmin(r32 a, r32 b) { return r32(::fminf(a.v, b.v)); }
BN_DEV bool isnan(r32 a) { return a.v != a.v; }
BN_DEV bool isfinite(r32 a) { return ::fabsf(a.v) <= 3.402823466e38f; }
BN_DEV r32 rsqrt(r32 a) {
#ifdef __CUDA_ARCH__
    return r32(::rsqrtf(a.v));
#else
    return r32(1.0f / ::sqrtf(a.v));
#endif
}
#ifdef __CUDACC__
__device__ __forceinline__ r32 __shfl_sync(unsigned m, r32 a, int l, int w = 32) { return r32(::__shfl_sync(m, a.v, l, w)); }
__device__ __forceinline__ r32 __shfl_up_sync(unsigned m, r32 a, unsigned l, int w = 32) { return r32(::__shfl_up_sync(m, a.v, l, w)); }
__device__ __forceinline__ r32 __shfl_down_sync(unsigned m, r32 a, unsigned l, int w = 32) { return r32(::__shfl_down_sync(m, a.v, l, w)); }
__device__ __forceinline__ r32 __shfl_xor_sync(unsigned m, r32 a, int l, int w = 32) { return r32(::__shfl_xor_sync(m, a.v, l, w)); }
__device__ __forceinline__ r32 __ldg(const r32* p) { return r32(::__ldg(&p->v)); }
#endif
using real = r32;
constexpr bool kReal32 = true;
#else
using real = double;
constexpr bool kReal32 = false;
#endif

}  // namespace BN_NS
