// Forward-mode dual number (value, derivative): the closed-form discretisations of gen.cuh are
// templates over the scalar type, so d A_k / d lambda and d Pinf / d lengthscale come from the SAME
// formulas (kernels.py:158-165,207-224,273-293,344-382) evaluated on Dual -- no second hand-written copy.
#pragma once
#include "smallmat.cuh"

namespace BN_NS {

using ::exp;   // keep the double overloads visible next to the Dual ones below
using ::sqrt;

struct Dual {
    double v, d;
    BN_DEV Dual() : v(0.0), d(0.0) {}
    BN_DEV Dual(double v_) : v(v_), d(0.0) {}
    BN_DEV Dual(double v_, double d_) : v(v_), d(d_) {}
};
BN_DEV Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
BN_DEV Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
BN_DEV Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
BN_DEV Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, fma(a.v, b.d, a.d * b.v)); }
BN_DEV Dual operator/(Dual a, Dual b) {
    const double r = 1.0 / b.v, q = a.v * r;
    return Dual(q, (a.d - q * b.d) * r);
}
BN_DEV Dual exp(Dual a) {
    const double e = ::exp(a.v);
    return Dual(e, e * a.d);
}
BN_DEV Dual sqrt(Dual a) {
    const double s = ::sqrt(a.v);
    return Dual(s, 0.5 * a.d / s);
}

}  // namespace BN_NS
