// compute_conditional_statistics (bayesnewton/utils.py:173-215): the conditional of the state at a time t on the
// states u_-, u_+ of the two neighbouring inducing / training inputs,
//     p(x_t | u_-, u_+) = N([P1, W] [u_-; u_+], T),
// from the two gaps dt_fwd = t - t_-, dt_back = t_+ - t.  Everything in registers; A(dt), Q(dt) from the closed forms.
#pragma once
#include "gen.cuh"

namespace bn {

template <class Gen>
BN_DEV void cond_stats(const Gen& gen, double dt_fwd, double dt_back, double* P1, double* W, double* Tm) {
    constexpr int d = Gen::d;
    double Af[d * d], Ab[d * d], Qf[symn(d)], Qb[symn(d)];
    gen.step_dt(dt_fwd, Af, Qf);
    gen.step_dt(dt_back, Ab, Qb);
    // Q_mp = Q_back + A_back Q_fwd A_back^T + 1e-8 I;  V = Q_mp^-1 A_back   (utils.py:196-202)
    double AbQf[d * d];
    mat_sym<d, d>(Ab, Qf, AbQf);
    double Qmp[symn(d)];
    abt_sym<d, d>(AbQf, Ab, Qb, Qmp);
#pragma unroll
    for (int i = 0; i < d; ++i) Qmp[sidx(i, i)] += 1e-8;
    chol<d>(Qmp);
    double V[d * d];
#pragma unroll
    for (int i = 0; i < d * d; ++i) V[i] = Ab[i];
    chol_solve<d, d>(Qmp, V);
    // W = Q_fwd V^T;  T = Q_fwd - (A_back Q_fwd)^T V Q_fwd = Q_fwd - W A_back Q_fwd   (:204-207)
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) {
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(Qf[sidx(i, l)], V[j * d + l], s);
            W[i * d + j] = s;
        }
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) {
            double s = Qf[sidx(i, j)];
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(-W[i * d + l], AbQf[l * d + j], s);
            Tm[i * d + j] = s;
        }
    // P = [A_fwd - W A_back A_fwd, W]   (:208)
    double WAb[d * d];
    matmul<d, d, d>(W, Ab, WAb);
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) {
            double s = Af[i * d + j];
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(-WAb[i * d + l], Af[l * d + j], s);
            P1[i * d + j] = s;
        }
}

}  // namespace bn
