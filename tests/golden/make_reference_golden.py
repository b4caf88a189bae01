"""Writes tests/golden/reference_*.npz: seeded inputs and the outputs of THE REFERENCE'S OWN CODE for the hot path.

The reference (AaltoML/BayesNewton v1.3.4) is pure Python on jax 0.4.14 / objax, which this image does not have.  Its
source is imported UNMODIFIED from /root/reference and executed on oracle/jaxshim -- a NumPy/SciPy stand-in for the
slice of the jax / objax API it uses (vmap = a loop, lax.scan = a loop, lax.associative_scan = jax's published odd/even
recursion, LAPACK Cholesky, forward-mode duals for grad / jacrev; float64 throughout).  The arithmetic recorded here is
therefore what the reference's functions compute, statement by statement; only the array library underneath differs
from XLA's (rounding-level differences in the order of fused / vectorised operations).

Run in the build container (needs /root/reference):   python tests/golden/make_reference_golden.py
The .npz files are committed; the tests that read them (tests/test_reference_golden.py) never import the shim or the
reference, so they also run on the GPU box, where neither exists.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('BN_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'jaxshim'))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

import bayesnewton as bn  # noqa: E402  (the reference, on the shim)


def A(x):
    return np.array(getattr(x, 'value', x), dtype=np.float64)


def filter_problem(N, D=1, seed=0, missing=0.1):
    rng = np.random.default_rng(seed)
    dt = np.concatenate([[0.0], 0.1 + 0.4 * rng.uniform(size=N - 1)])
    y = rng.standard_normal((N, D, 1))
    R = np.zeros((N, D, D))
    for i in range(D):
        R[:, i, i] = 0.3 + rng.uniform(size=N)
    if D == 2:
        off = 0.1 * rng.standard_normal(N)
        R[:, 0, 1] = off
        R[:, 1, 0] = off
    mask = rng.uniform(size=(N, D, 1)) < missing
    return dt, y, R, mask


def kernels():
    K = bn.kernels
    return {
        'm12': (lambda: K.Matern12(variance=0.8, lengthscale=1.7), 1),
        'm32': (lambda: K.Matern32(variance=1.1, lengthscale=0.6), 1),
        'm52': (lambda: K.Matern52(variance=1.3, lengthscale=0.9), 1),
        'm72': (lambda: K.Matern72(variance=0.7, lengthscale=1.4), 1),
        'ind32': (lambda: K.Independent([K.Matern32(variance=1.0, lengthscale=1.0), K.Matern32(variance=0.5, lengthscale=2.0)]), 2),
        'ind52': (lambda: K.Independent([K.Matern52(variance=1.3, lengthscale=0.9), K.Matern52(variance=0.7, lengthscale=2.1)]), 2),
    }


def ops_cases():
    """kalman_filter / rauch_tung_striebel_smoother (ops.py:256-285, 357-380), both forms, masks, return_predict / return_full;
    the discretisation arrays (kernels.py state_transition, ops.py:149-151)"""
    out = {}
    for name, (mk, D) in kernels().items():
        k = mk()
        N = 61
        dt, y, R, mask = filter_problem(N, D=D, seed=len(name) + D)
        out['%s_dt' % name], out['%s_y' % name], out['%s_R' % name], out['%s_mask' % name] = dt, y, R, mask
        mask = mask.reshape(N, D)  # the reference's own callers pass the [N, D] form (basemodels.py:138): mvn_logpdf's
        # np.diag(mask) then builds the D x D selector; the [N, D, 1] form of the docstring takes np.diag of a column
        As = np.stack([A(k.state_transition(d)) for d in dt])
        Pinf = A(k.stationary_covariance())
        out['%s_As' % name] = As
        out['%s_Qs' % name] = np.stack([A(bn.ops.process_noise_covariance(a, Pinf)) for a in As])
        out['%s_Pinf' % name] = Pinf
        out['%s_H' % name] = A(k.measurement_model())
        for par in (False, True):
            tag = '%s_%s' % (name, 'par' if par else 'seq')
            for rp in (False, True):
                ell, (fm, fP) = bn.ops.kalman_filter(dt, k, y, R, mask, parallel=par, return_predict=rp)
                sfx = '_pred' if rp else ''
                out[tag + '_ell' + sfx], out[tag + '_fm' + sfx], out[tag + '_fP' + sfx] = A(ell), A(fm), A(fP)
            ell, (fm, fP) = bn.ops.kalman_filter(dt, k, y, R, mask, parallel=par)
            dts = np.concatenate([dt[1:], [0.0]])
            for rf in (False, True):
                sm, sP, G = bn.ops.rauch_tung_striebel_smoother(dts, k, fm, fP, return_full=rf, parallel=par)
                sfx = '_full' if rf else ''
                out[tag + '_sm' + sfx], out[tag + '_sP' + sfx], out[tag + '_gain' + sfx] = A(sm), A(sP), A(G)
    return out


def classification_data(N, seed):
    rng = np.random.default_rng(seed)
    x = np.sort(100 * rng.random(N))
    f = 6 * np.sin(np.pi * x / 10.0) / (np.pi * x / 10.0 + 1)
    y = (f + np.sqrt(0.05) * rng.standard_normal(N) > 0).astype(np.float64)
    return x, y


def model_cases():
    """model.inference(lr) x 3 and model.energy() (inference.py:65-90 + the scheme's update_variational_params / energy)
    for Markov{Variational, ExpectationPropagation, Laplace, PosteriorLinearisation}GP with single-latent likelihoods,
    sequential and parallel forms, missing observations"""
    out = {}
    M, Lk = bn.models, bn.likelihoods
    x, yc = classification_data(80, seed=3)
    rng = np.random.default_rng(5)
    data = {
        'probit': (lambda: Lk.Bernoulli(link='probit'), yc.copy()),
        'logit': (lambda: Lk.Bernoulli(link='logit'), yc.copy()),
        'gaussian': (lambda: Lk.Gaussian(variance=0.3), np.sin(0.3 * x) + 0.4 * rng.standard_normal(x.shape[0])),
        'poisson': (lambda: Lk.Poisson(binsize=1.0, link='exp'), rng.poisson(np.exp(0.5 * np.sin(0.3 * x))).astype(np.float64)),
    }
    out['x'] = x
    methods = {'vi': (M.MarkovVariationalGP, {}), 'ep': (M.MarkovExpectationPropagationGP, dict(power=0.5)),
               'newton': (M.MarkovLaplaceGP, {}), 'pl': (M.MarkovPosteriorLinearisationGP, {})}
    for lname, (mk_lik, y) in data.items():
        y = y.copy()
        y[::13] = np.nan
        out['y_%s' % lname] = y
        for mname, (cls, kw) in methods.items():
            if lname in ('logit', 'poisson') and mname == 'pl':
                continue
            for par in (False, True):
                m = cls(kernel=bn.kernels.Matern52(variance=1.5, lengthscale=0.75), likelihood=mk_lik(), X=x, Y=y,
                        parallel=par, **kw)
                tag = '%s_%s_%s' % (lname, mname, 'par' if par else 'seq')
                diffs, energies = [], []
                for it in range(3):
                    _, (d1, d2) = m.inference(lr=0.6)
                    diffs.append([float(d1), float(d2)])
                    energies.append(float(m.energy()))
                out[tag + '_post_mean'], out[tag + '_post_var'] = A(m.posterior_mean), A(m.posterior_variance)
                out[tag + '_site_mean'], out[tag + '_site_cov'] = A(m.pseudo_likelihood.mean), A(m.pseudo_likelihood.covariance)
                out[tag + '_site_nat1'], out[tag + '_site_nat2'] = A(m.pseudo_likelihood.nat1), A(m.pseudo_likelihood.nat2)
                out[tag + '_diffs'], out[tag + '_energy'] = np.array(diffs), np.array(energies)
                out[tag + '_log_lik'] = A(m.compute_log_lik())
                out[tag + '_kl'] = A(m.compute_kl())
                if not par and mname == 'vi':
                    xt = np.linspace(x[0] - 3.0, x[-1] + 3.0, 37)
                    pm, pv = m.predict(X=xt)
                    out[tag + '_xtest'], out[tag + '_pred_mean'], out[tag + '_pred_var'] = xt, A(pm), A(pv)
                    if lname in ('probit', 'gaussian'):
                        ym, yv = m.predict_y(X=xt)
                        out[tag + '_predy_mean'], out[tag + '_predy_var'] = A(ym), A(yv)
    return out


def likelihood_cases():
    """the per-step statistics the schemes vmap over time (likelihoods.py:336-355, 363-383, 401-412, 436-446;
    cubature.py:198-246, 310-435): variational_expectation, moment_match, log_likelihood_gradients,
    statistical_linear_regression at a grid of (y, m, v)"""
    out = {}
    Lk = bn.likelihoods
    rng = np.random.default_rng(9)
    n = 40
    m = 2.0 * rng.standard_normal(n)
    v = np.exp(rng.uniform(-4.0, 1.0, size=n))
    out['m'], out['v'] = m, v
    liks = {'probit': (Lk.Bernoulli(link='probit'), (rng.random(n) < 0.5).astype(np.float64)),
            'logit': (Lk.Bernoulli(link='logit'), (rng.random(n) < 0.5).astype(np.float64)),
            'gaussian': (Lk.Gaussian(variance=0.3), rng.standard_normal(n)),
            'poisson': (Lk.Poisson(binsize=1.0, link='exp'), rng.poisson(1.5, size=n).astype(np.float64))}
    for name, (lik, y) in liks.items():
        out['y_%s' % name] = y
        ve, mm, ll, slr = [], [], [], []
        for i in range(n):
            yi, mi, vi = np.array([y[i]]), np.array([[m[i]]]), np.array([[v[i]]])
            e, d1, d2 = lik.variational_expectation(yi, mi, vi, None)
            ve.append([float(np.squeeze(A(e))), float(np.squeeze(A(d1))), float(np.squeeze(A(d2)))])
            for power in (1.0, 0.5):
                z, z1, z2 = lik.moment_match(yi, mi, vi, power, None)
                mm.append([float(np.squeeze(A(z))), float(np.squeeze(A(z1))), float(np.squeeze(A(z2)))])
            l0, j, h = lik.log_likelihood_gradients(yi, mi)
            ll.append([float(np.squeeze(A(l0))), float(np.squeeze(A(j))), float(np.squeeze(A(h)))])
            if name in ('probit', 'gaussian'):
                mu, om, dmu, _ = lik.statistical_linear_regression(mi, vi, None)
                slr.append([float(np.squeeze(A(mu))), float(np.squeeze(A(om))), float(np.squeeze(A(dmu)))])
        out['%s_ve' % name], out['%s_mm' % name], out['%s_ll' % name] = np.array(ve), np.array(mm).reshape(n, 2, 3), np.array(ll)
        if slr:
            out['%s_slr' % name] = np.array(slr)
    return out


def more_likelihoods(rng, n):
    """StudentsT / Gamma / NegativeBinomial / Beta (likelihoods.py:1011-1189) with the observations each one expects"""
    Lk = bn.likelihoods
    gam = Lk.Gamma(link='exp')                    # shape = softplus(softplus_inv(1)) (likelihoods.py:1117)
    return {'studentst': (lambda: Lk.StudentsT(scale=0.7, df=4.0), 1.5 * rng.standard_t(4.0, size=n)),
            'gamma': (lambda: Lk.Gamma(link='exp'), rng.gamma(float(gam.shape), 1.3, size=n) + 1e-3),
            'negbin': (lambda: Lk.NegativeBinomial(alpha=0.6, link='exp', scale=1.5), rng.negative_binomial(2, 0.4, size=n).astype(np.float64)),
            'beta': (lambda: Lk.Beta(link='probit', scale=3.0), np.clip(rng.beta(1.5, 2.0, size=n), 1e-3, 1 - 1e-3))}


def likelihood2_cases():
    """the per-step statistics of the remaining single-latent likelihoods, same grid of (m, v) as likelihood_cases"""
    out = {}
    rng = np.random.default_rng(19)
    n = 40
    m = 1.2 * rng.standard_normal(n)
    v = np.exp(rng.uniform(-4.0, 0.5, size=n))
    out['m'], out['v'] = m, v
    for name, (mk, y) in more_likelihoods(rng, n).items():
        lik = mk()
        out['y_%s' % name] = y
        ve, mm, ll, slr = [], [], [], []
        for i in range(n):
            yi, mi, vi = np.array([y[i]]), np.array([[m[i]]]), np.array([[v[i]]])
            e, d1, d2 = lik.variational_expectation(yi, mi, vi, None)
            ve.append([float(np.squeeze(A(e))), float(np.squeeze(A(d1))), float(np.squeeze(A(d2)))])
            for power in (1.0, 0.5):
                z, z1, z2 = lik.moment_match(yi, mi, vi, power, None)
                mm.append([float(np.squeeze(A(z))), float(np.squeeze(A(z1))), float(np.squeeze(A(z2)))])
            l0, j, h = lik.log_likelihood_gradients(yi, mi)
            ll.append([float(np.squeeze(A(l0))), float(np.squeeze(A(j))), float(np.squeeze(A(h)))])
            mu, om, dmu, _ = lik.statistical_linear_regression(mi, vi, None)
            slr.append([float(np.squeeze(A(mu))), float(np.squeeze(A(om))), float(np.squeeze(A(dmu)))])
        out['%s_ve' % name], out['%s_mm' % name], out['%s_ll' % name] = np.array(ve), np.array(mm).reshape(n, 2, 3), np.array(ll)
        out['%s_slr' % name] = np.array(slr)
        xs = np.linspace(-1.0, 1.0, 9)
        ym, yv = [], []
        for a, b in zip(xs, np.linspace(0.05, 0.6, 9)):
            p1, p2 = lik.predict(np.array([[a]]), np.array([[b]]), None)
            ym.append(float(np.squeeze(A(p1))))
            yv.append(float(np.squeeze(A(p2))))
        out['%s_pred_in' % name], out['%s_pred_y' % name] = np.stack([xs, np.linspace(0.05, 0.6, 9)]), np.array([ym, yv])
    return out


def model2_cases():
    """model.inference(lr) x 3 + energy with the remaining likelihoods: Markov{Variational, ExpectationPropagation, Laplace,
    PosteriorLinearisation}GP, sequential form, missing observations"""
    out = {}
    M = bn.models
    rng = np.random.default_rng(23)
    N = 60
    x = np.sort(40 * rng.random(N))
    out['x'] = x
    methods = {'vi': (M.MarkovVariationalGP, {}), 'ep': (M.MarkovExpectationPropagationGP, dict(power=0.5)),
               'newton': (M.MarkovLaplaceGP, {}), 'pl': (M.MarkovPosteriorLinearisationGP, {})}
    for lname, (mk_lik, y) in more_likelihoods(rng, N).items():
        y = y.copy()
        y[::11] = np.nan
        out['y_%s' % lname] = y
        for mname, (cls, kw) in methods.items():
            if lname == 'negbin' and mname == 'newton':
                # the reference itself fails here: NegativeBinomial.evaluate_log_likelihood does not squeeze
                # (likelihoods.py:1179-1182), so log_likelihood_gradients hands a [N, 1] Hessian to utils.diag (utils.py:42)
                continue
            m = cls(kernel=bn.kernels.Matern32(variance=0.8, lengthscale=2.5), likelihood=mk_lik(), X=x, Y=y, parallel=False, **kw)
            tag = '%s_%s' % (lname, mname)
            diffs, energies = [], []
            for it in range(3):
                _, (d1, d2) = m.inference(lr=0.3)
                diffs.append([float(d1), float(d2)])
                energies.append(float(m.energy()))
            out[tag + '_post_mean'], out[tag + '_post_var'] = A(m.posterior_mean), A(m.posterior_variance)
            out[tag + '_site_nat1'], out[tag + '_site_nat2'] = A(m.pseudo_likelihood.nat1), A(m.pseudo_likelihood.nat2)
            out[tag + '_diffs'], out[tag + '_energy'] = np.array(diffs), np.array(energies)
    return out


def heteroscedastic_cases():
    """config C3 in miniature: Independent[Matern32 x 2] + HeteroscedasticNoise (demos/heteroscedastic.py:49-56), VI / EP / Newton"""
    out = {}
    rng = np.random.default_rng(3)
    N = 70
    x = np.sort(15 * rng.random(N))
    y = np.sin(x) + 0.3 * (1 + np.cos(x)) * rng.standard_normal(N)
    out['x'], out['y'] = x, y
    K, M = bn.kernels, bn.models
    for mname, (cls, kw) in {'vi': (M.MarkovVariationalGP, {}), 'ep': (M.MarkovExpectationPropagationGP, dict(power=0.5)),
                             'newton': (M.MarkovNewtonGP, {})}.items():
        m = cls(kernel=K.Independent([K.Matern32(variance=1.0, lengthscale=1.0), K.Matern32(variance=1.0, lengthscale=1.0)]),
                likelihood=bn.likelihoods.HeteroscedasticNoise(), X=x, Y=y, parallel=False, **kw)
        energies = []
        for it in range(2):
            m.inference(lr=0.3)
            energies.append(float(m.energy()))
        out['%s_post_mean' % mname], out['%s_post_var' % mname] = A(m.posterior_mean), A(m.posterior_variance)
        out['%s_site_nat1' % mname], out['%s_site_nat2' % mname] = A(m.pseudo_likelihood.nat1), A(m.pseudo_likelihood.nat2)
        out['%s_energy' % mname] = np.array(energies)
    return out


def regression_case():
    """BASELINE config 1 in miniature (demos/regression.py with MarkovVariationalGP): Gaussian likelihood, lr = 1 is exact"""
    out = {}
    N = 200
    x = np.linspace(-17, 147, N)
    rng = np.random.default_rng(12345)
    y = np.cos(0.04 * x + 0.33 * np.pi) * np.sin(0.2 * x) + np.sqrt(0.2) * rng.standard_normal(N)
    out['x'], out['y'] = x, y
    for par in (False, True):
        m = bn.models.MarkovVariationalGP(kernel=bn.kernels.Matern52(variance=1.0, lengthscale=5.0),
                                          likelihood=bn.likelihoods.Gaussian(variance=0.2), X=x, Y=y, parallel=par)
        m.inference(lr=1.0)
        tag = 'par' if par else 'seq'
        out['%s_post_mean' % tag], out['%s_post_var' % tag] = A(m.posterior_mean), A(m.posterior_variance)
        out['%s_energy' % tag] = A(m.energy())
    return out


def sparse_cases():
    """SparseMarkovVariationalGP (basemodels.py:928-1152; ops.py:383-426): pairs filter, joint of neighbouring inducing states,
    grouped site update; Bernoulli-probit and Gaussian likelihoods"""
    out = {}
    rng = np.random.default_rng(21)
    N = 70
    x = np.sort(np.linspace(-10.0, 30.0, N) + 0.5 * rng.standard_normal(N))
    f = np.cos(0.04 * x + 0.33 * np.pi) * np.sin(0.2 * x)
    z = np.linspace(x[0] + 0.7, x[-1] - 0.4, 15)
    out['x'], out['z'] = x, z
    for lname, (mk, y) in {'gaussian': (lambda: bn.likelihoods.Gaussian(variance=0.15), f + np.sqrt(0.15) * rng.standard_normal(N)),
                           'probit': (lambda: bn.likelihoods.Bernoulli(link='probit'), (f + 0.3 * rng.standard_normal(N) > 0).astype(np.float64))}.items():
        out['y_' + lname] = y
        m = bn.models.SparseMarkovVariationalGP(kernel=bn.kernels.Matern52(variance=1.2, lengthscale=4.0), likelihood=mk(),
                                                X=x, Y=y, Z=z)
        energies = []
        for it in range(3):
            m.inference(lr=0.7)
            energies.append(float(m.energy()))
        out[lname + '_energy'] = np.array(energies)
        out[lname + '_post_mean'], out[lname + '_post_var'] = A(m.posterior_mean), A(m.posterior_variance)
        out[lname + '_site_nat1'], out[lname + '_site_nat2'] = A(m.pseudo_likelihood.nat1), A(m.pseudo_likelihood.nat2)
        xt = np.linspace(x[0] - 2.0, x[-1] + 2.0, 31)
        pm, pv = m.predict(X=xt)
        out[lname + '_xtest'], out[lname + '_pred_mean'], out[lname + '_pred_var'] = xt, A(pm), A(pv)
    return out


def spacetime_cases():
    """MarkovVariationalGP / MarkovVariationalMeanFieldGP with a SpatioTemporalKernel (kernels.py:385-586; ops.py:429-706;
    basemodels.py:676-687, 743-764, 1155-1175): gridded data, missing values, Gaussian likelihood"""
    out = {}
    rng = np.random.default_rng(31)
    Nt, Ns = 14, 6
    t = np.sort(np.linspace(0.0, 6.0, Nt) + 0.05 * rng.standard_normal(Nt))
    r = np.linspace(-1.5, 1.5, Ns)
    T, Rr = np.meshgrid(t, r, indexing='ij')
    Y = np.sin(T) + np.cos(2 * Rr) + 0.1 * rng.standard_normal((Nt, Ns))
    Y[2, 1] = np.nan
    Y[9, 4] = np.nan
    X = t[:, None]
    R = np.tile(r[None, :, None], [Nt, 1, 1])
    out['t'], out['r'], out['Y'] = t, r, Y
    for name, cls in {'full': bn.models.MarkovVariationalGP, 'meanfield': bn.models.MarkovVariationalMeanFieldGP}.items():
        kern = bn.kernels.SpatioTemporalKernel(temporal_kernel=bn.kernels.Matern32(variance=1.0, lengthscale=2.0),
                                               spatial_kernel=bn.kernels.Matern32(variance=1.0, lengthscale=1.0),
                                               z=r[:, None], sparse=True, opt_z=False, conditional='Full')
        m = cls(kernel=kern, likelihood=bn.likelihoods.Gaussian(variance=0.5), X=X, R=R, Y=Y)
        energies = []
        for it in range(2):
            m.inference(lr=0.7)
            energies.append(float(m.energy()))
        out[name + '_energy'] = np.array(energies)
        out[name + '_post_mean'], out[name + '_post_var'] = A(m.posterior_mean), A(m.posterior_variance)
        out[name + '_site_nat1'], out[name + '_site_nat2'] = A(m.pseudo_likelihood.nat1), A(m.pseudo_likelihood.nat2)
        out[name + '_log_lik'] = A(m.compute_log_lik())
    return out


def infinite_horizon_cases():
    """InfiniteHorizonVariationalGP (basemodels.py:1257-1363; ops.py:796-1186): steady-state gains from the DARE fixed point,
    evenly spaced inputs, sequential and parallel forms"""
    out = {}
    rng = np.random.default_rng(41)
    N = 120
    x = np.linspace(0.0, 24.0, N)
    f = np.sin(x) + 0.5 * np.cos(0.3 * x)
    out['x'] = x
    for lname, (mk, y) in {'gaussian': (lambda: bn.likelihoods.Gaussian(variance=0.2), f + np.sqrt(0.2) * rng.standard_normal(N)),
                           'probit': (lambda: bn.likelihoods.Bernoulli(link='probit'), (f + 0.3 * rng.standard_normal(N) > 0).astype(np.float64))}.items():
        out['y_' + lname] = y
        for par in (False, True):
            for kname, mkk in {'m32': lambda: bn.kernels.Matern32(variance=1.0, lengthscale=1.0),
                               'm52': lambda: bn.kernels.Matern52(variance=1.3, lengthscale=1.5)}.items():
                m = bn.models.InfiniteHorizonVariationalGP(kernel=mkk(), likelihood=mk(), X=x, Y=y, parallel=par)
                tag = '%s_%s_%s' % (lname, kname, 'par' if par else 'seq')
                energies = []
                for it in range(3):
                    m.inference(lr=0.6)
                    energies.append(float(m.energy()))
                out[tag + '_energy'] = np.array(energies)
                out[tag + '_post_mean'], out[tag + '_post_var'] = A(m.posterior_mean), A(m.posterior_variance)
                out[tag + '_site_nat1'], out[tag + '_site_nat2'] = A(m.pseudo_likelihood.nat1), A(m.pseudo_likelihood.nat2)
                out[tag + '_log_lik'] = A(m.compute_log_lik())
    return out


def gradient_cases():
    """d energy / d (kernel variance, lengthscale, Gaussian likelihood variance) of the REFERENCE'S energy() with sites and
    posterior held fixed -- what objax.GradValues(model.energy, model.vars()) differentiates (README.md:56-70; the shim has
    no reverse mode): Richardson-extrapolated central differences (h, h/2 -> O(h^4)) of the reference's own function"""
    from bayesnewton.utils import softplus_inv
    out = {}
    M = bn.models
    rng = np.random.default_rng(31)
    N = 120
    x = np.sort(60 * rng.random(N))
    y = np.sin(0.3 * x) + 0.4 * rng.standard_normal(N)
    out['x'], out['y'] = x, y
    base = dict(var_f=1.2, len_f=4.0, var_y=0.3)
    for mname, (cls, kw) in {'vi': (M.MarkovVariationalGP, {}), 'newton': (M.MarkovLaplaceGP, {}),
                             'ep': (M.MarkovExpectationPropagationGP, dict(power=0.5))}.items():
        m = cls(kernel=bn.kernels.Matern52(variance=base['var_f'], lengthscale=base['len_f']),
                likelihood=bn.likelihoods.Gaussian(variance=base['var_y']), X=x, Y=y, parallel=False, **kw)
        m.inference(lr=0.7)

        def energy_at(**over):
            p = dict(base, **over)
            m.kernel.transformed_variance.value = np.array(softplus_inv(p['var_f']))
            m.kernel.transformed_lengthscale.value = np.array(softplus_inv(p['len_f']))
            m.likelihood.transformed_variance.value = np.array(softplus_inv(p['var_y']))
            return float(m.energy())
        out[mname + '_energy'] = energy_at()
        grads = []
        for name in ('var_f', 'len_f', 'var_y'):
            h = 2e-3 * base[name]
            d1 = (energy_at(**{name: base[name] + h}) - energy_at(**{name: base[name] - h})) / (2 * h)
            d2 = (energy_at(**{name: base[name] + h / 2}) - energy_at(**{name: base[name] - h / 2})) / h
            grads.append([(4 * d2 - d1) / 3, d2 - d1])
        out[mname + '_grad'] = np.array(grads)   # [3, 2]: derivative, and the h -> h/2 change as an error scale
        out[mname + '_site_mean'], out[mname + '_site_cov'] = A(m.pseudo_likelihood.mean), A(m.pseudo_likelihood.covariance)
    return out


def model_family_cases():
    """the model-level iteration for the other Matern families (the `models` cases are Matern-5/2): VI and Newton with a
    Bernoulli-probit likelihood, both forms, missing observations"""
    out = {}
    M = bn.models
    x, y = classification_data(70, seed=17)
    y = y.copy()
    y[::9] = np.nan
    out['x'], out['y'] = x, y
    fams = {'m12': lambda: bn.kernels.Matern12(variance=0.8, lengthscale=1.7), 'm32': lambda: bn.kernels.Matern32(variance=1.1, lengthscale=0.6),
            'm72': lambda: bn.kernels.Matern72(variance=0.7, lengthscale=1.4)}
    for fname, mk in fams.items():
        for mname, cls in {'vi': M.MarkovVariationalGP, 'newton': M.MarkovLaplaceGP}.items():
            for par in (False, True):
                m = cls(kernel=mk(), likelihood=bn.likelihoods.Bernoulli(link='probit'), X=x, Y=y, parallel=par)
                tag = '%s_%s_%s' % (fname, mname, 'par' if par else 'seq')
                energies = []
                for it in range(3):
                    m.inference(lr=0.6)
                    energies.append(float(m.energy()))
                out[tag + '_energy'] = np.array(energies)
                out[tag + '_post_mean'], out[tag + '_post_var'] = A(m.posterior_mean), A(m.posterior_variance)
                out[tag + '_site_nat1'], out[tag + '_site_nat2'] = A(m.pseudo_likelihood.nat1), A(m.pseudo_likelihood.nat2)
    return out


CASES = {'ops': ops_cases, 'models': model_cases, 'likelihoods': likelihood_cases, 'likelihoods2': likelihood2_cases, 'models2': model2_cases, 'heteroscedastic': heteroscedastic_cases,
         'regression': regression_case, 'sparse': sparse_cases, 'spacetime': spacetime_cases,
         'infinite_horizon': infinite_horizon_cases, 'gradient': gradient_cases, 'model_families': model_family_cases}

if __name__ == '__main__':
    which = sys.argv[1:] or sorted(CASES)
    for name in which:
        data = CASES[name]()
        path = os.path.join(HERE, 'reference_%s.npz' % name)
        np.savez_compressed(path, **data)
        print('wrote %s (%d arrays, %.0f KB)' % (path, len(data), os.path.getsize(path) / 1024))
