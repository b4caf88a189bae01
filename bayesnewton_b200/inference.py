"""Inference mixins with the reference's interface (bayesnewton/inference.py:42-428):
``model.inference(lr, **kw)`` runs update_posterior -> site statistics -> Newton step in natural
parameters -> damped update -> update_posterior, and ``model.energy()`` returns the scheme's
objective.  The per-step work between the two posterior updates (inference.py:72-86 plus the
``update_variational_params`` body of the scheme) is ONE fused kernel (bn_site_update).
"""
import torch

from . import _lib
from ._util import ptr, stream_ptr, workspace


class InferenceMixin:
    method = None
    power = 1.0

    def _site_args(self, cubature=None):
        a, keep = self.likelihood.site_args(self.method, self.Y, self.posterior_mean, self.posterior_variance,
                                            cubature, self.power)
        a.nat1, a.nat2 = self.pseudo_likelihood.nat1_.data_ptr(), self.pseudo_likelihood.nat2_.data_ptr()
        return a, keep

    def _inference_fused(self, lr, cubature, ensure_psd):
        """the iteration as two fused passes over tiled resident state (fused.py): filter, smoother + site update;
        filter (+ log-likelihood), smoother + the two sums energy() needs, which are kept for it"""
        from . import fused
        st = self._fused_state()
        pl = self.pseudo_likelihood
        _, d = st.run(fused.SITES, self.likelihood, self.method, cubature, lr, self.power, ensure_psd, want_ell=False)
        pl.version += 1  # the sites were rewritten in place (in the tiled arrays: pl reads them back on access)
        pl.source, st.sites_version = st, pl.version
        if fused.linear_posterior():  # the sweep stores the marginals in the reference's [N, 1, 1] layout itself
            ell, sums = st.run(fused.ENERGY, self.likelihood, self.method, cubature, lr, self.power, ensure_psd, want_ell=True,
                               post=(self.posterior_mean, self.posterior_variance))
        else:
            ell, sums = st.run(fused.ENERGY, self.likelihood, self.method, cubature, lr, self.power, ensure_psd, want_ell=True)
            self.posterior_mean, self.posterior_variance = st.posterior(self.posterior_mean, self.posterior_variance)
        ell, sums, d = ell.double(), sums.double(), d.double()  # (no-ops unless the shard runs the fp32 build)
        self._ell_cache = (ell, pl.version, self._hyper_key())
        self._grad_cache = None
        self._energy_cache = (sums[0], sums[1], self._energy_key(cubature))
        n = float(self.num_data)
        return (None, None, None), (d[0] / n, d[1] / n)

    def inference(self, lr=1., batch_ind=None, cubature=None, ensure_psd=True, return_state=False, want_grad=False,
                  **kwargs):
        """one iteration (inference.py:65-90).  Returns ((mean, jacobian, hessian), (diff1, diff2)); the triple (the
        line-search state of the reference) is only formed with return_state=True -- inside the reference's jitted
        train_op it is dead code and never materialised.
        want_grad: the closing posterior update also accumulates the hyper-gradient energy_and_grad() serves."""
        if batch_ind is not None and len(batch_ind) != self.num_data:
            raise NotImplementedError('mini-batched site updates are outside the hot-path scope (SURVEY A.15)')
        fused_ok = not return_state and getattr(self, '_fused_ok', lambda: False)()
        if fused_ok and not want_grad:
            return self._inference_fused(lr, cubature, ensure_psd)
        if fused_ok and getattr(self, '_fused_dtype', torch.float64) == torch.float64:
            # with the hyper-gradient: the first half (update_posterior + site update) as ONE fused pass; the closing
            # update is the stage-level one that forms the adjoint inside its smoother sweep
            from . import fused
            st = self._fused_state()
            pl = self.pseudo_likelihood
            _, d = st.run(fused.SITES, self.likelihood, self.method, cubature, lr, self.power, ensure_psd, want_ell=False)
            pl.version += 1
            pl.source, st.sites_version = st, pl.version
            self._energy_cache = None
            self.update_posterior(want_grad=True)
            n = float(self.num_data)
            return (None, None, None), (d[0] / n, d[1] / n)
        self.update_posterior()
        a, keep = self._site_args(cubature)
        N, D = a.N, a.D
        dev = self.posterior_mean.device
        a.lr, a.ensure_psd = float(lr), int(bool(ensure_psd))
        pl = self.pseudo_likelihood
        a.site_mean, a.site_cov = pl.mean_.data_ptr(), pl.covariance_.data_ptr()
        state = (None, None, None)
        if return_state:
            state = (torch.empty((N, D, 1), dtype=torch.float64, device=dev),
                     torch.empty((N, D, 1), dtype=torch.float64, device=dev),
                     torch.empty((N, D, D), dtype=torch.float64, device=dev))
            a.out_mean, a.out_jac, a.out_hess = (ptr(s) for s in state)
        diffs = torch.zeros((2,), dtype=torch.float64, device=dev)
        a.diffs = diffs.data_ptr()
        ws, nb = workspace(N, getattr(self, '_site_state_dim', self.state_dim), D)
        _lib.check(_lib.lib().bn_site_update(a, ptr(ws), nb, stream_ptr()))
        pl.version += 1  # the sites were rewritten in place
        if want_grad:
            self.update_posterior(want_grad=True)
        else:
            self.update_posterior()
        return state, (diffs[0], diffs[1])

    def expected_density(self, cubature=None):
        """nansum over steps of the scheme's likelihood term (VI: E_q[log p]; Newton: log p(y|m); EP/PL: log Z)"""
        a, keep = self._site_args(cubature)
        out = torch.zeros((), dtype=torch.float64, device=self.posterior_mean.device)
        ws, nb = workspace(a.N, getattr(self, '_site_state_dim', self.state_dim), a.D)
        _lib.check(_lib.lib().bn_expected_density(a, None, ptr(out), ptr(ws), nb, stream_ptr()))
        return out

    def energy(self, batch_ind=None, cubature=None, **kwargs):
        raise NotImplementedError

    def _likelihood_and_kl(self, cubature=None):
        """(likelihood term, KL[q || p]) of the VI / Newton energies; one fused pass over the posterior marginals
        (bn_energy_terms) where the model offers it, the two separate sums otherwise"""
        fused = getattr(self, '_energy_terms_fused', None)
        if fused is not None:
            r = fused(cubature)
            if r is not None:
                return r[0], r[1] - self.compute_log_lik()
        return self.expected_density(cubature), self.compute_kl()

    def energy_and_grad(self, cubature=None):
        """(energy, d energy / d [variance_c...; lengthscale_c...]): the kernel part of what
        objax.GradValues(model.energy, model.vars()) returns (README.md:56-70, demos/regression.py:63-70), for the
        untransformed hyper-parameters; `kernel.chain_to_transformed(grad)` applies the softplus of
        kernels.py:80-95.  In every scheme the kernel hyper-parameters reach the energy only through the filter
        log-likelihood (sites and posterior are StateVars): VI/Newton  E = -(L - (X - ell)),  EP/PL  E = -(ell + ...),
        so d E = - d ell."""
        g = self.log_lik_grad()
        return self.energy(cubature=cubature), -g

    def energy_grad_likelihood(self, cubature=None):
        """d energy / d (likelihood hyper-parameter): the Gaussian variance, the only trainable likelihood parameter on
        the path (likelihoods.py:700-703).  The energy holds it in the likelihood term alone -- VI / Newton
        E = -(L - KL), EP E = -(lZ + (lel - lel_pseudo) / power) -- so one sum over the steps (bn_likelihood_param_grad)
        gives it; `likelihood.chain_to_transformed` applies the softplus of the stored variable."""
        if self.likelihood.lik_id != _lib.BN_LIK_GAUSSIAN:
            raise NotImplementedError('only the Gaussian likelihood has a trainable hyper-parameter here')
        if self.method == _lib.BN_METHOD_PL:
            raise NotImplementedError('likelihood-parameter gradient for posterior linearisation')
        a, keep = self._site_args(cubature)
        out = torch.zeros((), dtype=torch.float64, device=self.posterior_mean.device)
        ws, nb = workspace(a.N, getattr(self, '_site_state_dim', self.state_dim), a.D)
        _lib.check(_lib.lib().bn_likelihood_param_grad(a, ptr(out), ptr(ws), nb, stream_ptr()))
        scale = 1.0 / self.power if self.method == _lib.BN_METHOD_EP else 1.0
        return -scale * out


class VariationalInference(InferenceMixin):
    """natural-gradient VI, CVI form (inference.py:160-222)"""
    method = _lib.BN_METHOD_VI

    def energy(self, batch_ind=None, cubature=None, **kwargs):
        lik, kl = self._likelihood_and_kl(cubature)
        return -(lik - kl)


class Newton(InferenceMixin):
    """Newton = Laplace (inference.py:99-157)"""
    method = _lib.BN_METHOD_NEWTON

    def energy(self, batch_ind=None, cubature=None, **kwargs):
        lik, kl = self._likelihood_and_kl(cubature)
        return -(lik - kl)


Laplace = Newton


class ExpectationPropagation(InferenceMixin):
    """power EP (inference.py:225-325)"""
    method = _lib.BN_METHOD_EP

    def _pseudo_density(self, power, with_const):
        pl = self.pseudo_likelihood
        N, D = pl.mean.shape[0], pl.mean.shape[1]
        out = torch.zeros((), dtype=torch.float64, device=pl.mean.device)
        ws, nb = workspace(N, getattr(self, '_site_state_dim', self.state_dim), D)
        _lib.check(_lib.lib().bn_ep_pseudo_density(
            N, D, float(power), int(with_const), ptr(pl.mean), ptr(pl.covariance), ptr(self.posterior_mean),
            ptr(self.posterior_variance), ptr(pl.nat1), ptr(pl.nat2), ptr(self.mask_pseudo_y), ptr(out), ptr(ws), nb,
            stream_ptr()))
        return out

    def energy(self, batch_ind=None, cubature=None, **kwargs):
        cache = getattr(self, '_energy_cache', None)
        if type(self).method == _lib.BN_METHOD_EP and cache is not None and cache[2] == self._energy_key(cubature):
            lel, lel_pseudo = cache[0], cache[1]  # summed in the smoother epilogue of the pass that closed inference()
        else:
            lel = self.expected_density(cubature)
            lel_pseudo = self._pseudo_density(self.power, True)
        lZ = self.compute_log_lik()
        return -(lZ + 1. / self.power * (lel - lel_pseudo))


class PosteriorLinearisation(ExpectationPropagation):
    """iterated statistical linear regression; its energy is the EP energy at power 1 (inference.py:328-428)"""
    method = _lib.BN_METHOD_PL

    def energy(self, batch_ind=None, cubature=None, **kwargs):
        lZ = self.expected_density(cubature)
        lZ_pseudo = self._pseudo_density(1.0, False)
        return -(self.compute_log_lik() + (lZ - lZ_pseudo))
