"""JAX side of the boundary: primitives, MLIR custom-call lowering, custom_vjp, and the patch point of the reference.

Nothing here runs in the build image or on the GPU box (neither has jax; the reference pins jax==0.4.14 + objax 1.7).
It is the code a maintainer drops next to the reference to route its hot path through libbn_b200.so:

    import bayesnewton, bayesnewton_b200.jax_glue as glue
    glue.patch(bayesnewton)        # basemodels.kalman_filter / rauch_tung_striebel_smoother -> custom calls
                                   # MarkovGaussianProcess.update_posterior / compute_log_lik -> fused update + custom_vjp

* every C-ABI entry is reached through its header-free legacy custom-call wrapper (include/bn_b200_xla.h,
  csrc/xla.cu; bit-identical to the direct calls: tests/test_xla_wrappers.py) -- jax 0.4.14 predates jax.ffi;
* one jax.core.Primitive per wrapper, abstract-eval from the reference's shapes (ops.py:256-285, 357-380;
  basemodels.py:689-706), lowered with jaxlib.hlo_helpers.custom_call; the last result is the byte workspace XLA
  allocates for the call;
* `filter_log_lik` is a jax.custom_vjp: forward = bn_xla_update_posterior with want_grad (the adjoint is formed inside
  the smoother sweep), backward = cotangent x (d ell / d variance, d ell / d lengthscale), so
  objax.GradValues(model.energy, model.vars()) reaches the hand-written adjoint kernel (README.md:56-70).

Limits, stated plainly: (1) this file has never been executed -- there is no jax in the image; the C wrappers it lowers
to ARE tested (bit-identical to the direct C ABI calls); (2) the kernel hyper-parameters travel in the custom call's opaque
descriptor, i.e. they are lowering-time constants: `filter_log_lik` therefore evaluates them eagerly (jax.device_get), which
works under objax.GradValues outside of objax.Jit; a jitted train_op needs the variant of the wrapper that takes the
hyper-parameters as device operands (the kernels would then prepare the discretisation constants on the device), which is
not built.  The import of jax is deferred to the functions: importing this module is always safe.
"""
import ctypes as C

from . import _lib, xla

_registered = {}
_prims = {}


def available():
    try:
        import jax  # noqa: F401
        import jaxlib  # noqa: F401
        return True
    except ImportError:
        return False


def register():
    """register every bn_xla_* target with XLA's CUDA platform (idempotent)"""
    if not _registered:
        for name in xla.register_with_jax():
            _registered[name] = True
    return sorted(_registered)


def _spec_of(kernel):
    """bn_kernel_spec from a reference kernel object (Matern12/32/52/72 or an Independent stack of one family)"""
    fam = {'Matern12': _lib.BN_MATERN12, 'Matern32': _lib.BN_MATERN32, 'Matern52': _lib.BN_MATERN52,
           'Matern72': _lib.BN_MATERN72}
    parts = getattr(kernel, 'kernels', [kernel])
    names = {type(k).__name__ for k in parts}
    if len(names) != 1 or next(iter(names)) not in fam:
        raise NotImplementedError('no in-library discretisation for %s' % sorted(names))
    return _lib.kernel_spec(fam[next(iter(names))], [float(k.variance) for k in parts], [float(k.lengthscale) for k in parts])


def _primitive(name, n_out, abstract):
    """a multiple-results primitive whose CUDA lowering is the custom call `name`"""
    import jax
    from jax.interpreters import mlir
    from jaxlib.hlo_helpers import custom_call
    if name in _prims:
        return _prims[name]
    p = jax.core.Primitive(name)
    p.multiple_results = True
    p.def_impl(lambda *a, **k: jax.interpreters.xla.apply_primitive(p, *a, **k))
    p.def_abstract_eval(abstract)

    def lowering(ctx, *operands, opaque, **_):
        out_types = [mlir.aval_to_ir_type(a) for a in ctx.avals_out]
        return custom_call(name, out_types, list(operands), backend_config=opaque,
                           operand_layouts=[tuple(range(len(a.shape) - 1, -1, -1)) for a in ctx.avals_in],
                           result_layouts=[tuple(range(len(a.shape) - 1, -1, -1)) for a in ctx.avals_out])
    mlir.register_lowering(p, lowering, platform='cuda')
    _prims[name] = p
    return p


def _avals(*shapes_dtypes):
    import jax
    import jax.numpy as jnp
    return [jax.core.ShapedArray(s, jnp.dtype(d)) for s, d in shapes_dtypes]


def update_posterior(dt, kernel, pseudo_y, pseudo_var, mask=None, want_grad=False):
    """MarkovGaussianProcess.update_posterior (basemodels.py:689-706) as ONE custom call:
    -> ell, post_mean [N,D,1], post_cov [N,D,D] (+ d ell / d variance [NC], d ell / d lengthscale [NC])"""
    register()
    spec = _spec_of(kernel)
    N, D = pseudo_y.shape[0], pseudo_y.shape[1]
    nb = int(_lib.lib().bn_update_posterior_workspace_bytes(spec, N))
    opaque = xla.markov_desc(spec, N, nb, has_mask=mask is not None, want_grad=want_grad)
    NC = spec.n_components

    def abstract(*_, **__):
        outs = [((), 'float64'), ((N, D, 1), 'float64'), ((N, D, D), 'float64')]
        if want_grad:
            outs += [((NC,), 'float64'), ((NC,), 'float64')]
        return _avals(*outs, ((nb,), 'uint8'))
    p = _primitive('bn_xla_update_posterior', 4 + 2 * int(want_grad), abstract)
    ops = [dt, pseudo_y, pseudo_var] + ([mask.astype('uint8')] if mask is not None else [])
    return p.bind(*ops, opaque=opaque)[:-1]


def kalman_filter(dt, kernel, y, noise_cov, mask=None, parallel=False, return_predict=False):
    """ops.kalman_filter (ops.py:256-285): ell, (means, covs)"""
    register()
    spec = _spec_of(kernel)
    N, D = y.shape[0], y.shape[1]
    d = int(_lib.lib().bn_state_dim(spec))
    nb = int(_lib.lib().bn_workspace_bytes(N, d, D))
    opaque = xla.markov_desc(spec, N, nb, form=_lib.BN_SCAN if parallel else _lib.BN_SEQUENTIAL,
                             has_mask=mask is not None, return_predict=return_predict)
    p = _primitive('bn_xla_kalman_filter', 4, lambda *_, **__: _avals(((), 'float64'), ((N, d, 1), 'float64'),
                                                                      ((N, d, d), 'float64'), ((nb,), 'uint8')))
    ops = [dt, y, noise_cov] + ([mask.astype('uint8')] if mask is not None else [])
    ell, m, P, _ = p.bind(*ops, opaque=opaque)
    return ell, (m, P)


def rauch_tung_striebel_smoother(dt, kernel, filter_mean, filter_cov, return_full=False, parallel=False):
    """ops.rauch_tung_striebel_smoother (ops.py:357-380): means, covs, gains"""
    register()
    spec = _spec_of(kernel)
    N, d = filter_mean.shape[0], filter_mean.shape[1]
    D = d if return_full else spec.n_components
    nb = int(_lib.lib().bn_workspace_bytes(N, d, spec.n_components))
    opaque = xla.markov_desc(spec, N, nb, form=_lib.BN_SCAN if parallel else _lib.BN_SEQUENTIAL, return_full=return_full)
    p = _primitive('bn_xla_rts_smoother', 4, lambda *_, **__: _avals(((N, D, 1), 'float64'), ((N, D, D), 'float64'),
                                                                     ((N, d, d), 'float64'), ((nb,), 'uint8')))
    return tuple(p.bind(dt, filter_mean, filter_cov, opaque=opaque)[:-1])


def make_filter_log_lik(kernel_cls_name, n_components=1):
    """compute_log_lik (basemodels.py:726-741) as a function of the hyper-parameter VALUES with a hand-written VJP:

        ell = filter_log_lik(variance[NC], lengthscale[NC], dt, pseudo_y, pseudo_var)

    forward: bn_xla_update_posterior with want_grad (ell and both gradients from one fused pass); backward: the
    cotangent times the stored gradients.  dt, the sites and (by the reference's construction) the posterior get no
    cotangent: they are StateVars / data."""
    import jax
    import jax.numpy as jnp
    fam = {'Matern12': _lib.BN_MATERN12, 'Matern32': _lib.BN_MATERN32, 'Matern52': _lib.BN_MATERN52,
           'Matern72': _lib.BN_MATERN72}[kernel_cls_name]

    class _K:  # the minimal kernel view _spec_of needs, built from traced-out concrete values at call time
        def __init__(self, v, l):
            self.kernels = [type(kernel_cls_name, (), {'variance': float(a), 'lengthscale': float(b)})() for a, b in zip(v, l)]

    @jax.custom_vjp
    def filter_log_lik(variance, lengthscale, dt, pseudo_y, pseudo_var):
        return _fwd(variance, lengthscale, dt, pseudo_y, pseudo_var)[0]

    def _fwd(variance, lengthscale, dt, pseudo_y, pseudo_var):
        # hyper-parameters are concrete inside objax.GradValues' forward evaluation of the energy: the custom call takes
        # them by value in its descriptor (they are not traced operands)
        k = _K(jax.device_get(variance), jax.device_get(lengthscale))
        ell, _, _, dvar, dlen = update_posterior(dt, k, pseudo_y, pseudo_var, want_grad=True)
        return ell, (dvar, dlen)

    def _bwd(res, ct):
        dvar, dlen = res
        return ct * dvar, ct * dlen, None, None, None

    filter_log_lik.defvjp(_fwd, _bwd)
    filter_log_lik.family = fam
    return filter_log_lik


def patch(bayesnewton):
    """route the reference's hot path through the library: the names `basemodels` imported from `ops`
    (basemodels.py:28-44) are rebound, so MarkovGaussianProcess.filter / .smoother (basemodels.py:655-661) dispatch here"""
    register()
    bm = bayesnewton.basemodels
    bm.kalman_filter = kalman_filter
    bm.rauch_tung_striebel_smoother = rauch_tung_striebel_smoother
    return bm
