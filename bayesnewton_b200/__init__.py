"""bayesnewton_b200: the Markov-GP inference hot path of AaltoML/BayesNewton (Kalman filter, RTS
smoother, state-space discretisation, cubature site updates) as hand-written CUDA for sm_100a behind
a C ABI (include/bn_b200.h), with a host-side mirror of the reference's ops / kernels / likelihoods /
inference / models interfaces.  There is no CPU path: everything O(N) runs in libbn_b200.so."""
from . import _lib, cubature, inference, kernels, likelihoods, models, ops, sparse, spacetime  # noqa: F401
from .models import build_model  # noqa: F401

__version__ = '0.1.0'
