"""Which filter log-likelihood is right at N = 1e7?  fused update vs stand-alone scan filter vs a long-double
per-step re-evaluation on the host from the GPU's own filtered states (run on the GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # tests/ab -> repo root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bayesnewton_b200 as bn
from _data import bench_inputs
from oracle import ssm

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
t, dt, y = bench_inputs(N)
kg = bn.kernels.Matern52(1.0, 1.0)
m = bn.models.MarkovVariationalGP(kernel=kg, likelihood=bn.likelihoods.Bernoulli(), X=t, Y=y, parallel=True)
m.inference(lr=1.0, return_state=False)
py, pv = m.pseudo_likelihood.mean, m.pseudo_likelihood.covariance
e_fused, _, _ = bn.ops.update_posterior(m.dt, kg, py, pv, want_ell=True)
e_scan, (fm, fP) = bn.ops.kalman_filter(m.dt, kg, py, pv, parallel=True)
e_scan_rp, (pm_, pP_) = bn.ops.kalman_filter(m.dt, kg, py, pv, parallel=True, return_predict=True)
print('fused %.17g\nscan  %.17g  (diff %.3e)' % (float(e_fused), float(e_scan), float(e_fused) - float(e_scan)))
# host long-double evaluation of sum_k log N(y_k | H m_k^-, H P_k^- H^T + R_k) from the predicted states
mp = pm_.cpu().numpy()[:, 0, 0].astype(np.longdouble)
Pp = pP_.cpu().numpy()[:, 0, 0].astype(np.longdouble)
yy = py.cpu().numpy().reshape(-1).astype(np.longdouble)
RR = pv.cpu().numpy().reshape(-1).astype(np.longdouble)
S = Pp + RR
terms = -0.5 * ((yy - mp) ** 2 / S + np.log(2 * np.longdouble(np.pi)) + np.log(S))
import math
print('host  %.17g  (long double terms from the predicted states of the scan filter, fsum)' % math.fsum(terms.astype(np.float64)))
print('host longdouble sum %.17g' % float(terms.sum()))
