"""Writes tests/golden/markov_small.npz: seeded inputs and oracle outputs for the hot path.

The reference ships no golden vectors and cannot run in this image (no jax), so these fixtures are
generated from the CPU oracle (oracle/), which tests/test_oracle_pinning.py ties to the reference's
own dense-GP cross-checks.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from _data import classification_data, filter_problem  # noqa: E402
from oracle import kalman, model, sites, ssm  # noqa: E402


def build_cases():
    out = {}
    # filter + smoother, every family, sequential form
    for name, k, D in [('m12', ssm.Matern12(0.8, 1.7), 1), ('m32', ssm.Matern32(1.1, 0.6), 1),
                       ('m52', ssm.Matern52(1.3, 0.9), 1), ('m72', ssm.Matern72(0.7, 1.4), 1),
                       ('ind32', ssm.Independent([ssm.Matern32(1.0, 1.0), ssm.Matern32(0.5, 2.0)]), 2)]:
        dt, y, R, mask = filter_problem(97, D=D, seed=7)
        ell, (fm, fP) = kalman.kalman_filter(dt, k, y, R, mask)
        dts = np.concatenate([dt[1:], [0.0]])
        sm, sP, G = kalman.rauch_tung_striebel_smoother(dts, k, fm, fP)
        for key, val in dict(dt=dt, y=y, R=R, mask=mask, ell=ell, fm=fm, fP=fP, sm=sm, sP=sP, gains=G).items():
            out['%s_%s' % (name, key)] = np.asarray(val)
    # one VI / EP iteration of the classification model (reference test grid point var_f=1.5, len_f=0.75, N=60)
    x, y = classification_data(60)
    for method in ('vi', 'ep', 'newton', 'pl'):
        m = model.MarkovGP(ssm.Matern52(1.5, 0.75), sites.Bernoulli(), x, y, method=method, power=0.5)
        m.inference(lr=0.7)
        m.inference(lr=0.7)
        out['cls_%s_post_mean' % method] = m.post_mean
        out['cls_%s_post_cov' % method] = m.post_cov
        out['cls_%s_site_nat1' % method] = m.site_nat1
        out['cls_%s_site_nat2' % method] = m.site_nat2
        out['cls_%s_energy' % method] = np.asarray(m.energy())
    out['cls_x'], out['cls_y'] = x, y
    return out


if __name__ == '__main__':
    np.savez_compressed(os.path.join(HERE, 'markov_small.npz'), **build_cases())
    print('wrote', os.path.join(HERE, 'markov_small.npz'))
