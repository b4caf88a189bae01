"""Writes tests/golden/spacetime_small.npz: seeded inputs and oracle outputs of one VI iteration of the
spatio-temporal Markov GP (3 x 3 spatial grid, 12 time steps, 10 % missing observations).

The reference cannot run in this image (no jax); the fixture comes from oracle/spacetime.py, which
tests/test_spacetime.py::test_oracle_markov_vs_dense_gp ties to the dense GP on the product kernel the way the
reference's own test does (tests/test_gp_vs_markovgp_spacetime.py).  Run:  python tests/golden/make_golden_spacetime.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from test_spacetime import GOLDEN, golden_case  # noqa: E402

if __name__ == '__main__':
    _, out = golden_case()
    np.savez_compressed(GOLDEN, **out)
    print('wrote', GOLDEN, {k: v.shape for k, v in out.items()})
