"""Model classes = inference scheme x Markov GP, assembled by multiple inheritance exactly like the
reference's glue file (bayesnewton/models.py:118-152, build_model in bayesnewton/__init__.py:13-14)."""
from .basemodels import InfiniteHorizonGaussianProcess, MarkovGaussianProcess, MarkovMeanFieldGaussianProcess
from .sparse import SparseMarkovGaussianProcess
from .inference import ExpectationPropagation, Newton, PosteriorLinearisation, VariationalInference


class MarkovVariationalGP(VariationalInference, MarkovGaussianProcess):
    pass


class MarkovExpectationPropagationGP(ExpectationPropagation, MarkovGaussianProcess):
    def __init__(self, kernel, likelihood, X, Y, R=None, power=1., parallel=None):
        self.power = power
        super().__init__(kernel, likelihood, X, Y, R=R, parallel=parallel)


class MarkovNewtonGP(Newton, MarkovGaussianProcess):
    pass


MarkovLaplaceGP = MarkovNewtonGP


class MarkovPosteriorLinearisationGP(PosteriorLinearisation, MarkovGaussianProcess):
    pass


class MarkovVariationalMeanFieldGP(VariationalInference, MarkovMeanFieldGaussianProcess):
    """models.py:154 of the reference"""
    pass


class InfiniteHorizonVariationalGP(VariationalInference, InfiniteHorizonGaussianProcess):
    """models.py:162 of the reference"""
    pass


class InfiniteHorizonExpectationPropagationGP(ExpectationPropagation, InfiniteHorizonGaussianProcess):
    """models.py:261 of the reference"""

    def __init__(self, kernel, likelihood, X, Y, R=None, power=1., dare_iters=20, parallel=None):
        self.power = power
        super().__init__(kernel, likelihood, X, Y, R=R, dare_iters=dare_iters, parallel=parallel)


class InfiniteHorizonNewtonGP(Newton, InfiniteHorizonGaussianProcess):
    """models.py:315 of the reference"""
    pass


class InfiniteHorizonPosteriorLinearisationGP(PosteriorLinearisation, InfiniteHorizonGaussianProcess):
    """models.py:388 of the reference"""
    pass


class SparseMarkovVariationalGP(SparseMarkovGaussianProcess):
    """VI on the sparse Markov GP (models.py of the reference); the VI iteration is built into the base class"""
    pass


def build_model(model, inf):
    """dynamic glue, as bayesnewton.build_model"""
    return type(inf.__name__ + model.__name__, (inf, model), {})
