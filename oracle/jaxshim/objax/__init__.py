"""The slice of objax the reference uses: Module, StateVar / TrainVar (a value holder), random.  TEST INFRASTRUCTURE."""
import numpy as _np


class BaseVar:
    def __init__(self, value):
        self._value = value

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, v):
        self._value = v

    def assign(self, v):
        self._value = v


class StateVar(BaseVar):
    pass


class TrainVar(BaseVar):
    pass


class Module:
    def vars(self):
        out = {}
        for k, v in self.__dict__.items():
            if isinstance(v, BaseVar):
                out[k] = v
            elif isinstance(v, Module):
                for kk, vv in v.vars().items():
                    out[k + '.' + kk] = vv
        return out


class ModuleList(list, Module):
    pass


class _Random:
    class Generator:
        def __init__(self, seed=0):
            self.rng = _np.random.default_rng(seed)

    DEFAULT_GENERATOR = None

    @staticmethod
    def normal(shape, mean=0.0, stddev=1.0, generator=None):
        rng = generator.rng if generator is not None else _np.random.default_rng(0)
        return mean + stddev * rng.standard_normal(shape)

    @staticmethod
    def uniform(shape, generator=None):
        rng = generator.rng if generator is not None else _np.random.default_rng(0)
        return rng.random(shape)


random = _Random()


class _Loss:
    @staticmethod
    def cross_entropy_logits_sparse(logits, labels):
        raise NotImplementedError


class _Functional:
    loss = _Loss()


functional = _Functional()
