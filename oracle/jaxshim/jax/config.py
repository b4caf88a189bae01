from . import config  # noqa: F401  (`from jax.config import config`)
