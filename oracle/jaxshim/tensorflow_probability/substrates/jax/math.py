from scipy.special import ive as bessel_ive  # noqa: F401
