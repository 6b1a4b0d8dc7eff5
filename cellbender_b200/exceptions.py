"""Exceptions mirroring the reference's (cellbender/remove_background/exceptions.py:4-24): same names,
same attributes, so callers that catch them (train.py:277-283, run.py:823) keep working."""


class NanException(ArithmeticError):
    """A NaN appeared in the likelihood (raised where the reference raises it,
    distributions/NegativeBinomialPoissonConvApprox.py:129-141)."""

    def __init__(self, param: str):
        self.param = param
        self.message = "A wild NaN appeared!  In param {" + self.param + "}"
        super().__init__(self.message)


class ElboException(ValueError):
    """Training failed a test-ELBO criterion (train.py:211-257, 287-295)."""

    def __init__(self, message: str):
        self.message = message
        super().__init__(message)
