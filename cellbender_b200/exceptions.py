"""Error types of the hot path.  Names, base classes and attributes follow the reference
(cellbender/remove_background/exceptions.py:4-24) because its callers catch them by type and read ``.param`` /
``.message`` (train.py:277-283, run.py:823); unlike the reference's, these also carry the text as ``str(exc)``."""


class _Described:
    """Keeps the human-readable text in ``message`` and hands it to the built-in exception as well."""

    def _describe(self, text: str) -> None:
        self.message = text
        self.args = (text,)


class NanException(_Described, ArithmeticError):
    """The likelihood produced a NaN; ``param`` names the offending tensors (raised where the reference raises it,
    distributions/NegativeBinomialPoissonConvApprox.py:129-141)."""

    def __init__(self, param: str):
        ArithmeticError.__init__(self)
        self.param = param
        self._describe(f"A wild NaN appeared!  In param {{{param}}}")


class ElboException(_Described, ValueError):
    """Training failed one of the test-ELBO criteria (train.py:211-257, 287-295); ``message`` goes to the log."""

    def __init__(self, message: str):
        ValueError.__init__(self)
        self._describe(message)
