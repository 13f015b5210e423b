"""Exception types of the drop-in path.

Same names and hierarchy as the reference (nn_fac/utils/errors.py:8-18), including its choice of
BaseException as the root, because callers and the reference's tests catch these exact classes.
"""


def _make(name, base, doc):
    return type(name, (base,), {"__doc__": doc, "__module__": __name__})


ArgumentException = _make("ArgumentException", BaseException, "An argument has an invalid shape, type or value.")
InvalidRanksException = _make("InvalidRanksException", ArgumentException, "Ranks do not match the tensor order.")
CustomNotEngouhFactors = _make("CustomNotEngouhFactors", ArgumentException, "Custom init with too few factors.")
CustomNotValidFactors = _make("CustomNotValidFactors", ArgumentException, "Custom init with a missing factor.")
CustomNotValidCore = _make("CustomNotValidCore", ArgumentException, "Custom init with a missing core.")
InvalidInitializationType = _make("InvalidInitializationType", ArgumentException, "Unknown init type.")
InvalidArgumentValue = _make("InvalidArgumentValue", ArgumentException, "An argument value is out of its domain.")

OptimException = _make("OptimException", BaseException, "The optimisation cannot proceed.")
ZeroColumnWhenUnautorized = _make("ZeroColumnWhenUnautorized", OptimException,
                                  "A zero column/diagonal was met while nonzero=True.")
