"""Exception types of the clustering path, named as in the reference
(/root/reference/enspara/exception.py) so that callers' ``except`` clauses keep working."""


class ImproperlyConfigured(Exception):
    """A configuration (estimator arguments, metric name, CLI flags) cannot be used."""


class DataInvalid(Exception):
    """Input data is structurally wrong: shapes, lengths or dtypes do not fit together."""


class InsufficientResourceError(Exception):
    """Valid request, but it does not fit the available device memory."""
