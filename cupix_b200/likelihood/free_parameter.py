"""Free parameter of a fit: name, bounds, start value, optional Gaussian prior.

Interface mirror of ``cupix/likelihood/free_parameter.py`` (same constructor keywords and the two
methods ``Posterior`` calls), so fit scripts written against the reference keep working."""

_FIELDS = ("min_value", "max_value", "ini_value", "true_value", "delta",
           "gauss_prior_mean", "gauss_prior_width")


class FreeParameter(object):

    def __init__(self, name, min_value=None, max_value=None, ini_value=None, true_value=None,
                 delta=None, gauss_prior_mean=None, gauss_prior_width=None, latex_label=None):
        given = locals()
        self.name = name
        for field in _FIELDS:
            setattr(self, field, given[field])
        self.latex_label = latex_label if latex_label is not None else name

    def out_of_bounds(self, value):
        """True when ``value`` lies outside [min_value, max_value]."""
        return not (self.min_value <= value <= self.max_value)

    def get_prior_chi2(self, value):
        """((value - mean) / width)^2, or 0 without a prior; giving only one of the two is an error."""
        mean, width = self.gauss_prior_mean, self.gauss_prior_width
        if mean is None and width is None:
            return 0.0
        assert mean is not None and width is not None, "Gaussian prior needs both mean and width"
        return ((value - mean) / width) ** 2
