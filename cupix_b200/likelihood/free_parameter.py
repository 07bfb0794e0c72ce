"""Free parameter of a fit (mirror of cupix/likelihood/free_parameter.py)."""


class FreeParameter(object):

    def __init__(self, name, min_value=None, max_value=None, ini_value=None, true_value=None,
                 delta=None, gauss_prior_mean=None, gauss_prior_width=None, latex_label=None):
        self.name = name
        self.min_value = min_value
        self.max_value = max_value
        self.ini_value = ini_value
        self.true_value = true_value
        self.delta = delta
        self.gauss_prior_mean = gauss_prior_mean
        self.gauss_prior_width = gauss_prior_width
        self.latex_label = self.name if latex_label is None else latex_label

    def out_of_bounds(self, value):
        return bool(value < self.min_value or value > self.max_value)

    def get_prior_chi2(self, value):
        if self.gauss_prior_mean is None and self.gauss_prior_width is None:
            return 0.0
        assert self.gauss_prior_mean is not None
        assert self.gauss_prior_width is not None
        return ((value - self.gauss_prior_mean) / self.gauss_prior_width) ** 2
