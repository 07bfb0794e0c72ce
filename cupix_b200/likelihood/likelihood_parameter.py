"""Legacy parameter objects of the previous cupix API generation
(``cupix/likelihood/likelihood_parameter.py``), kept so that callers passing
``get_chi2(like_params=[LikelihoodParameter, ...])`` keep working.  Only the pieces the hot path
touches are provided: the container, the unit-cube mapping and the list <-> dict helpers."""


class LikelihoodParameter(object):

    def __init__(self, name, min_value=None, max_value=None, ini_value=None, value=None,
                 Gauss_priors_width=None, fixed=False):
        self.name, self.value = name, value
        self.min_value, self.max_value = min_value, max_value
        self.Gauss_priors_width = Gauss_priors_width
        self.fixed = False          # sic: the reference ignores the argument too (:24)
        if ini_value is not None:
            self.ini_value = ini_value

    def _span(self):
        return self.max_value - self.min_value

    def value_in_cube(self):
        assert self.value is not None, "value not set in parameter " + self.name
        return (self.value - self.min_value) / self._span()

    def value_from_cube(self, x):
        return self.min_value + x * self._span()

    def set_from_cube(self, x):
        self.value = self.value_from_cube(x)

    def is_same_parameter(self, param):
        return self.name == param.name


def dict_from_likeparam(like_params):
    """list of LikelihoodParameter (or an already flat dict) -> {name: value}."""
    if isinstance(like_params, dict):
        return dict(like_params)
    return {p.name: p.value for p in like_params}


def likeparam_from_dict(params_dict):
    return [LikelihoodParameter(name=k, value=v) for k, v in params_dict.items()]
