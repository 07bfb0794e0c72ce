"""Legacy parameter objects (cupix/likelihood/likelihood_parameter.py), kept so that
``get_chi2(like_params=[LikelihoodParameter, ...])`` callers keep working."""


class LikelihoodParameter(object):
    def __init__(self, name, min_value=None, max_value=None, ini_value=None, value=None,
                 Gauss_priors_width=None, fixed=False):
        self.name = name
        self.min_value = min_value
        self.max_value = max_value
        self.value = value
        self.Gauss_priors_width = Gauss_priors_width
        self.fixed = False          # sic: the reference ignores the argument (likelihood_parameter.py:24)
        if ini_value is not None:
            self.ini_value = ini_value

    def value_in_cube(self):
        assert self.value is not None, "value not set in parameter " + self.name
        return (self.value - self.min_value) / (self.max_value - self.min_value)

    def value_from_cube(self, x):
        return self.min_value + x * (self.max_value - self.min_value)

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
