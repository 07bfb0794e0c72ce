"""Legacy parameter objects of the previous cupix API generation
(``cupix/likelihood/likelihood_parameter.py``), kept so that callers passing
``get_chi2(like_params=[LikelihoodParameter, ...])`` keep working: the container with its unit-cube
mapping (likelihood_parameter.py:4-82) and the module-level list / dict helpers (:84-141)."""
import numpy as np



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

    def get_value_in_cube(self, value):
        """``value`` mapped to [0, 1] between the bounds (the stored value is not touched)."""
        return (value - self.min_value) / self._span()

    def value_from_cube(self, x):
        return self.min_value + x * self._span()

    def err_from_cube(self, err):
        """An error (or any length) given in cube units, in parameter units."""
        return err * self._span()

    def set_from_cube(self, x):
        self.value = self.value_from_cube(x)

    def set_without_cube(self, value):
        """Set the value directly; it must lie strictly inside the bounds."""
        assert self.min_value < value < self.max_value, "Parameter name: %s" % self.name
        self.value = value

    def get_new_parameter(self, value_in_cube):
        """A copy of this parameter (same bounds and prior width) whose value is ``value_in_cube`` mapped back."""
        twin = LikelihoodParameter(self.name, self.min_value, self.max_value, Gauss_priors_width=self.Gauss_priors_width)
        twin.set_from_cube(value_in_cube)
        return twin

    def info_str(self, all_info=False):
        """``name = value`` (with the bounds appended if asked), for log lines."""
        parts = ["%s = %s" % (self.name, self.value)]
        if all_info:
            parts += [str(self.min_value), str(self.max_value)]
        return " , ".join(parts)

    def is_same_parameter(self, param):
        return self.name == param.name


def dict_from_likeparam(like_params):
    """list of LikelihoodParameter (or an already flat dict) -> {name: value}; None -> {}."""
    if like_params is None:
        return {}
    if isinstance(like_params, dict):
        return dict(like_params)
    return {p.name: p.value for p in like_params}


def likeparam_from_dict(params_dict):
    """{name: value} -> list of LikelihoodParameter with the reference's catch-all bounds (-1000, 1000); None -> []."""
    return [LikelihoodParameter(name=k, min_value=-1000, max_value=1000, value=v) for k, v in (params_dict or {}).items()]


def format_like_params_dict(iz_choice, param_dict):
    """Keys in the ``<name>_<iz>`` form the multi-z likelihood of the old API expects (likelihood_parameter.py:106-124):
    keys already ending in ``_<iz>`` for a chosen bin are kept; with a single chosen bin bare names get its suffix;
    with several bins bare names are dropped with a warning."""
    bins = np.atleast_1d(iz_choice)
    out = {}
    for iz in bins:
        suffix = "_%d" % iz
        for key, value in param_dict.items():
            if key.endswith(suffix):
                out[key] = value
            elif bins.size == 1:
                out[key + suffix] = value
            else:
                print("Warning: parameter", key, "not in correct format, must end in _{integer} to specify redshift bin "
                      "when evaluating multiple redshift bins. This parameter will be ignored.")
    return out


def index_by_name(like_params, pname):
    """Position of the parameter called ``pname`` in a list of parameters (IndexError if absent)."""
    return [i for i, p in enumerate(like_params) if p.name == pname][0]


def like_parameter_by_name(like_params, pname):
    return like_params[index_by_name(like_params, pname)]


def par_index(like_param_list, par_name):
    """Like ``index_by_name`` but with a ValueError naming the parameter when it is absent."""
    for i, p in enumerate(like_param_list):
        if p.name == par_name:
            return i
    raise ValueError("Parameter with name " + par_name + " not found in list of LikelihoodParameter objects.")
