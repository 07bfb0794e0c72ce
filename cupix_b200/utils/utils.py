"""Small host helpers (counterpart of cupix/utils/utils.py:73-136)."""
import os


def get_path_repo(name="cupix_b200"):
    """Directory holding the package's data files."""
    return os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def data_path(*parts):
    return os.path.join(get_path_repo(), "data", *parts)


def is_rank0():
    return int(os.environ.get("RANK", "0")) == 0


def print0(*args, **kwargs):
    if is_rank0():
        print(*args, **kwargs)
