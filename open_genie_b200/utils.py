"""Small helpers mirroring genie/utils.py:13-19 (exists / default / Blueprint)."""
from typing import Tuple, TypeVar

T = TypeVar('T')
D = TypeVar('D')

Blueprint = Tuple[str | Tuple[str, dict], ...]


def exists(var) -> bool:
    return var is not None


def default(var, val):
    return var if exists(var) else val
