"""Importable alias of the hyphenated package directory ``transformer-pointer-critic_b200/``.

The build contract names the package directory ``transformer-pointer-critic_b200`` (not a valid Python
identifier), so this shim points ``tpc_b200.__path__`` at it and runs its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "transformer-pointer-critic_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
