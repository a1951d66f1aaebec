"""Import alias for the package in ``semantic-nerf-for-satellite-data_b200/`` (a directory name that
is not a valid Python identifier).  ``import semnerf_b200.renderer`` resolves inside that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "semantic-nerf-for-satellite-data_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
