"""Import shim: the package directory is named `flow-guided-krylov_b200/` (not an
importable identifier); this module maps it to `flow_guided_krylov_b200`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "flow-guided-krylov_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
