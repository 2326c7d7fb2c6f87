"""Import shim: the package directory is `xfg-stark_b200/` (the name the build contract asks for), which Python cannot
import by name because of the hyphen.  `import xfg_stark_b200` loads that directory as this module."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "xfg-stark_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
