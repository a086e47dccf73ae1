"""Import alias: `mcgra_b200` resolves to the package directory `mc-gra_b200/` (whose name, fixed by the
project layout, is not a valid Python identifier)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mc-gra_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
