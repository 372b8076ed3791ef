"""Import shim: the package lives in the directory ``mm-pde_b200/`` (not a valid Python identifier),
so this module turns itself into that package: ``import mmpde_b200.gnn_2d`` resolves inside it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "mm-pde_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
