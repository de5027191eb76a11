"""Importable alias of the product package, whose directory name (``v-diffusion-torch_b200``,
fixed by the repo layout contract) is not a valid Python identifier."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "v-diffusion-torch_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
