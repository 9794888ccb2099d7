"""Import shim: makes the hyphenated package directory ``ps-slm_b200/`` importable as
``ps_slm_b200`` (same trick as the reference's ``model/ps-slm.py`` loader,
Multitask/utils/dataset_utils.py:14-25).  ``import ps_slm_b200.bridge`` etc. work as usual."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "ps-slm_b200")]
__package__ = __name__
if __spec__ is not None:                       # make the module a regular package for importlib
    __spec__.submodule_search_locations = __path__
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
