"""Importable alias for the package directory ``explicit-alignment-for-vqa-tasks_b200/``.

The directory name required by the repo layout contains hyphens and therefore is
not a Python identifier; ``import eavqa_b200`` loads that directory as the package
``eavqa_b200`` (sub-modules resolve through its ``__path__``).
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "explicit-alignment-for-vqa-tasks_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
