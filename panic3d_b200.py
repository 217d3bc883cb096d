"""Import alias: the package directory is named ``panic3d-anime-reconstruction_b200`` (not a valid
Python identifier), so ``import panic3d_b200`` loads that directory as the package ``panic3d_b200``."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), 'panic3d-anime-reconstruction_b200')
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_dir, '__init__.py'), submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod          # `import panic3d_b200` hands back this package object
_spec.loader.exec_module(_mod)
