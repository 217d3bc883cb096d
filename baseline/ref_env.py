"""Environment shims for running the UNMODIFIED reference (baseline/_ref) in this image - nothing here touches its code.

* kornia is not installed: a stub module (only paste_front / the loss use it).
* torch >= 2.x: `torch.utils.cpp_extension.load` no longer leaves the built extension importable by name, but the reference's
  `custom_ops.get_plugin` (custom_ops.py:144) does `importlib.import_module(module_name)` right after it.  A meta-path finder
  resolves `<name>_plugin` to the `.so` under TORCH_EXTENSIONS_DIR/<name>_plugin/<digest>-<gpu>/ that the reference itself
  just built (or that baseline/install_ref.sh pre-built).
"""
import glob
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TREE = os.path.join(ROOT, 'baseline', '_ref')


class _PluginFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path=None, target=None):
        if not name.endswith('_plugin'):
            return None
        ext = os.environ.get('TORCH_EXTENSIONS_DIR', '')
        hits = sorted(glob.glob(os.path.join(ext, name, '*', name + '.so')), key=os.path.getmtime)
        if not hits:
            return None
        return importlib.util.spec_from_file_location(name, hits[-1], loader=importlib.machinery.ExtensionFileLoader(name, hits[-1]))


def setup(need_tree=True):
    """PROJECT_DN, sys.path, TORCH_EXTENSIONS_DIR, the kornia stub and the plugin finder.  Returns REF_TREE."""
    if need_tree and not os.path.isdir(os.path.join(REF_TREE, '_train', 'eg3dc', 'src', 'training')):
        raise SystemExit('baseline/_ref is missing: run `bash baseline/install_ref.sh` in the build container')
    os.environ['PROJECT_DN'] = REF_TREE
    os.environ.setdefault('TORCH_EXTENSIONS_DIR', os.path.join(REF_TREE, '_torch_ext'))
    for q in (os.path.join(REF_TREE, '_train', 'eg3dc', 'src'), REF_TREE, ROOT):
        if q not in sys.path:
            sys.path.insert(0, q)
    sys.modules.setdefault('kornia', types.ModuleType('kornia'))
    if not any(isinstance(f, _PluginFinder) for f in sys.meta_path):
        sys.meta_path.append(_PluginFinder())
    return REF_TREE
