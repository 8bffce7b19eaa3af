"""Import helper: the product package lives in `saprobe-alac_b200/` (hyphen, not importable by name)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
_NAME = 'saprobe_alac_b200'


def load_package():
    """Import saprobe-alac_b200/ as module `saprobe_alac_b200` (fails loudly if libalacb200.so is missing)."""
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    path = os.path.join(ROOT, 'saprobe-alac_b200', '__init__.py')
    spec = importlib.util.spec_from_file_location(_NAME, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        del sys.modules[_NAME]
        raise
    return mod
