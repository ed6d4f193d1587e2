"""Import shim: registers the hyphenated package directory `stereo-visual-slam_b200/` as the module
`stereo_visual_slam_b200` (a hyphen is not a valid identifier)."""
import importlib.util
import os
import sys

_NAME = "stereo_visual_slam_b200"
_ROOT = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_ROOT, "stereo-visual-slam_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                  submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = load()
