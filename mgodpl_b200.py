"""Import alias: the package directory is named after the reference repository and contains hyphens."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("multigoal-orchard-drone-planning-library_b200")
sys.modules[__name__] = _pkg
