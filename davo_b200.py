"""Import alias: the package directory `deep-attention-visual-odometry_b200/` is not a valid Python
identifier, so `import davo_b200` loads it with importlib and registers it under this name."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("deep-attention-visual-odometry_b200")
sys.modules[__name__] = _pkg
