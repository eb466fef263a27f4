"""Importable alias of the `ctrl-v_b200/` package directory (a hyphen is not a valid Python
identifier, so the code lives in `ctrl-v_b200/` and is imported as `ctrlv_b200`)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ctrl-v_b200")
__path__.append(_real)
__version__ = "0.1.0"
