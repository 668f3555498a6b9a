"""Makes the repo root importable when this directory is put on sys.path / PYTHONPATH in place of the
reference's source directory."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
