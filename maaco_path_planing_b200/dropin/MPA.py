"""Drop-in for the reference module `MPA` (same module and class names): put
`maaco_path_planing_b200/dropin` on sys.path instead of the reference directory and
`from MPA import MPA` resolves to the B200 implementation."""
import _bootstrap  # noqa: F401
from maaco_path_planing_b200.mpa import MPA  # noqa: F401,E402
