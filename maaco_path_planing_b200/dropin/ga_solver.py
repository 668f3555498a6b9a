"""Drop-in for the reference module `ga_solver` (same module and class names): put
`maaco_path_planing_b200/dropin` on sys.path instead of the reference directory and
`from ga_solver import GASolver` resolves to the B200 implementation."""
import _bootstrap  # noqa: F401
from maaco_path_planing_b200.ga_solver import GASolver  # noqa: F401,E402
