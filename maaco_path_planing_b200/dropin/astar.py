"""Drop-in for the reference module `astar` (same module and class names): put
`maaco_path_planing_b200/dropin` on sys.path instead of the reference directory and
`from astar import AStarSolver` resolves to the B200 implementation."""
import _bootstrap  # noqa: F401
from maaco_path_planing_b200.astar import AStarSolver  # noqa: F401,E402
