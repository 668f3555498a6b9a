"""Drop-in for the reference module `dijkstra` (same module and class names)."""
import _bootstrap  # noqa: F401
from maaco_path_planing_b200.dijkstra import DijkstraSolver  # noqa: F401,E402
