"""flooder_b200 -- B200-native (sm_100a) Flood-complex hot path behind the reference's API.

    from flooder_b200 import flood_complex, generate_landmarks

mirrors ``from flooder import flood_complex, generate_landmarks`` of plus-rkwitt/flooder for the
path  landmark FPS -> (host Delaunay) -> per-simplex covering radius -> filtered complex.
"""
from .core import (flood_complex, generate_landmarks, generate_grid, generate_uniform_weights,
                   fps_indices, covering_values, PreparedCloud)
from .simplex_tree import SimplexTree
from .synthetic import (generate_swiss_cheese_points, generate_annulus_points_2d,
                        generate_noisy_torus_points_3d, generate_figure_eight_points_2d)

__version__ = "0.1.0"

__all__ = [
    "flood_complex", "generate_landmarks", "generate_grid", "generate_uniform_weights",
    "fps_indices", "covering_values", "PreparedCloud", "SimplexTree",
    "generate_swiss_cheese_points", "generate_annulus_points_2d",
    "generate_noisy_torus_points_3d", "generate_figure_eight_points_2d",
]
