"""``gudhi`` / ``fpsample`` stand-in modules so that the reference's own ``flooder/core.py``
runs UNMODIFIED in the authoring container.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Used only by ``tests/golden/make_golden.py`` (fixture generation) and the optional
``tests/test_oracle_vs_reference.py`` (skipped where ``/root/reference`` is absent, i.e. on
the GPU box).  Both third-party packages are pinned by the reference
(``environment.yml:38`` fpsample==0.3.3, ``:40`` gudhi==3.11.0) and neither is installable
here (no network, no wheel).  What is substituted:

* ``gudhi.DelaunayComplex(pts).create_simplex_tree()`` / ``gudhi.SimplexTree``
      -> ``oracle.simplex_tree`` (Qhull triangulation + dict container)
* ``fpsample.bucket_fps_kdline_sampling(pc, n, h=None, start_idx=None)``
      -> exact FPS (``oracle.native.fps``); ``h`` only tunes fpsample's KD-bucket depth and
         does not change the result; ``start_idx=None`` draws from ``np.random`` as the
         reference's tests assume when they seed it (``tests/test_flooder.py:32-33``).
"""
from __future__ import annotations

import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def install() -> None:
    """Register the stand-ins in ``sys.modules`` (idempotent; never shadows real packages)."""
    from . import native, simplex_tree

    if "gudhi" not in sys.modules:
        try:
            import gudhi  # noqa: F401  (a real install wins)
        except ImportError:
            g = types.ModuleType("gudhi")
            g.DelaunayComplex = simplex_tree.DelaunayComplex
            g.SimplexTree = simplex_tree.DictSimplexTree
            g.__version__ = "0+oracle-shim"
            sys.modules["gudhi"] = g
    if "fpsample" not in sys.modules:
        try:
            import fpsample  # noqa: F401
        except ImportError:
            f = types.ModuleType("fpsample")

            def bucket_fps_kdline_sampling(pc, n_samples, h=None, start_idx=None):
                pts = np.asarray(pc, dtype=np.float32)
                if start_idx is None:
                    start_idx = int(np.random.randint(pts.shape[0]))
                return native.fps(pts, n_samples, start_idx).astype(np.uint64)

            f.bucket_fps_kdline_sampling = bucket_fps_kdline_sampling
            f.__version__ = "0+oracle-shim"
            sys.modules["fpsample"] = f


def import_reference(root: str = REFERENCE_ROOT):
    """Import the reference package from ``root`` (``/root/reference`` in the authoring
    container, or the pip-installed copy under ``baseline/_ref`` on the GPU box) with the
    stand-ins in place."""
    import os

    if not os.path.isdir(os.path.join(root, "flooder")):
        raise ImportError(f"{root} holds no reference package on this machine")
    install()
    if root not in sys.path:
        sys.path.insert(0, root)
    import flooder  # the reference package

    return flooder
