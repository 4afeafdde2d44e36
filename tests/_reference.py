"""Import the real VisFly reference for oracle validation and golden generation (test-side alias of
``baseline/ref_loader.py``, which holds the loader, the RK4 repairs R1-R3 and the third-party stubs).

Only usable where the reference tree is present (``/root/reference`` in the build container, or the
``baseline/_ref`` copy made by ``__graft_entry__.build()``).  Nothing that runs on the GPU box may depend on
it: callers must check ``reference_available()`` and skip otherwise.
"""
from __future__ import annotations

import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from baseline.ref_loader import (REFERENCE_ROOT, default_dtype, load_reference, load_reference_algorithms,  # noqa: E402,F401
                                 load_reference_envs, make_reference_dynamics, reference_available,
                                 reference_on_device, reference_origin)
