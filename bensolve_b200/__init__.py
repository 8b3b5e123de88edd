"""bensolve_b200 -- B200-native polyhedral cut engine behind bensolve's ``poly__*`` C API.

The product is the shared library ``libbslv_poly_b200.so`` built from ``csrc/`` (hand-written sm_100a
kernels + the reference-facing C ABI, see ``include/bensolve_b200.h`` and ``INTEGRATION.md``).  This
package is the thin Python side used by tests and ``bench.py``:

* :mod:`bensolve_b200.capi`      -- ctypes mirror of the ABI, ``PolyEngine`` driver, canonical state comparison
* :mod:`bensolve_b200.polytopes` -- synthetic halfspace traces (BASELINE configs 3-5) and replay helpers
* :mod:`bensolve_b200.vlpgen`    -- synthetic ``.vlp`` problems for the closed Benson loop
* :mod:`bensolve_b200.dist`      -- multi-GPU communicator set-up (torch.distributed rendezvous, NCCL exchange)
* :mod:`bensolve_b200.build`     -- nvcc / gcc build recipes

There is no CPU fallback for the cut: :func:`bensolve_b200.capi.load_product` raises when the CUDA
library has not been built, and the library aborts with a message when no CUDA device is visible.
"""
from .capi import PolyEngine, compare_states, load_lib, load_product  # noqa: F401

__all__ = ["PolyEngine", "compare_states", "load_lib", "load_product"]
