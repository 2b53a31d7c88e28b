"""sea_b200 — B200-native (sm_100a) hot path for ParsaEsmati/SEA.

The package holds only what the State-Exchange-Attention hot path needs:

* ``csrc/``    hand-written CUDA kernels + the C ABI (``include/sea_b200.h``), built in-tree to
               ``sea_b200/lib/libsea_b200.so`` by ``__graft_entry__.build()``;
* ``_lib``     ctypes binding of that ABI (plain pointers/sizes; torch only supplies device
               memory and streams);
* ``ops``      one thin Python wrapper per C entry point (used by the parity tests);
* ``temporal`` / ``spatial``  host-side mirrors of the reference ``TemporalModel`` /
               ``SpatialModel`` interface (same names, arguments, state_dict);
* ``accelerate`` rebinding of ``forward`` on unchanged reference module instances;
* ``install()`` the one-line hook that wraps the reference's ``get_model`` / ``initialize_spatial_model``.

There is no CPU fallback: importing is cheap, but every op raises if the CUDA library is missing.
"""
from ._lib import lib, LibraryMissing, check  # noqa: F401
from .hooks import install, uninstall  # noqa: F401

__all__ = ["lib", "LibraryMissing", "check", "install", "uninstall"]
__version__ = "0.1.0"
