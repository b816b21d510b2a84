"""gpupfem2_b200 -- B200-native PFEM-2 particle step behind gpuPfem2's ParticleHandler2D interface.

The product is the CUDA library (csrc/ -> libpfem2_b200.so, C ABI in include/pfem2_b200.h).  This
package is the thin host-side mirror of the reference interface used by the tests and bench.py.
"""
from .mesh import HostMesh, load_dat, poiseuille_field, structured_channel, vortex_field  # noqa: F401

__all__ = ["HostMesh", "load_dat", "structured_channel", "poiseuille_field", "vortex_field"]
