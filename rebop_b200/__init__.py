"""rebop_b200 -- B200-native ensemble engine for rebop's Gillespie direct method."""
from rebop_b200 import _ffi  # noqa: F401  (fails loudly when the CUDA library is not built)

from rebop_b200.gillespie import Gillespie  # noqa: E402
from rebop_b200.system import define_system  # noqa: E402

__version__ = _ffi.lib.rebop_b200_version().decode()
__all__ = ("Gillespie", "define_system", "__version__")
