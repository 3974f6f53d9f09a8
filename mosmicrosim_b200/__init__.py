"""mosmicrosim_b200 -- B200 (sm_100a) backend for the mosmicrosim per-cell
simulation chain.  See DESIGN.md; the product is libmmsgpu.so (include/mmsgpu.h),
this package is its host-side mirror of the reference's R interface."""
from . import dates, lifehistory, synthetic  # noqa: F401

__all__ = ["backend", "pipeline", "dates", "lifehistory", "synthetic", "build"]
