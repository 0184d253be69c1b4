"""pyccl stand-in for running the UNMODIFIED reference's light-cone path in this container (pyccl ~=3.2 is absent).

The two names the reference uses (`ccl.Cosmology(...)`, `ccl.comoving_radial_distance(cosmo, a)`, measure_w_lightcone.py:123-130)
delegate to measure_ia_b200/cosmo.py, so the reference and the product see the SAME distances: what the fixtures pin is
everything downstream of the distance conversion, not CCL's integral."""
from measure_ia_b200.cosmo import Cosmology, comoving_radial_distance  # noqa: F401
