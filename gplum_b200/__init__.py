"""gplum_b200 -- B200-native soft-force path for GPLUM (P3T on FDPS).

Only the hot path lives here: the sm_100a kernels + C ABI (csrc/, libgplum_b200.so) and the
host-side mirror of the reference's functor / multi-walk interface.
"""
from . import structs  # noqa: F401
