"""piquasso_b200 -- B200-native (sm_100a) implementation of piquasso's
permanent hot path, behind the reference's own entry points.

``piquasso_b200._math.permanent`` mirrors ``piquasso._math.permanent``
(``permanent`` and ``permanent_laplace``); the arithmetic runs in
``libpqperm.so`` (C ABI: ``include/pqperm.h``).  There is no CPU fallback.
"""

__version__ = "0.1.0"
