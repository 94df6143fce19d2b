"""Plugin-level drop-in: a piquasso connector whose permanent entries run on
the GPU (``INTEGRATION.md`` section 4).

piquasso is not a dependency of this package; the connector class is built on
demand from whatever ``piquasso`` is importable in the user's environment:

    import piquasso as pq
    from piquasso_b200.connector import make_connector
    simulator = pq.PassiveSimulator(d=100, connector=make_connector())

Only ``permanent`` and ``permanent_laplace`` are overridden
(``piquasso/_simulators/connectors/numpy_/connector.py:53-54`` binds them to
the native module this package replaces); every other operation stays the
reference's ``NumpyConnector``.
"""

from __future__ import annotations

from ._math.permanent import permanent, permanent_laplace


def make_connector():
    """Return an instance of ``B200Connector(NumpyConnector)``.

    Raises ImportError (with the reason) when piquasso is not importable."""
    try:
        from piquasso._simulators.connectors import NumpyConnector
    except ImportError as exc:  # pragma: no cover - depends on the user's env
        raise ImportError(
            "piquasso is not importable here; piquasso_b200.connector needs the "
            "reference package only to subclass its NumpyConnector") from exc

    class B200Connector(NumpyConnector):
        """NumpyConnector with the permanent hot path on the B200."""

        def permanent(self, matrix, rows, cols):
            return permanent(matrix, rows, cols)

        def permanent_laplace(self, matrix, rows, cols):
            return permanent_laplace(matrix, rows, cols)

    return B200Connector()
