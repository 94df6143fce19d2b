"""Host-side mirror of ``piquasso._math`` for the permanent path."""
