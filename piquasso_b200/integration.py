"""Sampler-level drop-in for an installed piquasso.

Swapping the pybind module or the connector (INTEGRATION.md sections 2-4) puts
every ``permanent_laplace`` call on the GPU, but the reference's sampler still
issues them one photon of one shot at a time
(``piquasso/_simulators/passive/sampling.py:149-236``), which leaves the device
waiting on launch latency.  ``install()`` replaces the two functions
``particle_number_measurement`` calls
(``piquasso/_simulators/passive/simulation_steps.py:342-367``:
``generate_samples`` and ``generate_lossy_samples``) by adapters with the SAME
signatures that run all shots in lock step on the batched GPU path; per-shot
seeds and draw order are the reference's, so ``Result.samples`` is unchanged.

    import piquasso as pq
    from piquasso_b200 import integration
    handle = integration.install()      # ... simulate ...
    handle.uninstall()
"""

from __future__ import annotations

import importlib

from . import sampling

__all__ = ["install", "make_adapters"]


def _never_rejects(reject_condition):
    """True for the reference's ``lambda: False`` (simulation_steps.py:345): a
    closure-free function without global names cannot depend on any state, so one
    call tells what it always returns."""
    if reject_condition is None:
        return True
    code = getattr(reject_condition, "__code__", None)
    if code is None or getattr(reject_condition, "__closure__", None) is not None:
        return False
    if code.co_names or code.co_argcount:
        return False
    return reject_condition() is False


def make_adapters(pmf_rows=None, devices=None):
    """(generate_samples, generate_lossy_samples) with the reference's signatures
    (``sampling.py:33-42, 110-117``).  ``calculate_permanent_laplace`` and
    ``config.use_dask`` are ignored: the permanents run batched in libpqperm."""

    def generate_samples(input, shots, calculate_permanent_laplace, interferometer,
                         reject_condition, postselect_data, uniform_particle_overlap, config):
        if _never_rejects(reject_condition):
            reject_condition = None
        return sampling.generate_samples(
            input, shots, interferometer, config.seed_sequence,
            reject_condition=reject_condition, postselect_data=postselect_data,
            uniform_particle_overlap=uniform_particle_overlap, pmf_rows=pmf_rows,
            devices=devices)

    def generate_lossy_samples(input, shots, calculate_permanent_laplace, interferometer,
                               postselect_data, config):
        return sampling.generate_lossy_samples(
            input, shots, interferometer, config.seed_sequence,
            postselect_data=postselect_data, pmf_rows=pmf_rows, devices=devices)

    return generate_samples, generate_lossy_samples


class _Handle:
    def __init__(self, patched):
        self._patched = patched  # [(module, name, original)]

    def uninstall(self):
        for module, name, original in reversed(self._patched):
            setattr(module, name, original)
        self._patched = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.uninstall()
        return False


def install(modules=None, pmf_rows=None, devices=None):
    """Patch piquasso's passive sampler; returns a handle with ``uninstall()``
    (also a context manager).

    ``modules`` defaults to ``piquasso._simulators.passive.sampling`` and
    ``...simulation_steps`` (the latter imports the functions by name, so both
    bindings are replaced).  ``pmf_rows`` is passed to
    :func:`piquasso_b200.sampling.generate_samples` (CPU tests inject the oracle),
    and so is ``devices`` (CUDA device indices: the plain sampler shards its shots
    over them inside this process)."""
    if modules is None:
        try:
            modules = [importlib.import_module("piquasso._simulators.passive.sampling"),
                       importlib.import_module("piquasso._simulators.passive.simulation_steps")]
        except ImportError as exc:
            raise ImportError("piquasso is not importable here; nothing to patch") from exc
    adapters = dict(zip(("generate_samples", "generate_lossy_samples"),
                        make_adapters(pmf_rows, devices)))
    patched = []
    for module in modules:
        for name, adapter in adapters.items():
            if hasattr(module, name):
                patched.append((module, name, getattr(module, name)))
                setattr(module, name, adapter)
    if not patched:
        raise ImportError("none of the given modules exposes generate_samples")
    return _Handle(patched)
