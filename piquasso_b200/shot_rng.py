"""Per-shot numpy generators, drawn from in bulk.

The reference sampler gives shot ``idx`` its own
``np.random.default_rng(seed_sequence + idx)`` and draws from it twice per
photon (``piquasso/_simulators/passive/sampling.py:197-205, 752-753``):
``rng.choice(len(to_shrink))`` and ``rng.choice(arange(d), p=pmf)``.  Calling
10^4 generators 50 times each from Python costs more than the GPU work of the
small photon steps, so the lock-step sampler pulls every shot's RAW 64-bit
stream once (``PCG64.random_raw``) and replays numpy's own derivations on all
shots at once:

* ``Generator.random()``: ``(next_uint64 >> 11) * 2**-53``;
* ``Generator.integers(0, high)`` (what ``choice(n)`` calls), ``high <= 2**32``:
  Lemire's bounded rejection on ``next_uint32``
  (``buffered_bounded_lemire_uint32`` in numpy/random/src/distributions), and
  nothing at all is drawn when ``high == 1``;
* ``next_uint32`` of PCG64 hands out the low half of a fresh 64-bit output and
  keeps the high half for the next 32-bit request; 64-bit requests never look
  at that half.

``tests/test_host.py::test_shot_streams_replay_numpy_generators`` pins all of
this against real ``Generator`` objects, rejection path included.
"""

from __future__ import annotations

import numpy as np

_MASK32 = np.uint64(0xFFFFFFFF)


class ShotStreams:
    """``np.random.default_rng(seed_sequence + idx)`` for ``idx`` in
    ``[begin, end)``, addressed by shot number ``0 .. end-begin-1``."""

    def __init__(self, seed_sequence, begin, end, draws_per_shot):
        n = max(0, end - begin)
        self._chunk = max(4, int(draws_per_shot))
        self._seed0 = seed_sequence + begin
        self._bitgens = None
        self._raw = self._bulk(n, self._chunk)
        if self._raw is None:
            # seeds the library does not restate (negative, >= 2**64, not an int): real
            # numpy bit generators, one per shot
            self._bitgens = [np.random.PCG64(seed_sequence + idx) for idx in range(begin, end)]
            self._raw = np.empty((n, self._chunk), dtype=np.uint64)
            for i, bg in enumerate(self._bitgens):
                self._raw[i] = bg.random_raw(self._chunk)
        self._pos = np.zeros(n, dtype=np.int64)       # next unread raw output
        self._has32 = np.zeros(n, dtype=bool)         # PCG64's pending high half
        self._buf32 = np.zeros(n, dtype=np.uint64)

    def _bulk(self, n, draws):
        """All shots' first `draws` raw outputs from ONE library call
        (``pq_pcg64_streams``: numpy's SeedSequence + PCG64 restated on the host,
        threaded), or None when the seeds are outside what it restates."""
        import ctypes

        from . import _lib
        seed0 = self._seed0
        if not isinstance(seed0, (int, np.integer)) or seed0 < 0 or seed0 + n >= (1 << 64):
            return None
        out = np.empty((n, draws), dtype=np.uint64)
        rc = _lib.load().pq_pcg64_streams(
            ctypes.c_uint64(int(seed0)), n, draws,
            out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
        return out if rc == 0 else None

    # -- raw outputs ---------------------------------------------------------
    def _ensure(self, shots, count=1):
        """Make sure `count` more raw outputs exist for the given shots."""
        while (self._pos[shots] + count > self._raw.shape[1]).any():
            have = self._raw.shape[1]
            if self._bitgens is None:
                # the streams are pure functions of the seeds: regenerate, longer
                self._raw = self._bulk(self._raw.shape[0], have + self._chunk)
                continue
            more = np.empty((self._raw.shape[0], self._chunk), dtype=np.uint64)
            for i, bg in enumerate(self._bitgens):
                more[i] = bg.random_raw(self._chunk)
            self._raw = np.concatenate([self._raw, more], axis=1)

    def _next64(self, shots):
        self._ensure(shots)
        out = self._raw[shots, self._pos[shots]]
        self._pos[shots] += 1
        return out

    def _next32(self, shots):
        have = self._has32[shots]
        out = self._buf32[shots].copy()
        fresh = shots[~have]
        if fresh.size:
            raw = self._next64(fresh)
            out[~have] = raw & _MASK32
            self._buf32[fresh] = raw >> np.uint64(32)
        self._has32[shots] = ~have
        return out

    # -- numpy's derived draws -------------------------------------------------
    def random(self, shots):
        """``rng.random()`` of every listed shot."""
        shots = np.asarray(shots, dtype=np.int64)
        return (self._next64(shots) >> np.uint64(11)) * (1.0 / 9007199254740992.0)

    def integers(self, shots, high):
        """``rng.integers(0, high[i])`` (= ``rng.choice(high[i])``) of every listed
        shot; ``1 <= high <= 2**32``."""
        shots = np.asarray(shots, dtype=np.int64)
        high = np.broadcast_to(np.asarray(high, dtype=np.int64), shots.shape)
        if ((high < 1) | (high > (1 << 32))).any():
            raise ValueError("integers(): need 1 <= high <= 2**32")
        out = np.zeros(shots.shape, dtype=np.int64)
        draw = high > 1                       # a range of one value consumes nothing
        full = high == (1 << 32)              # the whole 32-bit range: no rejection
        if full.any():
            out[full] = self._next32(shots[full]).astype(np.int64)
            draw = draw & ~full
        if not draw.any():
            return out
        who = shots[draw]
        excl = high[draw].astype(np.uint64)   # rng_excl = rng + 1 = high
        m = self._next32(who) * excl
        leftover = m & _MASK32
        suspect = leftover < excl
        if suspect.any():
            threshold = (_MASK32 - (excl - np.uint64(1))) % excl
            for k in np.flatnonzero(suspect & (leftover < threshold)):
                one = who[k: k + 1]
                while (m[k] & _MASK32) < threshold[k]:
                    m[k] = self._next32(one)[0] * excl[k]
        out[draw] = (m >> np.uint64(32)).astype(np.int64)
        return out
