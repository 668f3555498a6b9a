"""Host-side view of the RNG contract: Philox4x32-10 streams keyed (seed, class, iteration,
individual) with an in-stream cursor; draw d of a stream is word pair (d & 1) of block d >> 1.
Used only where the reference's control flow consumes a *variable* number of draws on the host
(population initialisers); the per-iteration kernels regenerate the same streams on the device.
"""
from __future__ import annotations

import numpy as np

CLS_MAACO_TOUR, CLS_PSO_INIT, CLS_PSO_PAD, CLS_PSO_UPDATE = 1, 2, 3, 4
CLS_GA_INIT, CLS_GA_PAD, CLS_GA_SELECT, CLS_GA_BREED, CLS_MPA_PHASE, CLS_MPA_FADS = 5, 6, 7, 8, 9, 10

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint64 arrays holding 32-bit values."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> _S32) ^ c1 ^ np.uint64(k0)) & _MASK, p1 & _MASK, \
                         ((p0 >> _S32) ^ c3 ^ np.uint64(k1)) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _u53(a, b):
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0


def stream_block(seed, cls, it, ind, n_draws):
    """First n_draws uniforms of each stream (seed, cls, it, ind[i]) -> array [len(ind), n_draws]."""
    ind = np.atleast_1d(np.asarray(ind, dtype=np.uint64))
    nb = (n_draws + 1) // 2
    blk = np.arange(nb, dtype=np.uint64)[None, :]
    shape = (ind.size, nb)
    c0 = np.broadcast_to(blk, shape)
    c1 = np.broadcast_to(ind[:, None], shape)
    x, y, z, w = philox4x32_10(c0, c1, np.full(shape, it, np.uint64), np.full(shape, cls, np.uint64),
                               seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((ind.size, nb * 2))
    out[:, 0::2] = _u53(x, y)
    out[:, 1::2] = _u53(z, w)
    return out[:, :n_draws]


class Stream:
    """Sequential cursor over one stream (host control flow with a variable number of draws)."""

    def __init__(self, seed, cls, it, ind, prefetch=32):
        self.key = (seed, cls, it, ind)
        self.buf = stream_block(seed, cls, it, [ind], prefetch)[0]
        self.cursor = 0

    def random(self):
        if self.cursor >= self.buf.size:
            self.buf = stream_block(*self.key[:3], [self.key[3]], self.buf.size * 2)[0]
        u = float(self.buf[self.cursor])
        self.cursor += 1
        return u

    def below(self, n):
        j = int(self.random() * n)
        return j if j < n else n - 1

    def uniform(self, a, b):
        return a + (b - a) * self.random()
