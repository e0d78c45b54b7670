"""Keyed (counter-based) random word source -- TEST INFRASTRUCTURE (oracle/).

The reference draws every random decision from CPython's global MT19937 stream
(sugarscape.py:160-162, agent.py:369,405,1033,1108, sugarscape.py:517).  That
stream is inherently sequential, so the throughput mode of the
engine uses a *keyed* word source instead: the j-th 32-bit word consumed by
agent `id` during timestep `t` is a pure function of (seed, t, id, j), and the
per-step execution order is the ascending order of a keyed bijection of the
agent slot (id-1).  The CPython algorithms layered on top of the words
(`_randbelow` rejection sampling, Fisher-Yates `shuffle`, `choice`;
/usr/lib/python3.12/random.py:242,350,341) are kept verbatim.

This file is the plain-Python statement of that word source.  It is used by
  * oracle/ref_harness.py -- a shim object with the `random` module's API that
    is bound to `agent.random` / `sugarscape.random` so that the UNMODIFIED
    reference consumes keyed words, and
  * tests/ -- to cross-check the C oracle and the CUDA implementation.
The same constants appear in oracle/ss_oracle.c and csrc/ss_rng.cuh.
"""

M64 = (1 << 64) - 1
M32 = (1 << 32) - 1
GOLDEN = 0x9E3779B97F4A7C15
C_AGENT = 0xD1342543DE82EF95
C_ROUND = 0xA24BAED4963EE407


def mix64(z):
    """splitmix64 finaliser."""
    z &= M64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return z


def fmix32(h):
    """murmur3 32-bit finaliser."""
    h &= M32
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M32
    h ^= h >> 16
    return h


def step_key(seed, t):
    return mix64((seed + GOLDEN * (t + 1)) & M64)


def agent_key(skey, agent_id):
    return mix64(skey ^ ((agent_id * C_AGENT) & M64))


def word(akey, j):
    """j-th 32-bit word of one agent's per-step stream."""
    return mix64((akey + j * GOLDEN) & M64) >> 32


def order_half_bits(n_slots):
    """Half-width of the balanced Feistel domain covering [0, n_slots)."""
    kb = max(1, (max(n_slots, 1) - 1).bit_length())
    return max(1, (kb + 1) // 2)


def round_keys(skey):
    return [mix64((skey + (r + 1) * C_ROUND) & M64) >> 32 for r in range(4)]


def priority(rk, half, slot):
    """Keyed bijection on [0, 2^(2*half)): execution priority of agent slot."""
    mask = (1 << half) - 1
    L, R = (slot >> half) & mask, slot & mask
    for r in range(4):
        L, R = R, L ^ (fmix32(R ^ rk[r]) & mask)
    return (L << half) | R


def priority_inverse(rk, half, prio):
    mask = (1 << half) - 1
    L, R = (prio >> half) & mask, prio & mask
    for r in (3, 2, 1, 0):
        L, R = R ^ (fmix32(L ^ rk[r]) & mask), L
    return (L << half) | R


class KeyedStream:
    """CPython `random.Random` algorithms over one agent's keyed word stream."""

    def __init__(self, akey):
        self.akey = akey
        self.j = 0

    def getrandbits(self, k):
        assert 0 < k <= 32
        w = word(self.akey, self.j)
        self.j += 1
        return w >> (32 - k)

    def randbelow(self, n):
        k = n.bit_length()
        r = self.getrandbits(k)
        while r >= n:
            r = self.getrandbits(k)
        return r

    def shuffle(self, x):
        for i in reversed(range(1, len(x))):
            j = self.randbelow(i + 1)
            x[i], x[j] = x[j], x[i]

    def choice(self, seq):
        return seq[self.randbelow(len(seq))]
