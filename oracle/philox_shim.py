"""TEST INFRASTRUCTURE -- not part of the product path.

A counter-based stand-in for CPython's ``random`` module, used to drive the
*unmodified* reference ``Environment`` (``shipping/environment.py``) on the same
word stream the CUDA kernels consume.  The reference draws from the module
global ``random`` (``shipping/environment.py:3``; call sites ``:63,64,104,157,
160,163,167,177,195,320``), so rebinding ``shipping.environment.random`` to an
instance of :class:`PhiloxRandom` is all that is needed -- no reference file is
edited.

Word-stream contract (normative; mirrored by ``shippingenv_b200/csrc`` and by
``oracle/shipenv_oracle.c``):

* ``word(i) = Philox4x32-10(counter=(i >> 2, c1, c2, c3), key=(k0, k1))[i & 3]``
  where ``(c1, c2, c3, k0, k1)`` is the *cell* selected with :meth:`seek`.
  - step path:   c1 = env clock (u32), c2/c3 = global env id lo/hi,
                 key = (seed_lo, seed_hi)
  - rollouts:    c1 = simulation index, c2/c3 = root id lo/hi,
                 key = (seed_lo, seed_hi ^ 0x524F4C4C)      ("ROLL")
  - port stock:  c1 = port index, c2 = c3 = 0xFFFFFFFF,
                 key = (seed_lo, seed_hi ^ 0x504F5254)      ("PORT")
* ``random()``       = word * 2**-32                    (fp64, exact)
* ``uniform(a, b)``  = a + (b - a) * random()            (two rounded fp64 ops,
                       the same expression as CPython ``random.py:uniform``)
* ``_randbelow(n)``  = (word * n) >> 32                  (one word)
* ``randint(a, b)``  = a + _randbelow(b - a + 1)
* ``choice(seq)``    = seq[_randbelow(len(seq))]
* ``betavariate(2, 2)`` = median of three ``random()`` draws (the 2nd order
  statistic of three uniforms is Beta(2,2)); other shapes are not needed by the
  reference (``shipping/environment.py:195``) and raise.
"""

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
ROLLOUT_KEY_TWEAK = 0x524F4C4C
PORT_KEY_TWEAK = 0x504F5254


def philox4x32_10(ctr, key):
    """Philox4x32 with 10 rounds (Salmon et al., Random123).  ctr: 4 u32, key: 2 u32."""
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = (
            ((p1 >> 32) ^ c1 ^ k0) & MASK,
            p1 & MASK,
            ((p0 >> 32) ^ c3 ^ k1) & MASK,
            p0 & MASK,
        )
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


class PhiloxRandom:
    """Drop-in for the subset of ``random`` the reference environment calls."""

    def __init__(self, seed=0):
        self.seed = seed & 0xFFFFFFFFFFFFFFFF
        self.seek(0, 0, 0)
        self.trace = None  # set to a list to record (fn, value) per draw

    # -- cell selection ---------------------------------------------------
    def seek(self, c1, c2, c3, key_tweak=0, word=0):
        self._c = (c1 & MASK, c2 & MASK, c3 & MASK)
        self._key = (self.seed & MASK, ((self.seed >> 32) ^ key_tweak) & MASK)
        self.i = word
        self._blk = None
        self._blk_idx = -1

    def seek_env(self, env_id, clock, word=0):
        self.seek(clock, env_id & MASK, (env_id >> 32) & MASK, 0, word)

    def seek_rollout(self, root_id, sim):
        self.seek(sim, root_id & MASK, (root_id >> 32) & MASK, ROLLOUT_KEY_TWEAK)

    def seek_port(self, port_idx):
        self.seek(port_idx, MASK, MASK, PORT_KEY_TWEAK)

    # -- raw words --------------------------------------------------------
    def next_word(self):
        b = self.i >> 2
        if b != self._blk_idx:
            self._blk = philox4x32_10((b & MASK,) + self._c, self._key)
            self._blk_idx = b
        w = self._blk[self.i & 3]
        self.i += 1
        return w

    # -- the ``random`` API subset ---------------------------------------
    def random(self):
        v = self.next_word() * 2.0 ** -32
        if self.trace is not None:
            self.trace.append(("random", v))
        return v

    def uniform(self, a, b):
        v = a + (b - a) * (self.next_word() * 2.0 ** -32)
        if self.trace is not None:
            self.trace.append(("uniform", v))
        return v

    def _randbelow(self, n):
        return (self.next_word() * n) >> 32

    def randint(self, a, b):
        v = a + self._randbelow(b - a + 1)
        if self.trace is not None:
            self.trace.append(("randint", v))
        return v

    def choice(self, seq):
        v = seq[self._randbelow(len(seq))]
        if self.trace is not None:
            self.trace.append(("choice", v))
        return v

    def betavariate(self, alpha, beta):
        if alpha != 2 or beta != 2:
            raise NotImplementedError("shim implements betavariate(2, 2) only")
        a = self.next_word()
        b = self.next_word()
        c = self.next_word()
        m = max(min(a, b), min(max(a, b), c))
        v = m * 2.0 ** -32
        if self.trace is not None:
            self.trace.append(("betavariate", v))
        return v

    def seed_(self, *a, **k):  # pragma: no cover - parity runs never reseed
        raise RuntimeError("PhiloxRandom is positioned with seek(), not seed()")
