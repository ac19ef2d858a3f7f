"""Pure-Python restatement of the numpy Generator(PCG64) draws used on the hot path (oracle, test-only).

Third-party arithmetic: numpy (reference pins 1.22.4, requirements.txt:18; this image has 2.3.5 which is the
oracle of record).  Algorithms restated from numpy's published PCG64 / bounded-integer / shuffle routines
(SURVEY.md Appendix B); call sites in the reference:
  envs/car_flag.py:147,158   integers(0, 2, size=1), uniform(-0.2, 0.2)
  envs/memory_cards.py:73,77,110-113   shuffle(state[10]), integers(10)
  utils/context.py:50        integers(A, size=(ctx, 1))
  dtqn/agents/dtqn.py:78-79  random(), integers(A)
The class mirrors numpy's ``bit_generator.state`` dict exactly so device RNG state can be compared bit-for-bit.
"""
import numpy as np

MASK64 = (1 << 64) - 1
MASK128 = (1 << 128) - 1
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645


class PCG64:
    __slots__ = ("state", "inc", "has_uint32", "uinteger")

    def __init__(self, state: int, inc: int, has_uint32: int = 0, uinteger: int = 0):
        self.state, self.inc, self.has_uint32, self.uinteger = state, inc, has_uint32, uinteger

    @classmethod
    def from_seed(cls, seed: int) -> "PCG64":
        """Seeding is done with numpy itself (SeedSequence hashing is not restated): car_flag.py:73-74."""
        st = np.random.PCG64(np.random.SeedSequence(seed)).state
        return cls(st["state"]["state"], st["state"]["inc"], st["has_uint32"], st["uinteger"])

    @classmethod
    def from_numpy(cls, gen) -> "PCG64":
        st = gen.bit_generator.state
        return cls(st["state"]["state"], st["state"]["inc"], st["has_uint32"], st["uinteger"])

    def as_tuple(self):
        return (self.state >> 64, self.state & MASK64, self.inc >> 64, self.inc & MASK64, self.has_uint32, self.uinteger)

    def next64(self) -> int:
        self.state = (self.state * PCG_MULT + self.inc) & MASK128
        hi, lo = self.state >> 64, self.state & MASK64
        x = hi ^ lo
        r = self.state >> 122
        return ((x >> r) | (x << ((64 - r) & 63))) & MASK64

    def next32(self) -> int:
        if self.has_uint32:
            self.has_uint32 = 0
            return self.uinteger
        v = self.next64()
        self.has_uint32 = 1
        self.uinteger = v >> 32
        return v & 0xFFFFFFFF

    def next_double(self) -> float:
        return (self.next64() >> 11) * (1.0 / 9007199254740992.0)

    # Generator.integers(low, high) for int64 dtype with range <= 2**32-1 (Lemire, 32-bit buffered path)
    def integers(self, low: int, high: int = None) -> int:
        if high is None:
            low, high = 0, low
        rng = high - low - 1
        if rng == 0:
            return low
        ex = rng + 1
        m = self.next32() * ex
        left = m & 0xFFFFFFFF
        if left < ex:
            thresh = (0xFFFFFFFF - rng) % ex
            while left < thresh:
                m = self.next32() * ex
                left = m & 0xFFFFFFFF
        return low + (m >> 32)

    def uniform(self, a: float, b: float) -> float:
        return a + (b - a) * self.next_double()

    def random(self) -> float:
        return self.next_double()

    def interval(self, mx: int) -> int:
        """random_interval: masked rejection on 32-bit draws (used by Generator.shuffle)."""
        if mx == 0:
            return 0
        mask = mx
        for s in (1, 2, 4, 8, 16, 32):
            mask |= mask >> s
        while True:
            v = self.next32() & mask
            if v <= mx:
                return v

    def shuffle(self, x) -> None:
        for i in range(len(x) - 1, 0, -1):
            j = self.interval(i)
            x[i], x[j] = x[j], x[i]
