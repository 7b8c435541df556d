"""Restatement of numpy's LEGACY global RNG as used by ItemsGenerator.items_generator (BinPackingGame.py:257-285):
`np.random.seed(int)` (MT19937 init_genrand) and `np.random.randint(low, high)` (masked rejection over 32-bit draws,
no draw at all when the range has a single value).  TEST INFRASTRUCTURE ONLY: pins the device-side generator
(csrc: k_items_generate) in environments without the reference; itself pinned against numpy in
tests/test_oracle_golden.py.  numpy is a pinned dependency of the reference (not vendored): numpy/random/_mt19937.pyx
`_legacy_seeding`, numpy/random/src/distributions `buffered_bounded_masked_uint32`."""


class LegacyMT19937:
    def __init__(self, seed):
        s = seed & 0xFFFFFFFF
        self.mt = [0] * 624
        for i in range(624):
            self.mt[i] = s
            s = (1812433253 * (s ^ (s >> 30)) + i + 1) & 0xFFFFFFFF
        self.pos = 624

    def _twist(self):
        mt = self.mt
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
            mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.pos = 0

    def u32(self):
        if self.pos >= 624:
            self._twist()
        y = self.mt[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def randint(self, low, high=None):
        if high is None:
            low, high = 0, low
        rng = high - 1 - low
        if rng == 0:
            return low  # numpy returns the only value without consuming a draw
        mask = rng
        for sh in (1, 2, 4, 8, 16):
            mask |= mask >> sh
        while True:
            v = self.u32() & mask
            if v <= rng:
                return low + v


def items_generator(W, H, N, seed):
    """[w, h, a, b] lists exactly like ItemsGenerator(W, H, N).items_generator(seed)"""
    r = LegacyMT19937(seed)
    rects = [[W, H, 0, 0]]
    while len(rects) < N:
        axis = r.randint(2)
        k = r.randint(len(rects))
        w, h, a, b = rects[k]
        if axis == 0:
            if w == 1:
                continue
            cut = r.randint(a + 1, a + w)
            rects += [[cut - a, h, a, b], [w - (cut - a), h, cut, b]]
        else:
            if h == 1:
                continue
            cut = r.randint(b + 1, b + h)
            rects += [[w, cut - b, a, b], [w, h - (cut - b), a, cut]]
        del rects[k]
    return rects
