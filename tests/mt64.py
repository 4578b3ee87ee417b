"""MT19937-64 (Matsumoto & Nishimura) in Python integers — an independent restatement of
std::mt19937_64 used only to pin the oracle's bootstrap stream in the CPU tests."""
M64 = (1 << 64) - 1


class MT19937_64:
    NN, MM = 312, 156
    MATRIX_A, UM, LM = 0xB5026F5AA96619E9, 0xFFFFFFFF80000000, 0x7FFFFFFF

    def __init__(self, seed: int):
        self.mt = [0] * self.NN
        self.mt[0] = seed & M64
        for i in range(1, self.NN):
            self.mt[i] = (6364136223846793005 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 62)) + i) & M64
        self.mti = self.NN

    def next(self) -> int:
        if self.mti >= self.NN:
            mt, NN, MM = self.mt, self.NN, self.MM
            for i in range(NN):
                x = (mt[i] & self.UM) | (mt[(i + 1) % NN] & self.LM)
                mt[i] = mt[(i + MM) % NN] ^ (x >> 1) ^ (self.MATRIX_A if x & 1 else 0)
            self.mti = 0
        x = self.mt[self.mti]
        self.mti += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        return x & M64
