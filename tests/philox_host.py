"""Host (numpy) restatement of the device random streams of the device-resident chain (st_kernels.cu: philox_normal,
philox_uniform, philox_gamma; Philox4x32-10 + Box-Muller + Marsaglia-Tsang) — test infrastructure: it lets a test REPLAY a
device-resident run through the host-driven model-layer calls with exactly the random numbers the device drew."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)
K_STREAM_U, K_STREAM_ACCEPT, K_STREAM_GAMMA, K_STREAM_BETA = 1 << 48, 2 << 48, 3 << 48, 4 << 48
K_COUNTER_YHAT = 1 << 62


def _philox(key, counter, seed):
    key = np.atleast_1d(np.asarray(key, dtype=np.uint64))
    counter = np.broadcast_to(np.asarray(counter, dtype=np.uint64), key.shape)
    c0, c1, c2, c3 = key & M32, key >> np.uint64(32), counter & M32, counter >> np.uint64(32)
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c0, np.uint64(0xCD9E8D57) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & M32, p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & M32, (k1 + np.uint64(0xBB67AE85)) & M32
    return c0, c1, c2, c3


def _u01(hi, lo):
    return ((((hi << np.uint64(32)) | lo) >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def normal(seed, key, counter):
    c0, c1, c2, c3 = _philox(key, counter, seed)
    return np.sqrt(-2.0 * np.log(_u01(c0, c1))) * np.cos(2.0 * np.pi * _u01(c2, c3))


def uniform(seed, key, counter):
    c0, c1, _, _ = _philox(key, counter, seed)
    return _u01(c0, c1)


def gamma(seed, key, counter, shape, scale):
    d = shape - 1.0 / 3.0
    c = 1.0 / np.sqrt(9.0 * d)
    att = 0
    while True:
        x = float(normal(seed, key + (att << 8), counter)[0])
        v = 1.0 + c * x
        if v > 0:
            v = v * v * v
            u = float(uniform(seed, key + (att << 8) + 128, counter)[0])
            if u < 1.0 - 0.0331 * x ** 4 or np.log(u) < 0.5 * x * x + d * (1.0 - v + np.log(v)):
                return d * v * scale
        att += 1
