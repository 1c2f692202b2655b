# PROBE-ONLY: NumPy restatement of JAX's threefry2x32 PRNG (PRNGKey, split, uniform) for seed < 2**32.
import numpy as np
_R = [[13, 15, 26, 6], [17, 29, 16, 24]]
def _rotl(x, d): return ((x << np.uint32(d)) | (x >> np.uint32(32 - d))).astype(np.uint32)
def threefry2x32(key, x0, x1):
    k0, k1 = np.uint32(key[0]), np.uint32(key[1]); k2 = np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA))
    ks = [k0, k1, k2]
    x0 = (x0 + ks[0]).astype(np.uint32); x1 = (x1 + ks[1]).astype(np.uint32)
    for r in range(5):
        for d in _R[r % 2]:
            x0 = (x0 + x1).astype(np.uint32); x1 = _rotl(x1, d); x1 = x1 ^ x0
        x0 = (x0 + ks[(r + 1) % 3]).astype(np.uint32)
        x1 = (x1 + ks[(r + 2) % 3] + np.uint32(r + 1)).astype(np.uint32)
    return x0, x1
def random_bits(key, n):
    cnt = np.arange(n, dtype=np.uint32)
    if n % 2: cnt = np.concatenate([cnt, np.zeros(1, np.uint32)])
    h = len(cnt) // 2
    a, b = threefry2x32(key, cnt[:h], cnt[h:])
    return np.concatenate([a, b])[:n]
def PRNGKey(seed): return np.array([0, np.uint32(int(seed) & 0xffffffff)], dtype=np.uint32)
def split(key, num=2): return random_bits(key, 2 * num).reshape(num, 2)
def uniform(key, shape=(), minval=0.0, maxval=1.0):
    n = int(np.prod(shape)) if shape != () else 1
    bits = random_bits(key, n)
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    out = np.maximum(np.float32(minval), f * np.float32(maxval - minval) + np.float32(minval))
    return out.reshape(shape)
