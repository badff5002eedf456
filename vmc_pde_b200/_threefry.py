"""Host-side threefry2x32 key arithmetic in JAX 0.2.18's layout (PRNGKey / split / small draws used at init).

Product code (key chains are host logic in the reference too: sampler.py:57-60,73; var_state.py:111-116).
The bulk draws happen on the device (csrc/rng.cuh); this module only derives keys and the few
initialisation-time draws.  Independent of oracle/ by design.
"""
import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _block(k0, k1, x0, x1):
    with np.errstate(over="ignore"):
        k0, k1 = np.uint32(k0), np.uint32(k1)
        ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
        x0 = (np.asarray(x0, np.uint32) + ks[0]).astype(np.uint32)
        x1 = (np.asarray(x1, np.uint32) + ks[1]).astype(np.uint32)
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = ((x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))).astype(np.uint32)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(i + 2) % 3] + np.uint32(i + 1)).astype(np.uint32)
    return x0, x1


def PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def _bits(key, count):
    c = np.asarray(count, np.uint32).ravel()
    odd = c.size % 2
    if odd:
        c = np.concatenate([c, np.zeros(1, np.uint32)])
    h = c.size // 2
    a, b = _block(key[0], key[1], c[:h], c[h:])
    out = np.concatenate([a, b])
    return out[:-1] if odd else out


def split(key, num=2):
    return _bits(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)


def random_bits32(key, size):
    return _bits(key, np.arange(size, dtype=np.uint32))


def uniform01(key, size):
    """float64 U[0,1) in JAX's layout (64 random bits per element from block (i, size + i))."""
    bits = _bits(key, np.arange(2 * size, dtype=np.uint32))
    v = (bits[:size].astype(np.uint64) << np.uint64(32)) | bits[size:].astype(np.uint64)
    return ((v >> np.uint64(12)) | np.float64(1.0).view(np.uint64)).view(np.float64) - 1.0


def uniform01_f32(key, size):
    """float32 U[0,1) in JAX's layout (one 32-bit word per element): what jax.nn.initializers.uniform draws."""
    bits = random_bits32(key, size)
    return ((bits >> np.uint32(9)) | np.float32(1.0).view(np.uint32)).view(np.float32) - np.float32(1.0)


def fold_in(key, data):
    """jax.random.fold_in(key, data) = threefry_2x32(key, PRNGKey(data)) for a 32-bit `data`."""
    a, b = _block(key[0], key[1], np.zeros(1, np.uint32), np.array([int(data) & 0xFFFFFFFF], np.uint32))
    return np.array([a[0], b[0]], dtype=np.uint32)


def fold_in_str(key, name):
    """flax.core.scope._fold_in_str (flax 0.3.x): fold in the first 4 bytes (big endian) of the SHA-1 of `name`."""
    import hashlib
    return fold_in(key, int.from_bytes(hashlib.sha1(name.encode("utf-8")).digest()[:4], byteorder="big"))


def permutation(key, n):
    """jax.random.permutation(key, n): rounds of sort-by-random-32-bit-keys (jax._src.random._shuffle)."""
    x = np.arange(n)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    for _ in range(rounds):
        key, sub = split(key)
        x = x[np.argsort(random_bits32(sub, n), kind="stable")]
    return x


def choice_no_replace(key, n, k):
    return permutation(key, n)[:k]
