"""Restatement of the JAX 0.2.18 threefry2x32 PRNG layout used by the reference.

The arithmetic lives in a third-party dependency that is not vendored under /root/reference:
jax 0.2.18 / jaxlib 0.1.74 (reference README.md:13-20).  Call sites on the hot path:
sampler.py:26,33,58-60,73 (PRNGKey / split / multivariate_normal), var_state.py:111,115-116
(PRNGKey / split / choice), tdvp.py:154-155 (normal / uniform).  This file restates the
published algorithm (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11,
Threefry-2x32 with 20 rounds) and JAX's counter layout (jax/_src/random.py of that release:
PRNGKey, _threefry_split, threefry_2x32, _random_bits, _uniform, _normal_real, _shuffle).

Pins: Random123 known-answer vectors, jax.random.split(PRNGKey(0)), jax.random.uniform(PRNGKey(0)) == 0.41845703
(tests/test_oracle.py), and -- reference-held -- the stored outputs of the reference's own JAX runs
(tests/golden/ref_*.npz, tests/test_reference_pins.py): `normal(PRNGKey(0), (10^4, 6))` reproduces the t=0 record of
paper_plot/data_phaseSpace/Wiener/*/infos.hdf5 (mean, covariance, ball counts) to 1e-14, and split / per-particle
split / normal reproduce all 1201 records of the stored Wiener trajectory to round-off.
"""
import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = np.uint64(0xFFFFFFFF)


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def threefry2x32(k0, k1, x0, x1):
    """One Threefry-2x32-20 block per element.  k0,k1: uint32 scalars; x0,x1: uint32 arrays."""
    with np.errstate(over="ignore"):
        k0 = np.uint32(k0)
        k1 = np.uint32(k1)
        ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
        x0 = (np.asarray(x0, dtype=np.uint32) + ks[0]).astype(np.uint32)
        x1 = (np.asarray(x1, dtype=np.uint32) + ks[1]).astype(np.uint32)
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(i + 2) % 3] + np.uint32(i + 1)).astype(np.uint32)
    return x0, x1


def prng_key(seed):
    """jax.random.PRNGKey(int): [high 32 bits, low 32 bits]."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def threefry_2x32(key, count):
    """jax._src.random.threefry_2x32: counts are split into two halves that form the block pairs."""
    count = np.asarray(count, dtype=np.uint32).ravel()
    odd = count.size % 2
    if odd:
        count = np.concatenate([count, np.zeros(1, np.uint32)])
    half = count.size // 2
    o0, o1 = threefry2x32(key[0], key[1], count[:half], count[half:])
    out = np.concatenate([o0, o1])
    return out[:-1] if odd else out


def split(key, num=2):
    """jax.random.split (sampler.py:58-60,73; var_state.py:115)."""
    return threefry_2x32(key, np.arange(num * 2, dtype=np.uint32)).reshape(num, 2)


def random_bits(key, bit_width, size):
    """jax._src.random._random_bits for a flat array of `size` elements (32 or 64 bit)."""
    max_count = int(np.ceil(bit_width * size / 32))
    bits = threefry_2x32(key, np.arange(max_count, dtype=np.uint32))
    if bit_width == 64:
        hi, lo = bits[:size].astype(np.uint64), bits[size:].astype(np.uint64)
        return (hi << np.uint64(32)) | lo
    assert bit_width == 32
    return bits


def uniform01(key, size, dtype=np.float64):
    """_uniform with minval=0, maxval=1: mantissa bits | 1.0, minus 1."""
    if dtype == np.float64:
        bits = random_bits(key, 64, size)
        f = ((bits >> np.uint64(12)) | np.float64(1.0).view(np.uint64)).view(np.float64)
        return f - 1.0
    bits = random_bits(key, 32, size)
    f = ((bits >> np.uint32(9)) | np.float32(1.0).view(np.uint32)).view(np.float32)
    return f - np.float32(1.0)


def uniform(key, size, minval=0.0, maxval=1.0, dtype=np.float64):
    f = uniform01(key, size, dtype)
    minval, maxval = dtype(minval), dtype(maxval)
    return np.maximum(minval, f * (maxval - minval) + minval)


def normal(key, size, dtype=np.float64):
    """_normal_real: sqrt(2) * erfinv(uniform(nextafter(-1,0), 1)).

    XLA's erf_inv is a polynomial (Giles); here scipy.special.erfinv (correct to ~1 ulp) is
    used, so normals agree with JAX's to a few ulp, not bitwise -- stated in DESIGN.md.
    """
    from scipy.special import erfinv
    lo = np.nextafter(dtype(-1.0), dtype(0.0))
    u = uniform(key, size, lo, 1.0, dtype)
    return (dtype(np.sqrt(2)) * erfinv(u)).astype(dtype)


def fold_in(key, data):
    """jax.random.fold_in(key, data) = threefry_2x32(key, PRNGKey(data)), data a 32-bit integer."""
    o0, o1 = threefry2x32(key[0], key[1], np.zeros(1, np.uint32), np.array([int(data) & 0xFFFFFFFF], np.uint32))
    return np.array([o0[0], o1[0]], dtype=np.uint32)


def fold_in_str(key, name):
    """flax 0.3.6 flax/core/scope.py `_fold_in_str` (the unvendored dependency that seeds net.init, var_state.py:123):
    fold in int.from_bytes(sha1(name)[:4], 'big')."""
    import hashlib
    return fold_in(key, int.from_bytes(hashlib.sha1(name.encode("utf-8")).digest()[:4], byteorder="big"))


def split_each(keys, num):
    """jax.vmap(lambda k: jax.random.split(k, num)) over an (N, 2) array of keys -> (N, num, 2)
    (exact_dyn.py:71 under the vmap of :82)."""
    keys = np.asarray(keys, dtype=np.uint32)
    cnt = np.arange(2 * num, dtype=np.uint32)
    o0, o1 = threefry2x32(keys[:, 0:1], keys[:, 1:2], cnt[None, :num], cnt[None, num:])
    return np.concatenate([o0, o1], axis=1).reshape(-1, num, 2)


def normal_each(keys, size):
    """jax.vmap(lambda k: jax.random.normal(k, (size,))) in float64 over an (N, 2) array of keys -> (N, size)
    (exact_dyn.py:59,66 under the vmap of :82)."""
    from scipy.special import erfinv
    keys = np.asarray(keys, dtype=np.uint32)
    cnt = np.arange(2 * size, dtype=np.uint32)
    o0, o1 = threefry2x32(keys[:, 0:1], keys[:, 1:2], cnt[None, :size], cnt[None, size:])
    bits = (o0.astype(np.uint64) << np.uint64(32)) | o1.astype(np.uint64)
    f = ((bits >> np.uint64(12)) | np.float64(1.0).view(np.uint64)).view(np.float64) - 1.0
    lo = np.nextafter(np.float64(-1.0), np.float64(0.0))
    u = np.maximum(lo, f * (np.float64(1.0) - lo) + lo)
    return np.sqrt(2.0) * erfinv(u)


def shuffle(key, n):
    """jax.random.permutation(key, n) via _shuffle (sort by random 32-bit keys, `rounds` times)."""
    x = np.arange(n)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    for _ in range(rounds):
        key, sub = split(key)
        sk = random_bits(sub, 32, n)
        x = x[np.argsort(sk, kind="stable")]
    return x


def choice_no_replace(key, n, k):
    """jax.random.choice(key, n, shape=(k,), replace=False) (var_state.py:116)."""
    return shuffle(key, n)[:k]
