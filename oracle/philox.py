"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

NumPy restatement of the counter-based generator the engine uses in place of the
reference's ``self.np_random.uniform`` (``opfgym/opf_env.py:278``): Philox4x32-10
(Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), key =
seed, counter = (column pair, env id low, env id high, stream), two 53-bit
doubles per block.  Pinned by the published Random123 known-answer vectors in
tests/test_philox.py.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & MASK
        n1 = p1 & MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & MASK
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniform(seed: int, first_env: int, stream: int, n_env: int, n_cols: int) -> np.ndarray:
    """Same numbers as ``opfg_philox_uniform``: out[b, j] in [0, 1)."""
    pairs = (n_cols + 1) // 2
    env = (np.arange(n_env, dtype=np.uint64) + np.uint64(first_env))[:, None]
    pair = np.arange(pairs, dtype=np.uint64)[None, :]
    s_mix = ((stream & 0xFFFFFFFF) ^ (((stream >> 32) * 0x9E3779B9) & 0xFFFFFFFF)) & 0xFFFFFFFF
    o0, o1, o2, o3 = philox4x32_10(pair + np.uint64(0) * env, env & MASK + np.uint64(0) * pair,
                                   (env >> np.uint64(32)) + np.uint64(0) * pair,
                                   np.uint64(s_mix) + np.uint64(0) * (env + pair),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    x = (o1 << np.uint64(32)) | o0
    y = (o3 << np.uint64(32)) | o2
    a = (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    b = (y >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    out = np.empty((n_env, 2 * pairs))
    out[:, 0::2] = a
    out[:, 1::2] = b
    return out[:, :n_cols]
