"""Counter-based RNG: oracle restatement pinned on the published Random123
known-answer vectors; host-sim library agrees bit for bit; sharding invariance."""
import ctypes as C

import numpy as np

from oracle import philox
from tests.hostsim import harness

KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def test_random123_known_answers():
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def lib_uniform(seed, first, stream, n_env, n_cols):
    lib = harness.load()
    out = np.zeros((n_env, n_cols))
    assert lib.opfg_philox_uniform(seed, first, stream, n_env, n_cols, out.ctypes.data, None) == 0
    return out


def test_library_matches_oracle_bit_for_bit():
    for n_cols in (1, 2, 7):
        got = lib_uniform(1234, 5, 77, 9, n_cols)
        np.testing.assert_array_equal(got, philox.uniform(1234, 5, 77, 9, n_cols))
        assert (got >= 0).all() and (got < 1).all()


def test_sharding_invariance_and_stream_separation():
    whole = lib_uniform(99, 0, 3, 16, 6)
    np.testing.assert_array_equal(lib_uniform(99, 8, 3, 8, 6), whole[8:])
    assert not np.array_equal(lib_uniform(99, 0, 4, 16, 6), whole)
    assert not np.array_equal(lib_uniform(98, 0, 3, 16, 6), whole)
    assert abs(lib_uniform(1, 0, 1, 4096, 8).mean() - 0.5) < 0.01


def test_sample_uniform_semantics():
    """opf_env.py:278-284: state = U(low, high) / scaling."""
    lib = harness.load()
    n_env, n_state = 5, 6
    slots = np.array([4, 1, 2], np.int32)
    lo, hi, dv = np.array([0.0, -2.0, 5.0]), np.array([1.0, 2.0, 5.0]), np.array([1.0, 2.0, 0.5])
    state = np.full((n_env, n_state), -7.0)
    assert lib.opfg_sample_uniform(11, 100, 2, n_env, 3, slots.ctypes.data, lo.ctypes.data,
                                   hi.ctypes.data, dv.ctypes.data, state.ctypes.data, n_state, None) == 0
    u = philox.uniform(11, 100, 2, n_env, 3)
    np.testing.assert_allclose(state[:, slots], (lo + (hi - lo) * u) / dv, rtol=0, atol=0)
    assert (state[:, [0, 3, 5]] == -7.0).all()
