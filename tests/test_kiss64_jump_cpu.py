"""The product's KISS64 (pic1dp_b200/csrc/rng_kernels.cuh, the code the device runs, here through its host entry
points) against the reference's known-answer vector (/root/reference/src/multirand.F90:396-401) and against the
KAT-pinned oracle's sequential stream at large offsets: the O(log n) jump of the three recurrences must land exactly
where n sequential calls do."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from pic1dp_b200 import host as H

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "multirand_kat.json")))
DEFAULT_SEEDS = [1234567890987654321, 362436362436362436, 1066149217761810, 123456123456123456]  # :481-484


def _to_real(i):
    return float(np.int64(i)) / 18446744073709551615.0 + 0.5   # INT2REAL64, :49


def test_known_answer_vector():
    u, _ = H.host_kiss64_fill(DEFAULT_SEEDS, 10)
    assert [float(t) for t in u] == [_to_real(i) for i in GOLD["kiss64"]]


@pytest.mark.parametrize("mype", [0, 3])
@pytest.mark.parametrize("offset", [0, 1, 2, 63, 1000003, 100000000])
def test_jump_equals_sequential_stream(mype, offset):
    g = O.MultiRand()
    g.init_const(1, mype, 5)           # seed_type 1, warm-up 5: multirand_init as particle_load calls it
    seeds = g.seeds4()
    g.skip(offset)
    want = g.real_array(4096)
    got, _ = H.host_kiss64_fill(H.host_kiss64_jump(seeds, offset), 4096)
    assert np.array_equal(got, want)


def test_jump_composes_and_handles_large_seed_carry():
    # a seed carry >= 2^58 (any 64-bit value can come out of the seeding) is walked sequentially until it is in range
    seeds = [0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, 0xF123456789ABCDEF]
    a = H.host_kiss64_jump(H.host_kiss64_jump(seeds, 12345), 987654321)
    b = H.host_kiss64_jump(seeds, 12345 + 987654321)
    assert a == b
    _, s = H.host_kiss64_fill(seeds, 5000)
    assert s == H.host_kiss64_jump(seeds, 5000)


def test_counter_uniforms_are_decomposition_independent():
    u_v, u_x = H.host_counter_uniforms(42, 0, 0, 5000)
    a_v, a_x = H.host_counter_uniforms(42, 0, 0, 2000)
    b_v, b_x = H.host_counter_uniforms(42, 0, 2000, 3000)
    assert np.array_equal(np.concatenate([a_v, b_v]), u_v) and np.array_equal(np.concatenate([a_x, b_x]), u_x)
    assert 0.0 <= u_v.min() and u_v.max() <= 1.0 and abs(u_v.mean() - 0.5) < 0.02 and abs(u_x.mean() - 0.5) < 0.02
    assert abs(np.corrcoef(u_v, u_x)[0, 1]) < 0.05
    o_v, _ = H.host_counter_uniforms(43, 0, 0, 5000)
    assert not np.array_equal(o_v, u_v)
