"""Host model of the integer accumulation that the fixed-point deposit (Depositor<DEP_FIXED>, native adds) and the
histogram kernel (k_diag_limb) perform with 32-bit shared-memory adds: the magic-number split of a scaled contribution into
{low 32 bits, I >> 32}, the carry rule (the add that wraps the low word carries one into the high word), and the
reconstruction at flush time.  numpy restates the device arithmetic word for word; the sums must be exact for any
order of arrival, for negative contributions, and at the bounds the kernels rely on."""
import numpy as np

MAGIC = 6755399441055744.0          # 1.5 * 2^52
HI_BIAS = 0x43380000                # high word of MAGIC


def split(val, scale):
    """limb_split / the native-add depositor: I = RN(val * scale) from the low mantissa bits of val * scale + MAGIC."""
    t = np.asarray(val, dtype=np.float64) * scale + MAGIC      # scale is a power of two: the product is exact
    bits = t.view(np.uint64)
    lo = (bits & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = ((bits >> np.uint64(32)).astype(np.int64) - HI_BIAS).astype(np.int64)   # floor(I / 2^32), signed
    return lo, hi


def accumulate(lo, hi, order):
    """One two-word counter receiving the contributions in the given order, 32-bit wrap-around adds as on the device."""
    c_lo, c_hi = np.uint32(0), np.uint32(0)
    with np.errstate(over="ignore"):
        for k in order:
            old = c_lo
            c_lo = np.uint32(old + lo[k])
            carry = np.uint32(1) if np.uint32(old + lo[k]) < lo[k] else np.uint32(0)
            c_hi = np.uint32(c_hi + np.uint32(hi[k] & 0xFFFFFFFF) + carry)
    return int(c_lo) + int(np.int32(c_hi)) * (1 << 32)          # limb_flush / fixed_flush


def test_split_recovers_the_rounded_integer():
    rng = np.random.default_rng(1)
    for ebits in (10, 31, 32, 33, 46):
        scale = 2.0 ** ebits
        val = rng.uniform(-1.0, 1.0, 20000)
        val[:6] = [1.0, -1.0, 0.0, 2.0 ** -ebits * 0.5, -2.0 ** -ebits * 0.5, 2.0 ** -ebits * 1.5]   # bounds and ties
        lo, hi = split(val, scale)
        got = lo.astype(np.int64) + hi * (1 << 32)
        want = np.rint(val * scale).astype(np.int64)            # round to nearest even, like the fp64 add
        assert np.array_equal(got, want), ebits


def test_two_word_counter_is_exact_in_any_order():
    rng = np.random.default_rng(2)
    for trial in range(20):
        n = 3000
        val = rng.uniform(-1.0, 1.0, n) * rng.choice([1.0, 1e-3, 1e-9], n)
        if trial % 4 == 0:
            val = np.abs(val)                                   # same sign: the low word wraps often
        lo, hi = split(val, 2.0 ** 46)
        exact = int(np.sum((lo.astype(np.int64) + hi * (1 << 32)).astype(object)))
        for order in (np.arange(n), np.arange(n)[::-1], rng.permutation(n)):
            assert accumulate(lo, hi, order) == exact


def test_high_word_capacity_of_the_histogram_counters():
    """k_diag_limb folds its counters every 64 tile steps of 1024 markers: 2^16 additions of |I| <= 2^46 (high part
    <= 2^14) plus one carry each stay inside a signed 32-bit word."""
    n = 1 << 16
    assert n * ((1 << 14) + 1) < (1 << 31)
    lo, hi = split(np.full(n, 1.0), 2.0 ** 46)                  # every contribution at the bound, same sign
    assert accumulate(lo, hi, range(n)) == n * (1 << 46)
    lo, hi = split(np.full(n, -1.0), 2.0 ** 46)
    assert accumulate(lo, hi, range(n)) == -n * (1 << 46)
    lo, hi = split(np.full(n, 1.0 - 2.0 ** -40), 2.0 ** 46)     # low word nearly full: a carry on almost every add
    assert accumulate(lo, hi, range(n)) == n * ((1 << 46) - (1 << 6))


def test_sixty_four_bit_grid_of_the_deposit():
    """The deposit keeps 64-bit sums (its scale bounds every slot sum by 2^62); the word pair behaves as one 64-bit
    two's-complement integer even while partial sums change sign."""
    rng = np.random.default_rng(3)
    val = rng.uniform(-1.0, 1.0, 5000)
    lo, hi = split(val, 2.0 ** 50)
    exact = int(np.sum(np.rint(val * 2.0 ** 50).astype(np.int64).astype(object)))
    assert accumulate(lo, hi, rng.permutation(5000)) == exact
