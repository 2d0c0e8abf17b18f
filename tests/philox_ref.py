"""Host replay of the library's counter-based RNG (thylacine_b200/csrc/philox.cuh) -- test infrastructure.
Philox4x32-10, keyed and indexed exactly as rng_uniform2 / rng_normal do on the device."""
import math

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
RNG_PRIOR, RNG_PRIOR_CHI, RNG_HMC_MOMENTUM, RNG_HMC_ACCEPT = 1, 2, 3, 4
RNG_LFM_PICK, RNG_LFM_PARTNER, RNG_LFM_ACCEPT = 5, 6, 7


def philox(c0, c1, c2, c3, k0, k1):
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def u01(hi, lo):
    return float(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0)


def uniform2(seed, chain, item, purpose, step):
    k0 = ((seed & MASK) ^ (((chain >> 32) * 0x85EBCA6B) & MASK)) & MASK
    k1 = (((seed >> 32) & MASK) ^ (((step >> 32) * 0xC2B2AE35) & MASK)) & MASK
    r = philox(chain & MASK, item & MASK, step & MASK, purpose, k0, k1)
    return u01(r[0], r[1]), u01(r[2], r[3])


def normal(seed, chain, item, purpose, step):
    u0, u1 = uniform2(seed, chain, item, purpose, step)
    return math.sqrt(-2.0 * math.log(1.0 - u0)) * math.cos(2.0 * math.pi * u1)
