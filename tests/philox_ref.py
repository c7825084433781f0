"""numpy restatement of the engine's counter-based streams (test infrastructure).

Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
SC'11) and the engine's parameter draw (csrc/vaxmc_kernels.cuh: draw_sample_params), so that
CPU tests can pin the counter layout and GPU tests can check the device draws bit for bit.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
TAG_DAY, TAG_PRO_ANTI, TAG_PARAMS_A, TAG_PARAMS_B, TAG_TEST = 0, 1, 2, 3, 7
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (any broadcastable shapes)."""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) & MASK for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in (c0, c1, c2, c3)]


def u53(a, b):
    a = np.asarray(a, dtype=np.uint32)
    b = np.asarray(b, dtype=np.uint32)
    return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0
            + (b >> np.uint32(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def draw_params(bounds12, seed, first, count):
    """(params [count,6], p_soft_no [count], rejections [count]) as the device draws them."""
    b = np.asarray(bounds12, dtype=np.float64).reshape(12)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    ids = np.arange(first, first + count, dtype=np.uint64)
    s_lo, s_hi = (ids & MASK), (ids >> np.uint64(32))
    params = np.empty((count, 6))
    rej = np.zeros(count, dtype=np.int64)
    todo = np.arange(count)
    j = 0
    while len(todo):
        r = philox4x32_10(s_lo[todo], s_hi[todo], np.uint64(j), np.uint64(TAG_PRO_ANTI), k0, k1)
        pro = b[0] + (b[1] - b[0]) * u53(r[0], r[1])
        anti = b[2] + (b[3] - b[2]) * u53(r[2], r[3])
        ok = ~(pro + anti > 1.0)
        params[todo[ok], 0] = pro[ok]
        params[todo[ok], 1] = anti[ok]
        rej[todo[~ok]] += 1
        todo = todo[~ok]
        j += 1
        if j > 100000:
            raise RuntimeError("acceptance probability too small")
    ra = philox4x32_10(s_lo, s_hi, np.uint64(0), np.uint64(TAG_PARAMS_A), k0, k1)
    rb = philox4x32_10(s_lo, s_hi, np.uint64(0), np.uint64(TAG_PARAMS_B), k0, k1)
    params[:, 2] = b[4] + (b[5] - b[4]) * u53(ra[0], ra[1])
    params[:, 3] = b[6] + (b[7] - b[6]) * u53(ra[2], ra[3])
    params[:, 4] = b[8] + (b[9] - b[8]) * u53(rb[0], rb[1])
    params[:, 5] = b[10] + (b[11] - b[10]) * u53(rb[2], rb[3])
    soft = 1 - (params[:, 0] + params[:, 1])
    return params, soft, rej


# Known-answer vectors of Philox4x32-10 from the Random123 distribution (kat_vectors).
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
