"""Independent NumPy mirror of the integer VIF / motion / ADM and float VIF / ADM / MS-SSIM arithmetic -- TEST INFRASTRUCTURE.

A second restatement of libvmaf's published algorithm (integer_vif.c, integer_motion.c, vif.c / vif_tools.c; reached
from the reference at app/vmaf_analyzer.py:417), written separately from
oracle/vmaf_oracle.c: whole-plane array operations instead of per-pixel loops, Python integers / float64 where libvmaf
uses 64-bit or double.  tests/test_oracle_mirror.py checks the C oracle against it accumulator for accumulator, so a
slip in one of the two shows up.  PARITY UNPINNED all the same: both restate the same third-party source from
memory; only a libvmaf log (tests/golden/libvmaf/) pins either.

Only tests/ may import this module."""
from __future__ import annotations

import numpy as np

VIF_TAPS = [
    [489, 935, 1640, 2640, 3896, 5274, 6547, 7455, 7784, 7455, 6547, 5274, 3896, 2640, 1640, 935, 489],
    [1244, 3663, 7925, 12590, 14692, 12590, 7925, 3663, 1244],
    [3571, 16004, 26386, 16004, 3571],
    [10904, 43728, 10904],
]
MOTION_TAPS = [3571, 16004, 26386, 16004, 3571]


def _pad_reflect101(a: np.ndarray, r: int, axis: int) -> np.ndarray:
    """integer_vif.c pads by reflection WITHOUT repeating the edge sample: index -i <- i, n-1+i <- n-1-i."""
    return np.pad(a, [(r, r) if k == axis else (0, 0) for k in range(a.ndim)], mode="reflect")


def _pad_mirror(a: np.ndarray, r: int, axis: int) -> np.ndarray:
    """integer_motion.c / the float convolutions: -i <- i at the top/left, n-1+i <- n-i at the bottom/right
    (index >= n maps to 2n - i - 1, i.e. the edge sample IS repeated there)."""
    n = a.shape[axis]
    idx = np.arange(-r, n + r)
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= n, 2 * n - idx - 1, idx)
    return np.take(a, idx, axis=axis)


def _fir(a: np.ndarray, taps, axis: int, pad) -> np.ndarray:
    """sum_k taps[k] * a[i - r + k] along `axis` (object-free: int64 / uint64 is wide enough for every caller)."""
    r = len(taps) // 2
    p = pad(a, r, axis)
    n = a.shape[axis]
    out = np.zeros_like(a)
    for k, t in enumerate(taps):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(k, k + n)
        out = out + p[tuple(sl)] * t
    return out


def _log2_table() -> np.ndarray:
    t = np.zeros(65536, np.int64)
    i = np.arange(32767, 65536)
    # C round(): ties away from zero (log2f(i) * 2048 lands exactly on .5 for 39 entries; round-half-even would differ there)
    v = np.log2(i.astype(np.float64)).astype(np.float32).astype(np.float64) * 2048.0
    t[32767:] = np.floor(v + 0.5).astype(np.int64)
    return t


_LOG2 = None


def _best16_from32(v: np.ndarray):
    """top 16 bits of a 32-bit value and the (negative) shift that produced them"""
    v = v.astype(np.int64)
    nbits = np.floor(np.log2(np.maximum(v, 1).astype(np.float64))).astype(np.int64) + 1      # bit length (exact below 2^53)
    k = nbits - 16                                   # = 16 - clz32(v)
    out = np.where(k >= 0, v >> np.maximum(k, 0), v << np.maximum(-k, 0))
    return out, -k


def _bitlen64(v) -> int:
    return int(v).bit_length()


def _best16_from64(values):
    """libvmaf get_best16_from64 on Python integers (general form, all three branches)."""
    outs, xs = [], []
    for v in values:
        v = int(v)
        k = 64 - v.bit_length()                      # clz64
        if k > 48:
            sh = k - 48
            outs.append((v << sh) & 0xFFFF); xs.append(sh)
        elif k < 47:
            sh = 48 - k
            outs.append((v >> sh) & 0xFFFF); xs.append(-sh)
        else:
            x = 0
            if v >> 16:
                v >>= 1
                x = -1
            outs.append(v & 0xFFFF); xs.append(x)
    return np.array(outs, np.int64), np.array(xs, np.int64)


def vif_int(ref: np.ndarray, dis: np.ndarray, bpc: int = 8, egl: float = 100.0) -> np.ndarray:
    """-> int64 [4, 7]: num_log, den_log, num_non_log, den_non_log, accum_x, accum_x2, count (= num_accum_x)
    per scale, the accumulators of integer_vif.c's vif_statistic_8/16 (layout of oracle.vif()['acc'])."""
    global _LOG2
    if _LOG2 is None:
        _LOG2 = _log2_table()
    x = ref.astype(np.int64)
    y = dis.astype(np.int64)
    acc = np.zeros((4, 7), np.int64)
    for scale in range(4):
        taps = VIF_TAPS[scale]
        if scale == 0:
            sh, rnd = bpc, 1 << (bpc - 1)
            sh_sq = 2 * (bpc - 8)
            rnd_sq = 0 if bpc == 8 else 1 << (sh_sq - 1)
        else:
            # subsample: THIS scale's taps on the previous level (vertical rounding as that level's statistic), keep even samples
            psh, prnd = (bpc, 1 << (bpc - 1)) if scale == 1 else (16, 32768)
            lv = []
            for p in (x, y):
                v = (_fir(p, taps, 0, _pad_reflect101) + prnd) >> psh
                hh = (_fir(v, taps, 1, _pad_reflect101) + 32768) >> 16
                lv.append(hh[: (p.shape[0] // 2) * 2: 2, : (p.shape[1] // 2) * 2: 2])
            x, y = lv
            sh, rnd, sh_sq, rnd_sq = 16, 32768, 16, 32768
        # vertical pass
        mu1 = (_fir(x, taps, 0, _pad_reflect101) + rnd) >> sh
        mu2 = (_fir(y, taps, 0, _pad_reflect101) + rnd) >> sh
        xx = (_fir(x * x, taps, 0, _pad_reflect101) + rnd_sq) >> sh_sq
        yy = (_fir(y * y, taps, 0, _pad_reflect101) + rnd_sq) >> sh_sq
        xy = (_fir(x * y, taps, 0, _pad_reflect101) + rnd_sq) >> sh_sq
        # horizontal pass: mu unshifted (u32), second moments (acc + 32768) >> 16
        mu1 = _fir(mu1, taps, 1, _pad_reflect101)
        mu2 = _fir(mu2, taps, 1, _pad_reflect101)
        xx = (_fir(xx, taps, 1, _pad_reflect101) + 32768) >> 16
        yy = (_fir(yy, taps, 1, _pad_reflect101) + 32768) >> 16
        xy = (_fir(xy, taps, 1, _pad_reflect101) + 32768) >> 16
        assert mu1.max() < 2 ** 32 and xx.max() < 2 ** 32
        # Python-int products for the three (a*b + 2^31) >> 32 terms: up to 2^64
        def mulhi_round(a, b):
            return ((a.astype(object) * b.astype(object) + 2147483648) >> 32).astype(np.int64)
        s1 = (xx - mulhi_round(mu1, mu1))
        s2 = np.maximum(yy - mulhi_round(mu2, mu2), 0)
        s12 = (xy - mulhi_round(mu1, mu2))
        # the C code keeps these in int32: wrap like it does
        s1 = ((s1 + 2 ** 31) % 2 ** 32) - 2 ** 31
        s12 = ((s12 + 2 ** 31) % 2 ** 32) - 2 ** 31
        nsq = 2 * 65536
        lg = s1 >= nsq
        d16, dx = _best16_from32((s1[lg] + nsq))
        acc[scale, 1] = int(_LOG2[d16].sum())
        acc[scale, 4] = int(dx.sum())
        acc[scale, 6] = int(lg.sum())
        gp = lg & (s12 > 0) & (s2 > 0)
        g = s12[gp].astype(np.float64) / (s1[gp].astype(np.float64) + 65536 * 1.0e-10)
        sv = np.trunc(s2[gp].astype(np.float64) - g * s12[gp].astype(np.float64)).astype(np.int64)
        sv = ((sv + 2 ** 31) % 2 ** 32) - 2 ** 31            # (int32_t) conversion of an in-range double
        sv = np.maximum(sv, 0)
        g = np.minimum(g, egl)
        n1 = sv + nsq
        n2 = np.trunc(g * g * s1[gp].astype(np.float64)).astype(np.int64) + n1
        t2, x1 = _best16_from64(n2)
        t1, x2 = _best16_from64(n1)
        acc[scale, 0] = int((_LOG2[t2] - _LOG2[t1]).sum())
        acc[scale, 5] = int((x2 - x1).sum())
        acc[scale, 2] = int(s2[~lg].sum())
        acc[scale, 3] = int((~lg).sum())
    return acc


def vif_scores(acc: np.ndarray) -> np.ndarray:
    """per-scale num / den as integer_vif.c finishes them (float num, den)."""
    out = np.zeros(4)
    for s in range(4):
        a = [int(v) for v in acc[s]]
        num = np.float32(a[0] / 2048.0 + a[5] + (a[3] - (a[2] / 16384.0) / 65025.0))
        den = np.float32(a[1] / 2048.0 - (a[4] + a[6] * 17) + a[3])
        out[s] = float(num / den)
    return out


def motion_blur_int(luma: np.ndarray, bpc: int = 8) -> np.ndarray:
    p = luma.astype(np.int64)
    v = (_fir(p, MOTION_TAPS, 0, _pad_mirror) + (1 << (bpc - 1))) >> bpc
    return ((_fir(v, MOTION_TAPS, 1, _pad_mirror) + 32768) >> 16).astype(np.uint16)


def motion_sad_int(a: np.ndarray, b: np.ndarray) -> int:
    return int(np.abs(a.astype(np.int64) - b.astype(np.int64)).sum())


# ---------------------------------------------------------------------------------------------
# float VIF with libvmaf's accumulation types: fp32 everywhere, per-row fp32 sums, rows added in double
# ---------------------------------------------------------------------------------------------
def _gauss(n: int) -> np.ndarray:
    """vif kernels: normalised Gaussian, sigma = n / 5, evaluated in double and rounded to float (vif_options.h tables)"""
    k = np.arange(n) - n // 2
    g = np.exp(-(k.astype(np.float64) ** 2) / (2.0 * (n / 5.0) ** 2))
    return (g / g.sum()).astype(np.float32)


def _fir_f32(a: np.ndarray, taps: np.ndarray, axis: int) -> np.ndarray:
    """libvmaf's scalar convolution order: acc = 0; acc += f[k] * v[k] left to right, every operation rounded to fp32"""
    r = len(taps) // 2
    p = _pad_mirror(a, r, axis)
    n = a.shape[axis]
    out = np.zeros(a.shape, np.float32)
    for k, t in enumerate(taps):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(k, k + n)
        out = (out + (p[tuple(sl)] * np.float32(t)).astype(np.float32)).astype(np.float32)
    return out


def _log2f_approx(x: np.ndarray) -> np.ndarray:
    """vif_tools.c log2f_approx: exponent + degree-8 polynomial of the mantissa (Horner, fp32)"""
    u = x.astype(np.float32).view(np.uint32)
    e = ((u >> 23) & 0xFF).astype(np.int32) - 127
    m = ((u & 0x007FFFFF) | 0x3F800000).view(np.float32) - np.float32(1.0)
    c = [-0.012671635276421, 0.064841182402670, -0.157048836463065, 0.257167726303123, -0.353800560300520,
         0.480131410397451, -0.721314327952201, 1.442694803896991, 0.0]
    v = np.zeros(x.shape, np.float32)
    for ck in c:
        v = ((v * m).astype(np.float32) + np.float32(ck)).astype(np.float32)
    return (e.astype(np.float32) + v).astype(np.float32)


def vif_float(ref_f: np.ndarray, dis_f: np.ndarray, egl: float = 100.0, row_acc=np.float32, taps_of=None):
    """-> (num[4], den[4]) doubles.  ref_f / dis_f: float32 luma - 128.  `row_acc` is the type of the per-row sums:
    np.float32 is libvmaf's vif_statistic_s (float accum_num / accum_den per row, rows added into double);
    np.float64 reproduces oracle/vmaf_float_oracle.c, which sums every pixel into double."""
    x, y = ref_f.astype(np.float32), dis_f.astype(np.float32)
    num, den = np.zeros(4), np.zeros(4)
    f32 = np.float32
    for scale in range(4):
        taps = taps_of(scale) if taps_of else _gauss([17, 9, 5, 3][scale])      # literal tables when the caller has them
        if scale > 0:
            x = _fir_f32(_fir_f32(x, taps, 0), taps, 1)[: (x.shape[0] // 2) * 2: 2, : (x.shape[1] // 2) * 2: 2]
            y = _fir_f32(_fir_f32(y, taps, 0), taps, 1)[: (y.shape[0] // 2) * 2: 2, : (y.shape[1] // 2) * 2: 2]
        F = lambda p: _fir_f32(_fir_f32(p, taps, 0), taps, 1)
        mu1, mu2 = F(x), F(y)
        xx, yy, xy = F((x * x).astype(f32)), F((y * y).astype(f32)), F((x * y).astype(f32))
        s1 = np.maximum((xx - (mu1 * mu1).astype(f32)).astype(f32), f32(0))
        s2 = np.maximum((yy - (mu2 * mu2).astype(f32)).astype(f32), f32(0))
        s12 = (xy - (mu1 * mu2).astype(f32)).astype(f32)
        eps, nsq = f32(1.0e-10), f32(2.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = (s12 / (s1 + eps).astype(f32)).astype(f32)
            sv = (s2 - (g * s12).astype(f32)).astype(f32)
            c1 = s1 < eps
            g = np.where(c1, f32(0), g); sv = np.where(c1, s2, sv)
            c2 = s2 < eps
            g = np.where(c2, f32(0), g); sv = np.where(c2, f32(0), sv)
            c3 = g < 0
            sv = np.where(c3, s2, sv); g = np.where(c3, f32(0), g)
            sv = np.maximum(sv, eps)
            g = np.minimum(g, f32(egl))
            s1z = np.where(c1, f32(0), s1)
            nv = _log2f_approx((f32(1) + (((g * g).astype(f32) * s1z).astype(f32) / (sv + nsq).astype(f32)).astype(f32)).astype(f32))
            dv = _log2f_approx((f32(1) + (s1z / nsq).astype(f32)).astype(f32))
        nv = np.where(s12 < 0, f32(0), nv)
        flat = s1z < nsq
        nv = np.where(flat, (f32(1) - (s2 * f32(4.0 / (255.0 * 255.0))).astype(f32)).astype(f32), nv).astype(f32)
        dv = np.where(flat, f32(1), dv).astype(f32)
        if row_acc is np.float32:
            # sequential fp32 sum along each row (cumsum is left to right), rows added in double
            num[scale] = float(np.cumsum(nv, axis=1, dtype=np.float32)[:, -1].astype(np.float64).sum())
            den[scale] = float(np.cumsum(dv, axis=1, dtype=np.float32)[:, -1].astype(np.float64).sum())
        else:
            num[scale] = float(nv.astype(np.float64).sum())
            den[scale] = float(dv.astype(np.float64).sum())
    return num, den


# ---------------------------------------------------------------------------------------------
# integer ADM, scale 0: the db2 wavelet decomposition (integer_adm.c adm_dwt2_8 / adm_dwt2_16)
# ---------------------------------------------------------------------------------------------
DWT_LO = [15826, 27411, 7345, -4240]
DWT_HI = [-4240, -7345, 27411, -15826]


def _dwt_index(n_out: int, n: int) -> np.ndarray:
    """input indices 2i-1, 2i, 2i+1, 2i+2 of output i, folded back at both ends: -1 -> 1; n + k -> n - 1 - k"""
    idx = 2 * np.arange(n_out)[:, None] + np.arange(-1, 3)[None, :]
    idx = np.abs(idx)
    return np.where(idx >= n, 2 * n - idx - 1, idx)


def adm_dwt_scale0(luma: np.ndarray, bpc: int = 8) -> np.ndarray:
    """-> int64 [4, (h+1)//2, (w+1)//2]: bands a, v, h, d of the picture as int16 values.  Vertical pass: the low band is
    re-centred by subtracting sum(lo) * 2^(bpc-1) (pixels are unsigned), both bands rounded with 2^(bpc-1) and shifted
    by bpc; horizontal pass: (acc + 32768) >> 16.  a = lo.lo, v = hi along x of the vertical low band, h = lo along x
    of the vertical high band, d = hi.hi."""
    p = luma.astype(np.int64)
    h, w = p.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    iy, jx = _dwt_index(oh, h), _dwt_index(ow, w)
    half = 1 << (bpc - 1)
    rows = p[iy]                                          # [oh, 4, w]
    lo = (np.tensordot(rows, np.array(DWT_LO), axes=([1], [0])) - sum(DWT_LO) * half + half) >> bpc
    hi = (np.tensordot(rows, np.array(DWT_HI), axes=([1], [0])) + half) >> bpc
    to16 = lambda a: ((a + 32768) % 65536) - 32768        # the C code stores these in int16_t
    lo, hi = to16(lo), to16(hi)
    out = []
    for src, taps in ((lo, DWT_LO), (lo, DWT_HI), (hi, DWT_LO), (hi, DWT_HI)):
        cols = src[:, jx]                                 # [oh, ow, 4]
        out.append(to16((cols @ np.array(taps) + 32768) >> 16))
    return np.stack(out)


# ---------------------------------------------------------------------------------------------
# integer ADM, all four scales (integer_adm.c): DWT of the approximation band, decoupling through the Q30 reciprocal
# table, fixed-point CSF, 3x3 contrast-masking threshold, per-row rounded cube sums -- whole-band int64 array code
# ---------------------------------------------------------------------------------------------
def _wrap(a: np.ndarray, bits: int) -> np.ndarray:
    """what storing into intN_t does on two's-complement machines"""
    half = 1 << (bits - 1)
    return ((a + half) % (1 << bits)) - half


def _adm_dwt_int(a_prev: np.ndarray, scale: int) -> np.ndarray:
    """scales 1..3: the previous approximation band (int32 values) -> a, v, h, d.  Shifts after the vertical / horizontal
    pass: scale 1: 0 / 15 (round 16384), scale 2: 16 / 16, scale 3: 16 / 15; 64-bit accumulation."""
    sh_v, rnd_v = [(0, 0), (16, 32768), (16, 32768)][scale - 1]
    sh_h, rnd_h = [(15, 16384), (16, 32768), (15, 16384)][scale - 1]
    h, w = a_prev.shape
    iy, jx = _dwt_index((h + 1) // 2, h), _dwt_index((w + 1) // 2, w)
    rows = a_prev.astype(np.int64)[iy]                     # [oh, 4, w]
    lo = _wrap((np.tensordot(rows, np.array(DWT_LO), axes=([1], [0])) + rnd_v) >> sh_v, 32)
    hi = _wrap((np.tensordot(rows, np.array(DWT_HI), axes=([1], [0])) + rnd_v) >> sh_v, 32)
    return np.stack([_wrap((src[:, jx] @ np.array(taps) + rnd_h) >> sh_h, 32)
                     for src, taps in ((lo, DWT_LO), (lo, DWT_HI), (hi, DWT_LO), (hi, DWT_HI))])


def _adm_recip(mag: np.ndarray) -> np.ndarray:
    """div_lookup[32768 + m] for m in 0..32768: floor(2^30 / m), 0 at m == 0"""
    return np.where(mag > 0, (1 << 30) // np.maximum(mag, 1), 0)


def _adm_gain_k(o: np.ndarray, t: np.ndarray, scale: int) -> np.ndarray:
    """k = clamp(t / o, 0, 1) in Q15.  Scale 0 (int16 bands): table entry of o itself; scales 1..3: |o| reduced to its 15
    leading bits (rounded), the entry applied with o's sign and the extra shift."""
    if scale == 0:
        k = (np.sign(o) * _adm_recip(np.abs(o)) * t + 16384) >> 15
    else:
        ao = np.abs(o)
        nbits = np.floor(np.log2(np.maximum(ao, 1).astype(np.float64))).astype(np.int64) + 1
        sh = np.where(ao < 32768, 0, nbits - 15)
        msb = np.where(ao < 32768, ao, (ao + (1 << np.maximum(sh - 1, 0))) >> sh)
        k = (_adm_recip(msb) * t * np.sign(o) + (1 << (14 + sh))) >> (15 + sh)
    return np.where(o == 0, 32768, np.clip(k, 0, 32768))


def _rows_exact(a: np.ndarray) -> list:
    """exact per-row sums of non-negative int64 values as Python integers (the C code keeps them in 64 bits)"""
    lo, hi = a & 0xFFFFFFFF, a >> 32
    return [int(l) + (int(h_) << 32) for l, h_ in zip(lo.sum(axis=1), hi.sum(axis=1))]


def adm_int(ref: np.ndarray, dis: np.ndarray, bpc: int = 8, egl: float = 100.0, view_dist: float = 3.0,
            display_h: int = 1080):
    """-> (cm int [4][3], den int [4][3]): the contrast-masked numerator and the CSF denominator accumulators of integer ADM
    per scale for the bands (h, v, d) -- layout of oracle.adm()['cm'] / ['den']."""
    f32, f64, i64 = np.float32, np.float64, np.int64
    cos2 = f32(np.cos(np.pi / 180.0) * np.cos(np.pi / 180.0))
    rb, db = adm_dwt_scale0(ref, bpc), adm_dwt_scale0(dis, bpc)
    cms, dens = [], []
    for scale in range(4):
        if scale:
            rb, db = _adm_dwt_int(rb[0], scale), _adm_dwt_int(db[0], scale)
        h, w = rb[0].shape
        O, T = [rb[2], rb[1], rb[3]], [db[2], db[1], db[3]]          # (h, v, d); stack order is a, v, h, d
        rf = rfactor(scale, view_dist, display_h)
        if scale == 0:
            i_rf = [36453, 36453, 49417] if abs(view_dist * display_h - 3240.0) < 1e-8 else \
                   [int(float(rf[0]) * 2 ** 21) & 0xFFFF, int(float(rf[1]) * 2 ** 21) & 0xFFFF, int(float(rf[2]) * 2 ** 23) & 0xFFFF]
        else:
            i_rf = [int(float(r) * 2.0 ** 32) for r in rf]
        # angle test: inner product and squared magnitudes of the (h, v) vectors, through float then double as the C does
        q = lambda v: v.astype(f32).astype(f64) / 4096.0
        ot, om, tm = q(O[0] * T[0] + O[1] * T[1]), q(O[0] * O[0] + O[1] * O[1]), q(T[0] * T[0] + T[1] * T[1])
        flag = (ot >= 0) & (ot * ot >= f64(cos2) * om * tm)
        R, CA, CF = [], [], []
        for b in range(3):
            o, t = O[b], T[b]
            k = _adm_gain_k(o, t, scale)
            rst = (k * o + 16384) >> 15
            if scale == 0:
                rst = _wrap(rst, 16)
            sgn = (k.astype(f32) / f32(32768)) * (o.astype(f32) / f32(64))
            v, tt = rst.astype(f64) * egl, t.astype(f64)
            lim = np.where(sgn > 0, np.minimum(v, tt), np.maximum(v, tt))
            rst = np.where(flag & (sgn != 0), np.trunc(lim).astype(i64), rst)
            a = t - rst
            if scale == 0:
                rst, a = _wrap(rst, 16), _wrap(a, 16)
                ca = _wrap((_wrap(i_rf[b] * a, 32) + [16384, 16384, 65536][b]) >> [15, 15, 17][b], 16)
                cf = _wrap((4369 * np.abs(ca) + 2048) >> 12, 16)
            else:
                ca = _wrap((i_rf[b] * a + (1 << 27)) >> 28, 32)
                cf = _wrap((143165577 * np.abs(ca) + (1 << 31)) >> 32, 32)
            R.append(rst); CA.append(ca); CF.append(cf)
        left, top = int(w * 0.1 - 0.5), int(h * 0.1 - 0.5)
        right, bottom = w - left, h - top
        ys, xs = slice(top, bottom), slice(left, right)
        # threshold: over the three bands, the eight neighbours' |csf| / 30 plus the centre's |csf| / 15
        thr = np.zeros((bottom - top, right - left), i64)
        for b in range(3):
            cfp = _pad_mirror(_pad_mirror(CF[b], 1, 0), 1, 1)
            box = sum(cfp[top + 1 + di:bottom + 1 + di, left + 1 + dj:right + 1 + dj] for di in (-1, 0, 1) for dj in (-1, 0, 1))
            c = np.abs(CA[b][ys, xs])
            centre = _wrap((8738 * c + 2048) >> 12, 16) if scale == 0 else _wrap((286331153 * c + (1 << 31)) >> 32, 32)
            thr = _wrap(thr + _wrap(box - CF[b][ys, xs] + centre, 32), 32)
        lg = lambda n: int(np.ceil(np.log2(n)))
        if scale == 0:
            sh_sub, sh_sq = [10, 10, 12], [29, 29, 30]
            sh_cub = [int(np.ceil(np.log2(w) - 4)), int(np.ceil(np.log2(w) - 4)), int(np.ceil(np.log2(w) - 3))]
        else:
            sh_sub, sh_sq, sh_cub = [0, 0, 0], [30, 30, 30], [lg(w)] * 3
        sh_in = lg(h)
        cm = []
        for b in range(3):
            x = _wrap(R[b][ys, xs] * i_rf[b], 32) if scale == 0 else _wrap((R[b][ys, xs] * i_rf[b] + (1 << 27)) >> 28, 32)
            x = np.maximum(_wrap(np.abs(x) - _wrap(thr << sh_sub[b], 32), 32), 0)
            x_sq = _wrap((x * x + (1 << (sh_sq[b] - 1))) >> sh_sq[b], 32)
            val = (x_sq * x + int(2.0 ** (sh_cub[b] - 1))) >> sh_cub[b]
            cm.append(sum((r + int(2.0 ** (sh_in - 1))) >> sh_in for r in _rows_exact(val)))
        cms.append(cm)
        # denominator: cubes of the reference bands over the same region, rounded per row
        dn = []
        for b in range(3):
            v = np.abs(O[b][ys, xs])
            if scale == 0:
                v = v & 0xFFFF
                sh_acc = max(int(np.ceil(np.log2((bottom - top) * (right - left)) - 20)), 0)
                add_acc = (1 << (sh_acc - 1)) if sh_acc > 0 else 0
                rows = _rows_exact(v * v * v)
            else:
                sh_sqd = [31, 30, 31][scale - 1]
                sh_c, sh_acc = lg(right - left), lg(bottom - top)
                add_acc = int(2.0 ** (sh_acc - 1))
                rows = _rows_exact(((((v * v + (1 << (sh_sqd - 1))) >> sh_sqd) * v) + int(2.0 ** (sh_c - 1))) >> sh_c)
            dn.append(sum((r + add_acc) >> sh_acc for r in rows))
        dens.append(dn)
    return cms, dens


# ---------------------------------------------------------------------------------------------
# float ADM (adm.c / adm_tools.c): DWT, decoupling, CSF, contrast masking -- whole-plane fp32 array code
# ---------------------------------------------------------------------------------------------
DWT_LO_F = np.array([0.482962913144690, 0.836516303737469, 0.224143868041857, -0.129409522550921], np.float32)
DWT_HI_F = np.array([-0.129409522550921, -0.224143868041857, 0.836516303737469, -0.482962913144690], np.float32)


def rfactor(scale: int, view_dist: float = 3.0, display_h: int = 1080) -> np.ndarray:
    """1 / Q(scale, theta) of Watson's 9/7 DWT quantisation model, theta = 1 for h and v, 2 for d; float temporaries
    around double libm calls as adm_tools.h writes them."""
    f32 = np.float32
    amp = [[0.62171, 0.67234, 0.72709, 0.67234], [0.34537, 0.41317, 0.49428, 0.41317],
           [0.18004, 0.22727, 0.28688, 0.22727], [0.091401, 0.11792, 0.15214, 0.11792]]
    g = [1.501, 1.0, 0.534, 1.0]
    out = []
    for theta in (1, 1, 2):
        r = f32(view_dist * display_h * np.pi / 180.0)
        temp = f32(np.log10(2.0 ** (scale + 1) * float(f32(0.401)) * float(f32(g[theta])) / float(r)))
        ktt = f32(f32(f32(0.466) * temp) * temp)             # float * float * float: evaluated in float
        q = f32(2.0 * float(f32(0.495)) * 10.0 ** float(ktt) / float(f32(amp[scale][theta])))
        out.append(f32(1.0) / q)
    return np.array(out, np.float32)


def _dwt_f32(p: np.ndarray):
    """one level of the float db2 decomposition; accumulation order lo/hi[0]*s0 + [1]*s1 + [2]*s2 + [3]*s3, fp32 each step"""
    h, w = p.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    iy, jx = _dwt_index(oh, h), _dwt_index(ow, w)
    f32 = np.float32

    def comb(vals, taps):                                  # vals: [..., 4] along the last axis
        acc = np.zeros(vals.shape[:-1], f32)
        for k in range(4):
            acc = (acc + (vals[..., k] * taps[k]).astype(f32)).astype(f32)
        return acc

    rows = np.moveaxis(p[iy], 1, -1)                       # [oh, w, 4]
    lo, hi = comb(rows, DWT_LO_F), comb(rows, DWT_HI_F)
    a, v = comb(lo[:, jx], DWT_LO_F), comb(lo[:, jx], DWT_HI_F)
    hb, d = comb(hi[:, jx], DWT_LO_F), comb(hi[:, jx], DWT_HI_F)
    return a, v, hb, d


def adm_float(ref_f: np.ndarray, dis_f: np.ndarray, egl: float = 100.0, rf_of=None):
    """-> (num_scale[4], den_scale[4], adm2).  ref_f / dis_f: float32 luma - 128.  `rf_of(scale)` supplies the CSF factors
    (default: rfactor())."""
    f32 = np.float32
    x, y = ref_f.astype(f32), dis_f.astype(f32)
    H0, W0 = x.shape
    cos2 = f32(np.cos(np.pi / 180.0) * np.cos(np.pi / 180.0))
    eps, by30, by15 = f32(1e-30), f32(0.0333333351), f32(0.0666666701)
    nums, dens = np.zeros(4), np.zeros(4)
    for scale in range(4):
        ra, rv, rh, rd = _dwt_f32(x)
        da, dv, dh, dd = _dwt_f32(y)
        x, y = ra, da
        h, w = ra.shape
        rf = (rf_of or rfactor)(scale).astype(f32)
        O, T = [rh, rv, rd], [dh, dv, dd]
        ot = ((rh * dh).astype(f32) + (rv * dv).astype(f32)).astype(f32)
        om = ((rh * rh).astype(f32) + (rv * rv).astype(f32)).astype(f32)
        tm = ((dh * dh).astype(f32) + (dv * dv).astype(f32)).astype(f32)
        flag = (ot >= 0) & ((ot * ot).astype(f32) >= ((cos2 * om).astype(f32) * tm).astype(f32))
        R, CA, CF = [], [], []
        for b in range(3):
            o, t = O[b], T[b]
            with np.errstate(divide="ignore", invalid="ignore"):
                k = np.clip((t / (o + eps).astype(f32)).astype(f32), f32(0), f32(1))
            rst = (k * o).astype(f32)
            v = rst.astype(np.float64) * egl
            t64 = t.astype(np.float64)
            lim = np.where(rst > 0, np.minimum(v, t64), np.where(rst < 0, np.maximum(v, t64), rst.astype(np.float64)))
            rst = np.where(flag, lim, rst.astype(np.float64)).astype(f32)
            ca = (rf[b] * (t - rst).astype(f32)).astype(f32)
            R.append(rst); CA.append(ca); CF.append((by30 * np.abs(ca)).astype(f32))
        left, top = int(w * 0.1 - 0.5), int(h * 0.1 - 0.5)
        right, bottom = w - left, h - top
        ys, xs = slice(top, bottom), slice(left, right)
        thr = np.zeros((bottom - top, right - left), f32)
        for b in range(3):
            s = np.zeros_like(thr)
            cfp = _pad_mirror(_pad_mirror(CF[b], 1, 0), 1, 1)        # neighbours past the band edge fold back (small bands)
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    if di == 0 and dj == 0:
                        term = (by15 * np.abs(CA[b][ys, xs])).astype(f32)
                    else:
                        term = cfp[top + 1 + di:bottom + 1 + di, left + 1 + dj:right + 1 + dj]
                    s = (s + term).astype(f32)
            thr = (thr + s).astype(f32)
        area = f32(np.power(f32((bottom - top) * (right - left)) / f32(32.0), f32(1.0 / 3.0), dtype=f32))
        ns = ds = f32(0)
        for b in range(3):
            xn = np.maximum((np.abs((R[b][ys, xs] * rf[b]).astype(f32)) - thr).astype(f32), f32(0))
            cn = ((xn * xn).astype(f32) * xn).astype(f32)
            vd = (np.abs(O[b][ys, xs]) * rf[b]).astype(f32)
            cd = ((vd * vd).astype(f32) * vd).astype(f32)
            # adm_tools.c order: a float sum per row, the rows added into another float
            an = np.cumsum(np.cumsum(cn, axis=1, dtype=f32)[:, -1], dtype=f32)[-1]
            ad = np.cumsum(np.cumsum(cd, axis=1, dtype=f32)[:, -1], dtype=f32)[-1]
            ns = f32(ns + f32(np.power(an, f32(1.0 / 3.0), dtype=f32) + area))
            ds = f32(ds + f32(np.power(ad, f32(1.0 / 3.0), dtype=f32) + area))
        nums[scale], dens[scale] = float(ns), float(ds)
    limit = 1e-10 * (W0 * H0) / (1920.0 * 1080.0)
    num, den = nums.sum(), dens.sum()
    num, den = (0.0 if num < limit else num), (0.0 if den < limit else den)
    return nums, dens, (1.0 if den == 0.0 else num / den)


# ---------------------------------------------------------------------------------------------
# iqa MS-SSIM (ms_ssim.c): float64 throughout -- a tolerance check of the fp32 oracle, not a bit check
# ---------------------------------------------------------------------------------------------
GAUSS11 = np.array([0.001028, 0.007599, 0.036001, 0.109361, 0.213006, 0.266012, 0.213006, 0.109361, 0.036001, 0.007599,
                    0.001028])
LPF9 = np.array([0.026727, -0.016828, -0.078201, 0.266846, 0.602914, 0.266846, -0.078201, -0.016828, 0.026727])


def _valid11(a: np.ndarray) -> np.ndarray:
    h, w = a.shape
    t = sum(a[:, k:w - 10 + k] * GAUSS11[k] for k in range(11))
    return sum(t[k:h - 10 + k, :] * GAUSS11[k] for k in range(11))


def _sym_index(n_out: int, n: int, step: int, off: int, taps: int) -> np.ndarray:
    idx = step * np.arange(n_out)[:, None] + off + np.arange(taps)[None, :]
    idx = np.where(idx < 0, -1 - idx, idx)
    return np.where(idx >= n, 2 * n - idx - 1, idx)


def _decimate2(a: np.ndarray) -> np.ndarray:
    """9-tap low-pass along x then y, keep every second sample, symmetric borders (-1 -> 0, n -> n - 1)"""
    h, w = a.shape
    dw, dh = w // 2 + (w & 1), h // 2 + (h & 1)
    t = a[:, _sym_index(dw, w, 2, -4, 9)] @ LPF9
    return np.tensordot(t[_sym_index(dh, h, 2, -4, 9)], LPF9, axes=([1], [0]))


def ms_ssim(ref0: np.ndarray, dis0: np.ndarray):
    """-> (score, lcs[5, 3]).  ref0 / dis0: luma as float in [0, 255]."""
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    c3 = c2 / 2
    expo = [0.0448, 0.2856, 0.3001, 0.2363, 0.1333]
    r, c = ref0.astype(np.float64), dis0.astype(np.float64)
    score, lcs = 1.0, np.zeros((5, 3))
    for s in range(5):
        if s:
            r, c = _decimate2(r), _decimate2(c)
        m1, m2 = _valid11(r), _valid11(c)
        v1 = np.maximum(_valid11(r * r) - m1 * m1, 0)
        v2 = np.maximum(_valid11(c * c) - m2 * m2, 0)
        cv = _valid11(r * c) - m1 * m2
        sr = np.sqrt(v1 * v2)
        lcs[s] = [np.mean((2 * m1 * m2 + c1) / (m1 * m1 + m2 * m2 + c1)), np.mean((2 * sr + c2) / (v1 + v2 + c2)),
                  np.mean((cv + c3) / (sr + c3))]
        score *= (lcs[s, 0] ** (expo[4] if s == 4 else 0.0)) * lcs[s, 1] ** expo[s] * lcs[s, 2] ** expo[s]
    return score, lcs


# ---------------------------------------------------------------------------------------------
# FFmpeg `ssim` filter (libavfilter/vf_ssim.c): 4x4 block sums, overlapping 8x8 windows, float row sums
# ---------------------------------------------------------------------------------------------
def ffssim_plane(a: np.ndarray, b: np.ndarray, bpc: int = 8) -> float:
    f32 = np.float32
    h, w = a.shape
    h4, w4 = h >> 2, w >> 2
    if h4 < 2 or w4 < 2:
        return 1.0
    p = a[:h4 * 4, :w4 * 4].astype(np.int64).reshape(h4, 4, w4, 4)
    q = b[:h4 * 4, :w4 * 4].astype(np.int64).reshape(h4, 4, w4, 4)
    win = lambda s: s[:-1, :-1] + s[:-1, 1:] + s[1:, :-1] + s[1:, 1:]          # four 4x4 blocks = one 8x8 window
    s1, s2 = win(p.sum(axis=(1, 3))), win(q.sum(axis=(1, 3)))
    ss, s12 = win((p * p + q * q).sum(axis=(1, 3))), win((p * q).sum(axis=(1, 3)))
    peak = (1 << bpc) - 1
    c1, c2 = int(.01 * .01 * peak * peak * 64 + .5), int(.03 * .03 * peak * peak * 64 * 63 + .5)
    var, cov = ss * 64 - s1 * s1 - s2 * s2, s12 * 64 - s1 * s2
    v = ((2 * s1 * s2 + c1).astype(f32) * (2 * cov + c2).astype(f32)) / \
        ((s1 * s1 + s2 * s2 + c1).astype(f32) * (var + c2).astype(f32))
    rows = np.cumsum(v, axis=1, dtype=f32)[:, -1]                              # vf_ssim sums a row of windows in float
    return float(rows.astype(np.float64).sum() / ((h4 - 1) * (w4 - 1)))


def sse_plane(a: np.ndarray, b: np.ndarray) -> int:
    d = a.astype(np.int64) - b.astype(np.int64)
    return int((d * d).sum())


# ---------------------------------------------------------------------------------------------
# float motion (motion.c): 5-tap blur, mean absolute difference of consecutive blurred frames
# ---------------------------------------------------------------------------------------------
MOTION_TAPS_F = np.array([0.054488685, 0.244201342, 0.402619947, 0.244201342, 0.054488685], np.float32)


def motion_blur_float(luma_f: np.ndarray, taps=MOTION_TAPS_F) -> np.ndarray:
    return _fir_f32(_fir_f32(luma_f.astype(np.float32), taps, 0), taps, 1)


def motion_sad_float(a: np.ndarray, b: np.ndarray) -> float:
    """a float sum per row, the rows added into another float, divided by the sample count in float"""
    f32 = np.float32
    rows = np.cumsum(np.abs((a - b).astype(f32)), axis=1, dtype=f32)[:, -1]
    return float(f32(np.cumsum(rows, dtype=f32)[-1] / f32(a.shape[0] * a.shape[1])))


# ---------------------------------------------------------------------------------------------
# iqa SSIM (ssim.c) at one scale, after the f x f box decimation float_ssim applies to large pictures -- float64
# ---------------------------------------------------------------------------------------------
def _decimate_box(a: np.ndarray, f: int) -> np.ndarray:
    h, w = a.shape
    dh, dw = h // f + (h & 1), w // f + (w & 1)
    iy, jx = _sym_index(dh, h, f, -(f // 2), f), _sym_index(dw, w, f, -(f // 2), f)
    return a[iy][:, :, jx].sum(axis=(1, 3)) / (f * f)


def ssim_float(ref0: np.ndarray, dis0: np.ndarray) -> float:
    """luma as float in [0, 255]; decimation factor max(1, round(min(w, h) / 256))"""
    r, c = ref0.astype(np.float64), dis0.astype(np.float64)
    f = max(1, int(np.floor(min(r.shape) / 256.0 + 0.5)))
    if f > 1:
        r, c = _decimate_box(r, f), _decimate_box(c, f)
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    m1, m2 = _valid11(r), _valid11(c)
    v1, v2, cv = _valid11(r * r) - m1 * m1, _valid11(c * c) - m2 * m2, _valid11(r * c) - m1 * m2
    return float(np.mean((2 * m1 * m2 + c1) * (2 * cv + c2) / ((m1 * m1 + m2 * m2 + c1) * (v1 + v2 + c2))))
