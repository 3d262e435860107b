"""Bound *specs* and their device serialisation.

A spec is a plain dictionary holding exactly the state the reference keeps on
a bound object:

    ell  = dict(c=f64[de], B=f64[de,de], B_inv=f64[de,de])
           (``Ellipsoid.c/B/B_inv``, nautilus/bounds/basic.py:303-309)
    mix  = dict(dim_cube=bool[d], ell=ell | None)
           (``UnitCubeEllipsoidMixture``, nautilus/bounds/basic.py:491-563)
    emu  = dict(mean, scale, coefs=[[W..] per net], intercepts=[[b..] per net])
           (``NeuralNetworkEmulator``, nautilus/neural.py:74-98)
    nb   = dict(ell=ell, emulator=emu | None, score_predict_min=float)
           (``NeuralBound``, nautilus/bounds/neural.py:61-97)
    spec = dict(kind='nautilus', n_dim, unit, log_v_all=f64[K],
                mixtures=[mix]*K, neural=[nb]*J)
           (``NautilusBound`` + its outer ``Union``,
            nautilus/bounds/nautilus.py:88-144, nautilus/bounds/union.py:117-151)
    spec = dict(kind='cube', n_dim)          (``UnitCube``, basic.py:9)

``pack_stack`` turns a list of specs (a bound and the later bounds that carve
its shell, nautilus/sampler.py:796-801) into the two flat arrays the CUDA
kernels read: ``meta`` (int32) and ``data`` (float64).  The layout is
documented in ``include/nautilus_b200.h`` and mirrored by ``nb200_blob.cuh``.
"""

import numpy as np

HDR = 16        # ints in a bound record header
MIX_REC = 8     # ints per mixture record
NB_REC = 12     # ints per neural-bound record
D_MAX = 128     # largest supported dimensionality
W_MAX = 256     # largest supported hidden width


# --------------------------------------------------------------------------
# flat <-> nested (for .npz fixtures and checkpoints)
# --------------------------------------------------------------------------

def spec_to_flat(spec, prefix=''):
    """Flatten a spec into ``{str: ndarray}`` suitable for ``np.savez``."""
    out = {prefix + 'kind': np.array(spec['kind']),
           prefix + 'n_dim': np.array(spec['n_dim'])}
    if spec['kind'] == 'cube':
        return out
    out[prefix + 'unit'] = np.array(bool(spec['unit']))
    out[prefix + 'log_v_all'] = np.asarray(spec['log_v_all'], dtype=float)
    out[prefix + 'K'] = np.array(len(spec['mixtures']))
    out[prefix + 'J'] = np.array(len(spec['neural']))
    for k, mix in enumerate(spec['mixtures']):
        p = '{}mix{}_'.format(prefix, k)
        out[p + 'dim_cube'] = np.asarray(mix['dim_cube'], dtype=bool)
        if mix['ell'] is not None:
            for key in ('c', 'B', 'B_inv'):
                out[p + key] = np.asarray(mix['ell'][key], dtype=float)
    for j, nb in enumerate(spec['neural']):
        p = '{}nb{}_'.format(prefix, j)
        for key in ('c', 'B', 'B_inv'):
            out[p + key] = np.asarray(nb['ell'][key], dtype=float)
        out[p + 'score_predict_min'] = np.array(float(
            nb['score_predict_min']))
        emu = nb['emulator']
        if emu is not None:
            out[p + 'mean'] = np.asarray(emu['mean'], dtype=float)
            out[p + 'scale'] = np.asarray(emu['scale'], dtype=float)
            out[p + 'n_net'] = np.array(len(emu['coefs']))
            out[p + 'n_lay'] = np.array(len(emu['coefs'][0]))
            for n, (ws, bs) in enumerate(zip(emu['coefs'],
                                             emu['intercepts'])):
                for i, (w, b) in enumerate(zip(ws, bs)):
                    out['{}W{}_{}'.format(p, n, i)] = np.asarray(w, float)
                    out['{}b{}_{}'.format(p, n, i)] = np.asarray(b, float)
    return out


def flat_to_spec(flat, prefix=''):
    """Inverse of :func:`spec_to_flat`."""
    kind = str(flat[prefix + 'kind'])
    spec = dict(kind=kind, n_dim=int(flat[prefix + 'n_dim']))
    if kind == 'cube':
        return spec
    spec['unit'] = bool(flat[prefix + 'unit'])
    spec['log_v_all'] = np.array(flat[prefix + 'log_v_all'], dtype=float)
    spec['mixtures'] = []
    for k in range(int(flat[prefix + 'K'])):
        p = '{}mix{}_'.format(prefix, k)
        ell = None
        if p + 'c' in flat:
            ell = {key: np.array(flat[p + key], dtype=float)
                   for key in ('c', 'B', 'B_inv')}
        spec['mixtures'].append(dict(
            dim_cube=np.array(flat[p + 'dim_cube'], dtype=bool), ell=ell))
    spec['neural'] = []
    for j in range(int(flat[prefix + 'J'])):
        p = '{}nb{}_'.format(prefix, j)
        ell = {key: np.array(flat[p + key], dtype=float)
               for key in ('c', 'B', 'B_inv')}
        emu = None
        if p + 'mean' in flat:
            n_net, n_lay = int(flat[p + 'n_net']), int(flat[p + 'n_lay'])
            emu = dict(
                mean=np.array(flat[p + 'mean'], dtype=float),
                scale=np.array(flat[p + 'scale'], dtype=float),
                coefs=[[np.array(flat['{}W{}_{}'.format(p, n, i)])
                        for i in range(n_lay)] for n in range(n_net)],
                intercepts=[[np.array(flat['{}b{}_{}'.format(p, n, i)])
                             for i in range(n_lay)] for n in range(n_net)])
        spec['neural'].append(dict(
            ell=ell, emulator=emu,
            score_predict_min=float(flat[p + 'score_predict_min'])))
    return spec


# --------------------------------------------------------------------------
# spec -> (meta, data) device blob
# --------------------------------------------------------------------------

class _Data:
    def __init__(self):
        self.chunks = []
        self.n = 0

    def add(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        off = self.n
        self.chunks.append(a)
        self.n += len(a)
        if self.n % 2:                      # keep every array 16-byte aligned
            self.chunks.append(np.zeros(1))
            self.n += 1
        return off

    def array(self):
        if not self.chunks:
            return np.zeros(2)
        return np.concatenate(self.chunks)


TC_MAGIC = 0x7F32
TC_MAX_HID = 4
TC_HDR_WORDS = 32
TC_SMEM_LIMIT = 220 * 1024     # bytes of resident weights per CTA
TC_COLS = 256                  # TMEM columns per tile group (2 groups/CTA)
TC_COLS_SINGLE = 512           # ... or all of TMEM for a single group


def _round_up(v, m):
    return (v + m - 1) // m * m


def tf32_round(a):
    """Round float32 values to tf32 (10 explicit mantissa bits), nearest with
    ties away from zero like ``cvt.rna.tf32.f32``."""
    bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    return ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(
        np.float32)


def _core_matrix_layout(wt, np_, kp):
    """K-major, no-swizzle UMMA operand layout of a [n, k] matrix zero-padded
    to [np_, kp]: 8-row x 16-byte core matrices, K chunks 128 B apart, 8-row
    groups kp*32 B apart (see smem_desc() in csrc/nb200_mlp_tc.cu)."""
    full = np.zeros((np_, kp), dtype=np.float32)
    full[:wt.shape[0], :wt.shape[1]] = wt
    return full.reshape(np_ // 8, 8, kp // 4, 4).transpose(0, 2, 1, 3).ravel()


def pack_tc(emu, score_predict_min):
    """Tensor-core blob of an emulator for ``k_mlp_tf32``.

    Returns ``(header int32[32], weights float32[])`` or ``None`` if the
    architecture does not fit the kernel's envelope (<= 4 hidden layers,
    resident weights <= 220 KB, <= 256 TMEM columns per tile group)."""
    coefs, intercepts = emu['coefs'], emu['intercepts']
    n_net, n_lay = len(coefs), len(coefs[0])
    n_hid = n_lay - 1
    if not 1 <= n_hid <= TC_MAX_HID:
        return None
    sizes = [coefs[0][0].shape[0]] + [w.shape[1] for w in coefs[0]]
    d = sizes[0]
    # Biases ride on the tensor cores: the input row carries a constant 1 in
    # its first padding column, every hidden layer regenerates that 1 in ITS
    # first padding column (a weight of exactly 1 from the previous constant
    # column; ReLU(1) = 1), and the bias vector is the weight row of that
    # column.  Hence fan_in is padded to hold one extra column.
    k0p = _round_up(d + 1, 8)
    np_ = [_round_up(sizes[l + 1] + 1, 16) for l in range(n_hid)]
    kp = [k0p] + [_round_up(sizes[l + 1] + 1, 8) for l in range(n_hid - 1)]
    # TMEM columns: input row, then one region per hidden layer (its
    # accumulator, rewritten in place as the next layer's A operand)
    a0_col, col = 0, _round_up(k0p, 32)
    d_col = []
    for l in range(n_hid):
        d_col.append(col)
        col += _round_up(np_[l], 32)
    w_off, b_off, off = [], [], 0
    for l in range(n_hid):
        w_off.append(off)
        off += np_[l] * kp[l]
    for l in range(n_hid):
        b_off.append(off)          # kept for layout stability; unused
        off += 0
    w_out_off = off
    off += np_[-1]
    b_out_off = off
    off += 4
    net_stride = _round_up(off, 4)
    total = net_stride * n_net
    if col <= TC_COLS_SINGLE and total * 4 <= TC_SMEM_LIMIT:
        # resident mode: every weight of every network stays in shared memory
        n_groups = 2 if col <= TC_COLS else 1
    else:
        # streamed mode (csrc/nb200_mlp_layer.cu): one launch per (network,
        # layer); a layer's weights and its TMEM footprint must fit
        n_groups = 0
        for l in range(n_hid):
            extra = np_[-1] + 4 if l == n_hid - 1 else 0
            cols = _round_up(kp[l], 32) + _round_up(np_[l], 32)
            if (np_[l] * kp[l] + extra) * 4 > TC_SMEM_LIMIT or \
                    cols > TC_COLS_SINGLE:
                return None
    blob = np.zeros(total, dtype=np.float32)
    for n in range(n_net):
        base = n * net_stride
        for l in range(n_hid):
            fi, fo = sizes[l], sizes[l + 1]
            wt = np.zeros((fo + 1, fi + 1))
            wt[:fo, :fi] = np.asarray(coefs[n][l], dtype=np.float64).T
            wt[:fo, fi] = np.asarray(intercepts[n][l], dtype=np.float64)
            if l + 1 < n_hid:
                wt[fo, fi] = 1.0   # regenerate the constant-one column
            blob[base + w_off[l]:base + w_off[l] + np_[l] * kp[l]] = \
                _core_matrix_layout(tf32_round(wt), np_[l], kp[l])
        w_out = np.asarray(coefs[n][-1], dtype=np.float32).ravel()
        blob[base + w_out_off:base + w_out_off + len(w_out)] = w_out
        blob[base + b_out_off] = np.float32(intercepts[n][-1][0])
    thr = np.array([float(score_predict_min) - 1e-9]).view(np.int32)
    pad4 = lambda v: list(v) + [0] * (TC_MAX_HID - len(v))  # noqa: E731
    hdr = [TC_MAGIC | (n_groups << 16), n_net, n_hid, d, k0p, net_stride,
           total, a0_col]
    hdr += pad4(np_) + pad4(kp) + pad4(w_off) + pad4(b_off) + pad4(d_col)
    hdr += [w_out_off, b_out_off, int(thr[0]), int(thr[1])]
    assert len(hdr) == TC_HDR_WORDS
    return np.asarray(hdr, dtype=np.int32), blob


TC16_MAGIC = 0x7F16


def f16_round(a):
    """Round to IEEE half precision (nearest even), returned as float32."""
    return np.asarray(a, dtype=np.float64).astype(np.float16).astype(
        np.float32)


def _core_matrix_layout16(wt, np_, kp):
    """K-major, no-swizzle UMMA operand layout of a [n, k] half-precision
    matrix zero-padded to [np_, kp]: 8-row x 16-byte (8 element) core
    matrices, K chunks 128 B apart, 8-row groups kp*16 B apart."""
    full = np.zeros((np_, kp), dtype=np.float16)
    full[:wt.shape[0], :wt.shape[1]] = wt
    return full.reshape(np_ // 8, 8, kp // 8, 8).transpose(0, 2, 1, 3).ravel()


def pack_tc16(emu, score_predict_min):
    """Tensor-core blob of an emulator for the kind::f16 form of the emulator
    kernel (fp16 x fp16 -> fp32: the same 10 explicit mantissa bits as tf32,
    half the shared memory, TMEM traffic and MMA time).

    Returns ``(header int32[32], blob float32[] holding the raw bytes)`` or
    ``None`` when the architecture does not fit (resident weights only: <= 4
    hidden layers, <= 220 KB, <= 512 TMEM columns).  Same header fields as
    :func:`pack_tc`; offsets count 4-byte words; ``b_off[0]`` is filled by
    the caller with the blob's offset in the data array."""
    coefs, intercepts = emu['coefs'], emu['intercepts']
    n_net, n_lay = len(coefs), len(coefs[0])
    n_hid = n_lay - 1
    if not 1 <= n_hid <= TC_MAX_HID:
        return None
    sizes = [coefs[0][0].shape[0]] + [w.shape[1] for w in coefs[0]]
    d = sizes[0]
    k0p = _round_up(d + 1, 16)
    np_ = [_round_up(sizes[l + 1] + 1, 16) for l in range(n_hid)]
    kp = [k0p] + np_[:-1]          # fan_in of layer l = padded fan_out of l-1
    a0_col, col = 0, _round_up(k0p // 2, 32)
    d_col = []
    for l in range(n_hid):
        d_col.append(col)
        col += _round_up(np_[l], 32)
    w_off, off = [], 0
    for l in range(n_hid):
        w_off.append(off)
        off += np_[l] * kp[l] // 2             # 4-byte words
    w_out_off = off
    off += np_[-1]
    b_out_off = off
    off += 4
    net_stride = _round_up(off, 4)
    total = net_stride * n_net
    if col > TC_COLS_SINGLE or total * 4 > TC_SMEM_LIMIT:
        return None
    n_groups = 2 if col <= TC_COLS else 1
    blob = np.zeros(total, dtype=np.float32)
    raw = blob.view(np.float16)
    for n in range(n_net):
        base = n * net_stride
        for l in range(n_hid):
            fi, fo = sizes[l], sizes[l + 1]
            wt = np.zeros((fo + 1, fi + 1))
            wt[:fo, :fi] = np.asarray(coefs[n][l], dtype=np.float64).T
            wt[:fo, fi] = np.asarray(intercepts[n][l], dtype=np.float64)
            if l + 1 < n_hid:
                wt[fo, fi] = 1.0   # regenerate the constant-one column
            lo = 2 * (base + w_off[l])
            raw[lo:lo + np_[l] * kp[l]] = _core_matrix_layout16(
                wt.astype(np.float16), np_[l], kp[l])
        w_out = np.asarray(coefs[n][-1], dtype=np.float32).ravel()
        blob[base + w_out_off:base + w_out_off + len(w_out)] = w_out
        blob[base + b_out_off] = np.float32(intercepts[n][-1][0])
    thr = np.array([float(score_predict_min) - 1e-9]).view(np.int32)
    pad4 = lambda v: list(v) + [0] * (TC_MAX_HID - len(v))  # noqa: E731
    hdr = [TC16_MAGIC | (n_groups << 16), n_net, n_hid, d, k0p, net_stride,
           total, a0_col]
    hdr += pad4(np_) + pad4(kp) + pad4(w_off) + pad4([]) + pad4(d_col)
    hdr += [w_out_off, b_out_off, int(thr[0]), int(thr[1])]
    assert len(hdr) == TC_HDR_WORDS
    return np.asarray(hdr, dtype=np.int32), blob


def _is_lower(m):
    return bool(np.all(np.triu(m, 1) == 0))


def _pack_ell(data, ell):
    off_c = data.add(ell['c'])
    off_b = data.add(ell['B'])
    off_binv = data.add(ell['B_inv'])
    return off_c, off_b, off_binv, int(_is_lower(ell['B_inv']))


def union_cdf(log_v_all):
    """CDF of the volume-proportional ellipsoid choice
    (p from nautilus/bounds/union.py:308); last entry forced to 1."""
    log_v_all = np.asarray(log_v_all, dtype=float)
    m = np.max(log_v_all)
    p = np.exp(log_v_all - (m + np.log(np.sum(np.exp(log_v_all - m)))))
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    return cdf


def pack_record(spec, data, cdf=None):
    """Serialise one bound; returns its int32 record (offsets into ``data``
    are absolute, offsets into ``meta`` are relative to the record start)."""
    d = int(spec['n_dim'])
    if d > D_MAX:
        raise ValueError('n_dim={} exceeds the supported maximum {}.'.format(
            d, D_MAX))
    if spec['kind'] == 'cube':
        rec = np.zeros(HDR, dtype=np.int32)
        rec[0], rec[1], rec[2] = HDR, 0, d
        return rec
    K, J = len(spec['mixtures']), len(spec['neural'])
    hdr = np.zeros(HDR, dtype=np.int32)
    hdr[1:6] = (1, d, K, J, int(bool(spec['unit'])))
    if cdf is None:
        cdf = union_cdf(spec['log_v_all'])
    hdr[6] = data.add(cdf)
    tail = []          # variable-length int tables appended after records
    base_tail = HDR + K * MIX_REC + J * NB_REC
    mix_recs = np.zeros((K, MIX_REC), dtype=np.int32)
    for k, mix in enumerate(spec['mixtures']):
        dim_cube = np.asarray(mix['dim_cube'], dtype=bool)
        idx_ell = np.flatnonzero(~dim_cube)
        idx_cube = np.flatnonzero(dim_cube)
        if mix['ell'] is None and len(idx_ell) > 0:
            raise ValueError('mixture without ellipsoid must be all-cube')
        off_idx = base_tail + sum(len(t) for t in tail)
        tail.append(np.concatenate([idx_ell, idx_cube]).astype(np.int32))
        if mix['ell'] is not None:
            off_c, off_b, off_binv, tri = _pack_ell(data, mix['ell'])
        else:
            off_c = off_b = off_binv = -1
            tri = 1
        mix_recs[k] = (len(idx_ell), len(idx_cube), off_idx, off_c, off_b,
                       off_binv, tri, 0)
    nb_recs = np.zeros((J, NB_REC), dtype=np.int32)
    max_width = 1
    for j, nb in enumerate(spec['neural']):
        off_c, _, off_binv, tri = _pack_ell(data, nb['ell'])
        emu = nb['emulator']
        if emu is None:
            nb_recs[j] = (off_c, off_binv, tri, 0, 0, -1, -1, -1, -1, -1, -1,
                          0)
            continue
        n_net, n_lay = len(emu['coefs']), len(emu['coefs'][0])
        sizes = [emu['coefs'][0][0].shape[0]] + [
            w.shape[1] for w in emu['coefs'][0]]
        if sizes[0] != d or sizes[-1] != 1:
            raise ValueError('emulator layer sizes {} do not map {} -> 1'
                             .format(sizes, d))
        if max(sizes) > W_MAX:
            raise ValueError('hidden width above {}'.format(W_MAX))
        max_width = max(max_width, max(sizes))
        off_mean = data.add(emu['mean'])
        off_scale = data.add(emu['scale'])
        off_thr = data.add([float(nb['score_predict_min']) - 1e-9,
                            float(nb['score_predict_min'])])
        off_sizes = base_tail + sum(len(t) for t in tail)
        tail.append(np.asarray(sizes, dtype=np.int32))
        wtab = []
        for ws, bs in zip(emu['coefs'], emu['intercepts']):
            if [w.shape for w in ws] != [
                    (sizes[i], sizes[i + 1]) for i in range(n_lay)]:
                raise ValueError('all networks must share one architecture')
            for w, b in zip(ws, bs):
                wtab += [data.add(w), data.add(b)]
        off_wtab = base_tail + sum(len(t) for t in tail)
        tail.append(np.asarray(wtab, dtype=np.int32))
        off_tc, off_tc_hdr = -1, 0
        tc = pack_tc(emu, nb['score_predict_min'])
        if tc is not None:
            off_tc = data.add(np.concatenate(
                [tc[1], np.zeros(len(tc[1]) % 2, np.float32)]).view(
                    np.float64))
            off_tc_hdr = base_tail + sum(len(t) for t in tail)
            tail.append(tc[0])
            # the kind::f16 form of the same weights: a second 32-int header
            # right behind the first (all zero when the architecture does not
            # fit), its b_off[0] = offset of the blob in the data array
            tc16 = pack_tc16(emu, nb['score_predict_min'])
            if tc16 is None:
                tail.append(np.zeros(TC_HDR_WORDS, dtype=np.int32))
            else:
                h16 = tc16[0].copy()
                h16[20] = data.add(np.concatenate(
                    [tc16[1], np.zeros(len(tc16[1]) % 2, np.float32)]).view(
                        np.float64))
                tail.append(h16)
        nb_recs[j] = (off_c, off_binv, tri, n_net, n_lay, off_mean, off_scale,
                      off_thr, off_sizes, off_wtab, off_tc, off_tc_hdr)
    hdr[7] = HDR
    hdr[8] = HDR + K * MIX_REC
    hdr[9] = max_width
    # the usual unimodal bound: the neural bound's ellipsoid IS mixture k's
    # (both are the MVEE of the same live points) -> one whitening serves both
    if J == 1:
        nbe = spec['neural'][0]['ell']
        for k, mix in enumerate(spec['mixtures']):
            e = mix['ell']
            if (e is not None and not np.any(mix['dim_cube']) and
                    np.array_equal(e['c'], nbe['c']) and
                    np.array_equal(e['B_inv'], nbe['B_inv'])):
                hdr[10] = k + 1
                # rounding amplification of x = c + B z -> B_inv (x - c):
                # the fused front end takes t = s z for the whitened point
                # unless the squared radius is within a guard band of 1
                # whose width follows from this number (nb200_front_mma.cu)
                amp = np.linalg.norm(e['B_inv'], 2) * (
                    np.linalg.norm(e['B'], 2) + np.max(np.abs(e['c'])))
                hdr[11] = int(np.clip(np.ceil(np.log2(max(amp, 1.0))), 0, 62)
                              ) if np.isfinite(amp) else 62
                break
    rec = np.concatenate([hdr, mix_recs.ravel(), nb_recs.ravel()] + tail)
    rec = rec.astype(np.int32)
    rec[0] = len(rec)
    return rec


def pack_stack(specs):
    """Serialise a list of bounds.

    Returns ``(meta int32[], data float64[])``; ``meta[0]`` = number of
    bounds L, ``meta[1 + i]`` = start of record i.
    """
    data = _Data()
    recs = [pack_record(s, data) for s in specs]
    L = len(recs)
    table = np.zeros(1 + L, dtype=np.int32)
    table[0] = L
    off = 1 + L
    for i, r in enumerate(recs):
        table[1 + i] = off
        off += len(r)
    meta = np.concatenate([table] + recs).astype(np.int32)
    return meta, data.array()
