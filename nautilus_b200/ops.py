"""Thin tensor-level wrappers over the C ABI (include/nautilus_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every call below
passes raw device pointers and the current CUDA stream to
``libnautilus_b200.so``.  All tensors must be CUDA, contiguous, float64 for
points / uint8 for masks.
"""

import ctypes

import numpy as np
import torch

from . import _lib
from ._pack import _Data, pack_record, pack_stack

MLP_F64 = 0
MLP_TF32 = 1
MLP_F16 = 2

LIKE_GAUSSIAN = 0
LIKE_ROSENBROCK = 1
LIKE_MIXTURE = 2
LIKE_EQUICORR = 3

CODE_CUBE_REJECT, CODE_OVERLAP_REJECT, CODE_NN_REJECT, CODE_EXCLUDED, \
    CODE_IN_SHELL = range(5)
CNT_RAW, CNT_CUBE_REJECT, CNT_OVERLAP_REJECT, CNT_NN_REJECT, CNT_EXCLUDED, \
    CNT_IN_SHELL, CNT_UPDATE = range(7)
N_CNT = 8
N_LSE = 4


def _ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_points(points, d=None):
    if not (points.is_cuda and points.dtype == torch.float64 and
            points.dim() == 2 and points.is_contiguous()):
        raise ValueError('points must be a contiguous CUDA float64 [n, d] '
                         'tensor')
    if d is not None and points.shape[1] != d:
        raise ValueError('points have {} dimensions, bound has {}'.format(
            points.shape[1], d))


def _chk_mask(mask, n):
    if mask is None:
        return None
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)
    if not (mask.is_cuda and mask.dtype == torch.uint8 and
            mask.is_contiguous() and mask.numel() == n):
        raise ValueError('mask must be a contiguous CUDA uint8/bool [n] tensor')
    return mask


def device_info():
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib().nb200_device_info(
        ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)))
    return sm.value, major.value, minor.value


def launch_count():
    return int(_lib.lib().nb200_launch_count())


N_STAGES = 10


def profile_enable(on=True):
    _lib.lib().nb200_profile_enable(int(bool(on)))


def profile_collect():
    """{stage name: (milliseconds, scopes)} accumulated since enable()."""
    ms = np.zeros(N_STAGES)
    calls = np.zeros(N_STAGES, dtype=np.int64)
    _lib.check(_lib.lib().nb200_profile_collect(
        ms.ctypes.data_as(ctypes.c_void_p),
        calls.ctypes.data_as(ctypes.c_void_p)))
    return {_lib.lib().nb200_profile_stage_name(i).decode():
            (float(ms[i]), int(calls[i])) for i in range(N_STAGES)}


class Workspace:
    """Grow-only scratch buffer handed to the library.

    The buffer is shared by every stack on the same (device, stream): a call
    of the library uses it from start to end of its own launches only, and
    launches on one stream are ordered.  (One buffer per stack meant a
    cudaMalloc of up to a few hundred MB for every new bound -- and a
    device-synchronising cudaFree when the stack died: ~1 s per config-2 run.)
    """

    _shared = {}

    def __init__(self, device):
        self.device = device

    def get(self, n, d, pairs=0):
        if pairs > 0:
            need = int(_lib.lib().nb200_cycle_workspace_bytes(
                int(n), int(d), int(pairs)))
        else:
            need = int(_lib.lib().nb200_workspace_bytes(int(n), int(d)))
        stream = torch.cuda.current_stream(self.device)
        key = (self.device.index, stream.cuda_stream)
        buf = Workspace._shared.get(key)
        if buf is None or buf.numel() < need:
            # grow by at least half so that a slowly rising demand does not
            # reallocate on every call
            have = 0 if buf is None else buf.numel()
            buf = None
            Workspace._shared.pop(key, None)
            buf = torch.empty(max(need, have + have // 2), dtype=torch.uint8,
                              device=self.device)
            Workspace._shared[key] = buf
        return buf, need


class DeviceStack:
    """A serialised list of bounds resident on one GPU."""

    def __init__(self, specs, device='cuda'):
        self.device = torch.device(device)
        self.n_dim = int(specs[0]['n_dim'])
        self._data = _Data()
        self._recs = []
        self._n_chunks = 0          # data chunks already on the device
        self._n_data = 0            # doubles already on the device
        self.data_d = None
        self.ws = Workspace(self.device)
        for spec in specs:
            self._recs.append(pack_record(spec, self._data))
        self._sync()

    def append(self, spec):
        """Add one more bound at the end of the stack (a new bound was
        accepted, sampler.py:1023-1039): only ITS parameters are serialised
        and uploaded, the records already on the device stay where they are
        (data offsets are absolute and do not move)."""
        self._recs.append(pack_record(spec, self._data))
        self._sync()

    def _sync(self):
        # new data chunks -> the tail of the device array (grow by doubling)
        new = self._data.chunks[self._n_chunks:]
        total = self._data.n
        if self.data_d is None or self.data_d.numel() < max(total, 2):
            grown = torch.empty(max(2 * total, 1024), dtype=torch.float64,
                                device=self.device)
            if self.data_d is not None and self._n_data:
                grown[:self._n_data] = self.data_d[:self._n_data]
            self.data_d = grown
        if new:
            chunk = torch.from_numpy(np.concatenate(new))
            self.data_d[self._n_data:self._n_data + chunk.numel()] = chunk.to(
                self.device)
        self._n_chunks = len(self._data.chunks)
        self._n_data = total
        L = len(self._recs)
        table = np.zeros(1 + L, dtype=np.int32)
        table[0] = L
        off = 1 + L
        for i, r in enumerate(self._recs):
            table[1 + i] = off
            off += len(r)
        self.meta_h = np.ascontiguousarray(
            np.concatenate([table] + self._recs), dtype=np.int32)
        self.n_bounds = L
        self.meta_d = torch.from_numpy(self.meta_h).to(self.device)

    @property
    def _meta_h_ptr(self):
        return self.meta_h.ctypes.data_as(ctypes.c_void_p)

    # -- Union ------------------------------------------------------------
    def union_count(self, bound, points, mask=None):
        """(n_bound int32[n], contains bool[n]) -- union.py:285-289,316-317."""
        _chk_points(points, self.n_dim)
        n = points.shape[0]
        mask = _chk_mask(mask, n)
        count = torch.empty(n, dtype=torch.int32, device=self.device)
        contains = torch.empty(n, dtype=torch.uint8, device=self.device)
        _lib.check(_lib.lib().nb200_union_count(
            self._meta_h_ptr, _ptr(self.meta_d), _ptr(self.data_d), bound,
            _ptr(points), _ptr(mask), n, _ptr(count), _ptr(contains),
            _stream()))
        return count, contains.bool()

    def propose(self, bound, n, seed=0, offset=0, stream_id=0, test=None):
        """Raw draws + cube filter + overlap acceptance (union.py:305-319).

        ``test`` = dict(k, z, cube_u, u, r) of CUDA tensors switches to the
        host-supplied randoms of the reference (test mode).
        Returns (points f64[n,d], code u8[n], n_bound i32[n]).
        """
        points = torch.empty((n, self.n_dim), dtype=torch.float64,
                             device=self.device)
        code = torch.empty(n, dtype=torch.uint8, device=self.device)
        n_bound = torch.empty(n, dtype=torch.int32, device=self.device)
        t = test or {}
        _lib.check(_lib.lib().nb200_union_propose(
            self._meta_h_ptr, _ptr(self.meta_d), _ptr(self.data_d), bound, n,
            seed, offset, stream_id, _ptr(t.get('k')), _ptr(t.get('z')),
            _ptr(t.get('cube_u')), _ptr(t.get('u')), _ptr(t.get('r')),
            _ptr(points), _ptr(code), _ptr(n_bound), _stream()))
        return points, code, n_bound

    # -- emulator ---------------------------------------------------------
    def mlp_predict(self, bound, j, x, mode=MLP_F64):
        """emulator.predict on whitened coordinates (neural.py:114-116)."""
        _chk_points(x, self.n_dim)
        n = x.shape[0]
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        buf, nbytes = self.ws.get(n, self.n_dim)
        _lib.check(_lib.lib().nb200_mlp_predict(
            self._meta_h_ptr, _ptr(self.meta_d), _ptr(self.data_d), bound, j,
            _ptr(x), n, _ptr(out), mode, _ptr(buf), nbytes, _stream()))
        return out

    # -- contains ---------------------------------------------------------
    def contains(self, bound, points, which=0, mask=None, mode=MLP_F64):
        """NautilusBound.contains (nautilus.py:146-169); which=1 union only,
        which=2 the neural filter of NautilusBound.sample (:217-218)."""
        _chk_points(points, self.n_dim)
        n = points.shape[0]
        mask = _chk_mask(mask, n)
        out = torch.empty(n, dtype=torch.uint8, device=self.device)
        buf, nbytes = self.ws.get(n, self.n_dim)
        _lib.check(_lib.lib().nb200_bound_contains(
            self._meta_h_ptr, _ptr(self.meta_d), _ptr(self.data_d), bound,
            which, _ptr(points), _ptr(mask), n, _ptr(out), mode, _ptr(buf),
            nbytes, _stream()))
        return out.bool()

    # -- full cycle ---------------------------------------------------------
    def cycle(self, bound, n, later=(0, 0), seed=0, offset=0, stream_id=0,
              like_id=-1, like_params=None, log_l_min=-np.inf, mode=MLP_F64,
              out=None):
        """One raw batch through propose -> NN filter -> exclusion ->
        likelihood -> sums (sampler.py:1093-1144).

        Returns dict(points, log_l, code, lse f64[4], counters i64[8]) of CUDA
        tensors; nothing is synchronised.
        """
        d = self.n_dim
        if out is None:
            out = dict(
                points=torch.empty((n, d), dtype=torch.float64,
                                   device=self.device),
                log_l=torch.empty(n, dtype=torch.float64, device=self.device),
                code=torch.empty(n, dtype=torch.uint8, device=self.device),
                lse=torch.empty(N_LSE, dtype=torch.float64,
                                device=self.device),
                counters=torch.empty(N_CNT, dtype=torch.int64,
                                     device=self.device))
        # (later bound, neural bound) pairs of the exclusion: record header
        # word 4 is J
        pairs = sum(int(self.meta_h[self.meta_h[1 + b] + 4])
                    for b in range(later[0], later[0] + later[1])) \
            if later[1] > 0 and mode in (MLP_TF32, MLP_F16) else 0
        buf, nbytes = self.ws.get(n, d, pairs)
        n_par = 0 if like_params is None else like_params.numel()
        _lib.check(_lib.lib().nb200_cycle(
            self._meta_h_ptr, _ptr(self.meta_d), _ptr(self.data_d), bound,
            later[0], later[1], n, seed, offset, stream_id, like_id,
            _ptr(like_params), n_par, float(log_l_min), mode,
            _ptr(out['points']), _ptr(out['log_l']), _ptr(out['code']),
            _ptr(out['lse']), _ptr(out['counters']), _ptr(buf), nbytes,
            _stream()))
        return out

    def compact(self, points, log_l, code, out_points=None, out_log_l=None):
        """Stable gather of the in-shell rows.  Returns (points, log_l, n)
        with n a CUDA int64 scalar tensor (no sync)."""
        n, d = points.shape
        if out_points is None:
            out_points = torch.empty_like(points)
        if out_log_l is None and log_l is not None:
            out_log_l = torch.empty_like(log_l)
        n_out = torch.empty(1, dtype=torch.int64, device=self.device)
        buf, nbytes = self.ws.get(n, d)
        _lib.check(_lib.lib().nb200_compact(
            _ptr(points), _ptr(log_l), _ptr(code), n, d, _ptr(out_points),
            _ptr(out_log_l), _ptr(n_out), _ptr(buf), nbytes, _stream()))
        return out_points, out_log_l, n_out


    def compact_index(self, log_l, code, offset=0):
        """(global index u64[n] as int64 tensor, log_l f64[n], n_out) of the
        in-shell rows (nb200_compact_index); nothing is synchronised and only
        the first n_out entries are meaningful."""
        n = code.numel()
        index = torch.empty(n, dtype=torch.int64, device=self.device)
        out_ll = None if log_l is None else torch.empty_like(log_l)
        n_out = torch.empty(1, dtype=torch.int64, device=self.device)
        buf, nbytes = self.ws.get(n, self.n_dim)
        _lib.check(_lib.lib().nb200_compact_index(
            _ptr(log_l), _ptr(code), n, int(offset), _ptr(index),
            _ptr(out_ll), _ptr(n_out), _ptr(buf), nbytes, _stream()))
        return index, out_ll, n_out

    def materialize(self, bound, index, seed=0, stream_id=0, mode=MLP_F64):
        """Rows f64[k,d] of the proposals with global indices ``index``
        (int64/uint64 CUDA tensor), bit-identical to what ``cycle`` wrote for
        them with the same ``mode`` (nb200_materialize)."""
        if not (index.is_cuda and index.is_contiguous() and
                index.dtype in (torch.int64, torch.uint64)):
            raise ValueError('index must be a contiguous CUDA int64 tensor')
        k = index.numel()
        out = torch.empty((k, self.n_dim), dtype=torch.float64,
                          device=self.device)
        _lib.check(_lib.lib().nb200_materialize(
            self._meta_h_ptr, _ptr(self.meta_d), _ptr(self.data_d), bound,
            int(seed), int(stream_id), int(mode), _ptr(index), k, _ptr(out),
            _stream()))
        return out


RETURN_ROWS = 0
RETURN_INDEX = 1


class HostSession:
    """Host-buffer form of the cycle (``nb200_session_*``): NumPy in, NumPy
    out, no torch tensors.  ``submit`` enqueues one raw batch and returns at
    once; ``wait`` returns what ``Sampler.add_samples`` appends to its host
    arrays (sampler.py:1135-1141) and what ``update_shell_info`` reduces
    (sampler.py:925-943).  With ``n_slots >= 2`` the device->host copy of one
    batch runs under the kernels of the next.

    The arrays returned by ``wait`` are views of the slot's pinned buffers:
    valid until that slot is submitted again (copy them to keep them).
    """

    def __init__(self, specs, n_max, cap=None, n_slots=2, like_params_max=0,
                 device=None, returns='rows'):
        if returns not in ('rows', 'index'):
            raise ValueError("returns must be 'rows' or 'index'")
        if device is not None:
            torch.cuda.set_device(device)
        meta, data = pack_stack(specs)
        self.meta_h = np.ascontiguousarray(meta, dtype=np.int32)
        self.data_h = np.ascontiguousarray(data, dtype=np.float64)
        self.n_dim = int(specs[0]['n_dim'])
        self.n_max = int(n_max)
        self.cap = int(cap if cap is not None else n_max)
        self.n_slots = int(n_slots)
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().nb200_session_create(
            self.meta_h.ctypes.data_as(ctypes.c_void_p), self.meta_h.size,
            self.data_h.ctypes.data_as(ctypes.c_void_p), self.data_h.size,
            self.n_max, self.cap, self.n_slots, int(like_params_max),
            ctypes.byref(self._h)))
        self.returns = returns
        if returns == 'index':
            _lib.check(_lib.lib().nb200_session_set_returns(
                self._h, RETURN_INDEX))

    @property
    def stack_bytes(self):
        return self.meta_h.nbytes + self.data_h.nbytes

    def set_stack(self, specs):
        meta, data = pack_stack(specs)
        self.meta_h = np.ascontiguousarray(meta, dtype=np.int32)
        self.data_h = np.ascontiguousarray(data, dtype=np.float64)
        _lib.check(_lib.lib().nb200_session_set_stack(
            self._h, self.meta_h.ctypes.data_as(ctypes.c_void_p),
            self.meta_h.size, self.data_h.ctypes.data_as(ctypes.c_void_p),
            self.data_h.size))

    def submit(self, slot, bound, n, later=(0, 0), seed=0, offset=0,
               stream_id=0, like_id=-1, like_params=None, log_l_min=-np.inf,
               mode=MLP_F64, upload_stack=False):
        if like_params is None:
            par, n_par = None, 0
        else:
            like_params = np.ascontiguousarray(like_params, dtype=np.float64)
            par = like_params.ctypes.data_as(ctypes.c_void_p)
            n_par = like_params.size
        _lib.check(_lib.lib().nb200_session_submit(
            self._h, int(slot), int(bool(upload_stack)), int(bound),
            int(later[0]), int(later[1]), int(n), int(seed), int(offset),
            int(stream_id), int(like_id), par, n_par, float(log_l_min),
            int(mode)))

    def wait(self, slot):
        """dict(points f64[k,d], log_l f64[k] | None, lse f64[4],
        counters i64[8]) for the batch submitted on ``slot``; in index mode
        ``index`` u64[k] (global proposal indices) replaces ``points``."""
        p_pts, p_ll = ctypes.c_void_p(), ctypes.c_void_p()
        k = ctypes.c_int64()
        lse = np.empty(N_LSE)
        counters = np.empty(N_CNT, dtype=np.int64)
        fn = (_lib.lib().nb200_session_wait_index if self.returns == 'index'
              else _lib.lib().nb200_session_wait)
        _lib.check(fn(
            self._h, int(slot), ctypes.byref(p_pts), ctypes.byref(p_ll),
            ctypes.byref(k), lse.ctypes.data_as(ctypes.c_void_p),
            counters.ctypes.data_as(ctypes.c_void_p)))
        k = int(k.value)
        d = self.n_dim

        def view(ptr, shape, dtype=np.float64):
            if not ptr.value or k == 0:
                return np.empty(shape, dtype=dtype)
            buf = (ctypes.c_double * int(np.prod(shape))).from_address(
                ptr.value)
            return np.frombuffer(buf, dtype=dtype).reshape(shape)

        out = dict(log_l=view(p_ll, (k,)) if p_ll.value else None,
                   lse=lse, counters=counters)
        if self.returns == 'index':
            out['index'] = view(p_pts, (k,), np.uint64)
        else:
            out['points'] = view(p_pts, (k, d))
        return out

    def materialize(self, bound, index, seed=0, stream_id=0, mode=MLP_F64):
        """Rows f64[k,d] (NumPy) of the proposals with global indices
        ``index`` of stack record ``bound``: bit-identical to the rows the
        cycle produced for them (nb200_session_materialize)."""
        index = np.ascontiguousarray(index, dtype=np.uint64)
        out = np.empty((index.size, self.n_dim))
        _lib.check(_lib.lib().nb200_session_materialize(
            self._h, int(bound), int(seed), int(stream_id), int(mode),
            index.ctypes.data_as(ctypes.c_void_p), index.size,
            out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def close(self):
        if self._h:
            _lib.lib().nb200_session_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# -- single-ellipsoid primitives -------------------------------------------

def ell_transform(points, c, M, inverse=False):
    """Ellipsoid.transform (basic.py:339-342)."""
    _chk_points(points)
    n, d = points.shape
    out = torch.empty_like(points)
    _lib.check(_lib.lib().nb200_ell_transform(
        _ptr(points), n, d, _ptr(c), _ptr(M), int(inverse), _ptr(out),
        _stream()))
    return out


def ell_contains(points, c, B_inv, return_r2=False):
    """Ellipsoid.contains (basic.py:360)."""
    _chk_points(points)
    n, d = points.shape
    out = torch.empty(n, dtype=torch.uint8, device=points.device)
    r2 = torch.empty(n, dtype=torch.float64, device=points.device) \
        if return_r2 else None
    _lib.check(_lib.lib().nb200_ell_contains(
        _ptr(points), n, d, _ptr(c), _ptr(B_inv), _ptr(out), _ptr(r2),
        _stream()))
    return (out.bool(), r2) if return_r2 else out.bool()


def ell_sample_from(z, u, c, B):
    """Ellipsoid.sample with explicit base randoms (basic.py:376-381)."""
    _chk_points(z)
    n, d = z.shape
    out = torch.empty_like(z)
    _lib.check(_lib.lib().nb200_ell_sample_from(
        _ptr(z), _ptr(u), n, d, _ptr(c), _ptr(B), _ptr(out), _stream()))
    return out


def stats(log_l, code=None, log_l_min=-np.inf):
    """(lse f64[4], counters i64[8]) -- sampler.py:934-937, 1144."""
    n = log_l.numel()
    dev = log_l.device
    lse = torch.empty(N_LSE, dtype=torch.float64, device=dev)
    counters = torch.empty(N_CNT, dtype=torch.int64, device=dev)
    nbytes = int(_lib.lib().nb200_workspace_bytes(1, 1))
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib().nb200_stats(
        _ptr(log_l), _ptr(code), n, float(log_l_min), _ptr(lse),
        _ptr(counters), _ptr(buf), nbytes, _stream()))
    return lse, counters


def loglike(points, like_id, params, code=None):
    _chk_points(points)
    n, d = points.shape
    out = torch.empty(n, dtype=torch.float64, device=points.device)
    _lib.check(_lib.lib().nb200_loglike(
        _ptr(points), _ptr(code), n, d, like_id, _ptr(params),
        params.numel(), _ptr(out), _stream()))
    return out


def kth_largest(values, k, mask=None):
    """(threshold f64[1], greater i64[1]) CUDA tensors: the k-th largest entry
    of the CUDA float64 vector ``values`` (among ``mask != 0`` if given) and
    how many are strictly greater (nb200_select_kth_largest: device radix
    select, no sort, no synchronisation)."""
    n = values.numel()
    if not (values.is_cuda and values.dtype == torch.float64 and
            values.is_contiguous()):
        raise ValueError('values must be a contiguous CUDA float64 tensor')
    mask = _chk_mask(mask, n)
    dev = values.device
    thr = torch.empty(1, dtype=torch.float64, device=dev)
    greater = torch.empty(1, dtype=torch.int64, device=dev)
    nbytes = int(_lib.lib().nb200_select_workspace_bytes())
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib().nb200_select_kth_largest(
        _ptr(values), _ptr(mask), n, int(k), _ptr(thr), _ptr(greater),
        _ptr(buf), nbytes, _stream()))
    return thr, greater


def top_k(values, k):
    """Indices (int64 CUDA tensor, unordered) of the k largest entries of the
    CUDA float64 vector ``values`` -- the live set of sampler.py:1007-1009,
    1160-1164 without a host argsort: device radix select of the k-th largest
    value, everything above it, and the first ties in index order to fill up
    (``argsort`` picks ties arbitrarily too)."""
    n = values.numel()
    if k >= n:
        return torch.arange(n, device=values.device)
    values = values.contiguous()
    thr, greater = kth_largest(values, k)
    above = torch.nonzero(values > thr).squeeze(1)
    short = int(k) - above.numel()
    if short > 0:
        ties = torch.nonzero(values == thr).squeeze(1)[:short]
        above = torch.cat([above, ties])
    return above


def mvee_weights(q_t, max_updates=1500, tol=1e-3):
    """Khachiyan weights of the MVEE of the points ``q_t`` f64[d, n] (CUDA,
    coordinate-major, ideally whitened) -- basic.py:175-241.  Returns
    (u f64[n] CUDA, iterations int32[1] CUDA)."""
    if not (q_t.is_cuda and q_t.dtype == torch.float64 and q_t.dim() == 2 and
            q_t.is_contiguous()):
        raise ValueError('q_t must be a contiguous CUDA float64 [d, n] tensor')
    d, n = q_t.shape
    u = torch.empty(n, dtype=torch.float64, device=q_t.device)
    iters = torch.zeros(1, dtype=torch.int32, device=q_t.device)
    nbytes = int(_lib.lib().nb200_mvee_workspace_bytes(n))
    buf = torch.empty(nbytes, dtype=torch.uint8, device=q_t.device)
    _lib.check(_lib.lib().nb200_mvee_weights(
        _ptr(q_t), n, d, int(max_updates), float(tol), _ptr(u), _ptr(iters),
        _ptr(buf), nbytes, _stream()))
    return u, iters


def gmm2_applicable(n, d):
    """Does ``gmm2_em`` take an [n, d] problem?"""
    return bool(_lib.lib().nb200_gmm2_applicable(int(n), int(d)))


def gmm2_em(x, labels, max_iter=100, tol=1e-3, reg=1e-6):
    """EM of a 2-component Gaussian mixture for all restarts in one launch
    (the GaussianMixture call of Union.split, union.py:185-187).  ``x``
    f64[n, d] CUDA, ``labels`` uint8[R, n] CUDA initial assignments.  Returns
    CUDA tensors (log_p f64[R, 2, n], score f64[R], iters int32[R])."""
    _chk_points(x)
    n, d = x.shape
    if not (labels.is_cuda and labels.dtype == torch.uint8 and
            labels.dim() == 2 and labels.shape[1] == n and
            labels.is_contiguous()):
        raise ValueError('labels must be a contiguous CUDA uint8 [R, n] '
                         'tensor')
    r = labels.shape[0]
    log_p = torch.empty((r, 2, n), dtype=torch.float64, device=x.device)
    score = torch.empty(r, dtype=torch.float64, device=x.device)
    iters = torch.empty(r, dtype=torch.int32, device=x.device)
    _lib.check(_lib.lib().nb200_gmm2_em(
        _ptr(x), n, d, _ptr(labels), r, int(max_iter), float(tol), float(reg),
        _ptr(log_p), _ptr(score), _ptr(iters), _stream()))
    return log_p, score, iters


def mlp_fit(x, y, sizes, n_networks, seed=0, lr=1e-2, beta1=0.9, beta2=0.999,
            eps=1e-8, batch_size=200, max_epochs=10000, tol=0.0, patience=10):
    """Train an ensemble on standardised inputs (neural.py:50-98).

    x f64[m,d], y f64[m] CUDA tensors.  Returns (params f64[n_net, n_params]
    CUDA, n_iter int32[n_net] CUDA, loss f64[n_net] CUDA)."""
    _chk_points(x)
    m, d = x.shape
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    n_lay = len(sizes) - 1
    n_params = int(sum(sizes[i] * sizes[i + 1] + sizes[i + 1]
                       for i in range(n_lay)))
    dev = x.device
    params = torch.empty((n_networks, n_params), dtype=torch.float64,
                         device=dev)
    n_iter = torch.empty(n_networks, dtype=torch.int32, device=dev)
    loss = torch.empty(n_networks, dtype=torch.float64, device=dev)
    nbytes = int(_lib.lib().nb200_mlp_fit_workspace_bytes(
        m, d, n_params, n_networks))
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib().nb200_mlp_fit(
        _ptr(x), _ptr(y.contiguous()), m, d,
        sizes.ctypes.data_as(ctypes.c_void_p), n_lay, n_networks, int(seed),
        lr, beta1, beta2, eps, int(batch_size), int(max_epochs), tol,
        int(patience), _ptr(params), _ptr(n_iter), _ptr(loss), _ptr(buf),
        nbytes, _stream()))
    return params, n_iter, loss
