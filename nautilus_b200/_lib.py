"""Loader of ``libnautilus_b200.so`` (the C ABI in include/nautilus_b200.h).

There is no CPU fallback: if the library is missing or fails to load, every
operation raises.  ``build()`` compiles the CUDA sources in-tree for sm_100a
(nvcc cross-compiles without a GPU).
"""

import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.path.join(_HERE, 'libnautilus_b200.so')

# --exclude-libs,ALL: nvcc links parts of libstdc++.a / libcudart_static.a
# into the library; their symbols must not be exported (a second, partial C++
# runtime in the lookup scope of whoever links against us breaks exception
# unwinding there -- seen with the torch.ops shim)
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-std=c++17', '-shared', '-Xcompiler', '-fPIC', '-Xlinker',
              '--exclude-libs,ALL']

_lib = None


class NautilusB200Error(RuntimeError):
    """Raised when a library call fails (mirrors RuntimeError semantics)."""


def sources():
    return sorted(glob.glob(os.path.join(_CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(_CSRC, '*.cuh')) + [
        os.path.join(_HERE, '..', 'include', 'nautilus_b200.h')]
    return any(os.path.getmtime(p) > t for p in deps)


def host_cxx():
    """The SYSTEM C++ compiler.  Deliberately not $CXX: in this image $CXX is
    a relocated gcc that only has a static libstdc++, and a second,
    uninitialised C++ runtime inside a dlopen'ed library crashes at the first
    ostream / exception (seen: a failed STD_TORCH_CHECK in the torch.ops shim
    segfaulted).  NB200_CXX overrides."""
    cxx = os.environ.get('NB200_CXX')
    if cxx:
        return cxx
    return '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'


def build(force=False, verbose=False):
    """Compile the extension in-tree with nvcc for sm_100a."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    extra = os.environ.get('NB200_EXTRA_FLAGS', '').split()
    cmd = [nvcc, '-ccbin', host_cxx()] + NVCC_FLAGS + extra + [
        '-o', LIB_PATH] + sources() + ['-lcuda']
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


TORCH_SRC = os.path.join(_HERE, 'csrc_torch', 'nb200_torch_ops.cpp')
TORCH_LIB_PATH = os.path.join(_HERE, 'libnautilus_b200_torch.so')
_torch_ops = None


def build_torch_ops(force=False, verbose=False):
    """Compile the STABLE_TORCH_LIBRARY shim (torch.ops.nautilus_b200.*,
    csrc_torch/nb200_torch_ops.cpp) in-tree: host C++ only, linked against
    libtorch's C shim libnautilus_b200.so."""
    build()
    if (not force and os.path.exists(TORCH_LIB_PATH) and
            os.path.getmtime(TORCH_LIB_PATH) >= max(
                os.path.getmtime(TORCH_SRC), os.path.getmtime(LIB_PATH))):
        return TORCH_LIB_PATH
    import torch
    troot = os.path.dirname(torch.__file__)
    inc, lib = os.path.join(troot, 'include'), os.path.join(troot, 'lib')
    cmd = [host_cxx(), '-O2', '-std=c++17', '-shared',
           '-fPIC', '-DUSE_CUDA', '-DTORCH_TARGET_VERSION=0x020B000000000000',
           '-I' + inc, '-I' + os.path.join(inc, 'torch', 'csrc', 'api',
                                           'include'),
           '-I/usr/local/cuda/include', '-o', TORCH_LIB_PATH, TORCH_SRC,
           '-L' + lib, '-ltorch', '-ltorch_cpu', '-ltorch_cuda', '-lc10',
           '-ldl', '-Wl,-rpath,' + lib]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return TORCH_LIB_PATH


def torch_ops():
    """``torch.ops.nautilus_b200`` (loads the shim on first use; raises if
    it is not built -- no fallback)."""
    global _torch_ops
    if _torch_ops is None:
        if not os.path.exists(TORCH_LIB_PATH):
            raise NautilusB200Error(
                'libnautilus_b200_torch.so is not built; run '
                '`python -c "import __graft_entry__ as g; g.build()"`.')
        import torch
        lib()                       # the C ABI it forwards to
        torch.ops.load_library(TORCH_LIB_PATH)
        _torch_ops = torch.ops.nautilus_b200
    return _torch_ops


_i32p = ctypes.c_void_p
_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_u64 = ctypes.c_uint64
_u32 = ctypes.c_uint32
_int = ctypes.c_int
_dbl = ctypes.c_double
_sz = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol of the public header
PROTOTYPES = {
    'nb200_last_error': (ctypes.c_char_p, []),
    'nb200_version': (_int, []),
    'nb200_device_info': (_int, [_vp, _vp, _vp]),
    'nb200_launch_count': (_i64, []),
    'nb200_profile_enable': (None, [_int]),
    'nb200_profile_collect': (_int, [_vp, _vp]),
    'nb200_profile_stage_name': (ctypes.c_char_p, [_int]),
    'nb200_workspace_bytes': (_sz, [_i64, _int]),
    'nb200_cycle_workspace_bytes': (_sz, [_i64, _int, _int]),
    'nb200_ell_transform': (_int, [_vp, _i64, _int, _vp, _vp, _int, _vp, _vp]),
    'nb200_ell_contains': (_int, [_vp, _i64, _int, _vp, _vp, _vp, _vp, _vp]),
    'nb200_ell_sample_from': (_int, [_vp, _vp, _i64, _int, _vp, _vp, _vp,
                                     _vp]),
    'nb200_union_count': (_int, [_i32p, _vp, _vp, _int, _vp, _vp, _i64, _vp,
                                 _vp, _vp]),
    'nb200_union_propose': (_int, [_i32p, _vp, _vp, _int, _i64, _u64, _u64,
                                   _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _vp]),
    'nb200_mlp_predict': (_int, [_i32p, _vp, _vp, _int, _int, _vp, _i64, _vp,
                                 _int, _vp, _sz, _vp]),
    'nb200_mlp_fit_workspace_bytes': (_sz, [_i64, _int, _int, _int]),
    'nb200_mlp_fit': (_int, [_vp, _vp, _i64, _int, _i32p, _int, _int, _u64,
                             _dbl, _dbl, _dbl, _dbl, _int, _int, _dbl, _int,
                             _vp, _vp, _vp, _vp, _sz, _vp]),
    'nb200_bound_contains': (_int, [_i32p, _vp, _vp, _int, _int, _vp, _vp,
                                    _i64, _vp, _int, _vp, _sz, _vp]),
    'nb200_stats': (_int, [_vp, _vp, _i64, _dbl, _vp, _vp, _vp, _sz, _vp]),
    'nb200_loglike': (_int, [_vp, _vp, _i64, _int, _int, _vp, _int, _vp,
                             _vp]),
    'nb200_compact': (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _vp,
                             _sz, _vp]),
    'nb200_compact_index': (_int, [_vp, _vp, _i64, _u64, _vp, _vp, _vp, _vp,
                                   _sz, _vp]),
    'nb200_materialize': (_int, [_i32p, _vp, _vp, _int, _u64, _u32, _int,
                                 _vp, _i64, _vp, _vp]),
    'nb200_cycle': (_int, [_i32p, _vp, _vp, _int, _int, _int, _i64, _u64,
                           _u64, _u32, _int, _vp, _int, _dbl, _int, _vp, _vp,
                           _vp, _vp, _vp, _vp, _sz, _vp]),
    'nb200_cycle_host': (_int, [_i32p, _i64, _vp, _i64, _int, _int, _int,
                                _i64, _u64, _u64, _u32, _int, _vp, _int, _dbl,
                                _int, _i64, _vp, _vp, _vp, _vp, _vp]),
    'nb200_select_workspace_bytes': (_sz, []),
    'nb200_select_kth_largest': (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp,
                                        _sz, _vp]),
    'nb200_mvee_workspace_bytes': (_sz, [_i64]),
    'nb200_mvee_weights': (_int, [_vp, _i64, _int, _int, _dbl, _vp, _vp, _vp,
                                  _sz, _vp]),
    'nb200_gmm2_applicable': (_int, [_i64, _int]),
    'nb200_gmm2_em': (_int, [_vp, _i64, _int, _vp, _int, _int, _dbl, _dbl,
                             _vp, _vp, _vp, _vp]),
    'nb200_session_create': (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _int,
                                    _int, _vp]),
    'nb200_session_destroy': (_int, [_vp]),
    'nb200_session_set_stack': (_int, [_vp, _vp, _i64, _vp, _i64]),
    'nb200_session_submit': (_int, [_vp, _int, _int, _int, _int, _int, _i64,
                                    _u64, _u64, _u32, _int, _vp, _int, _dbl,
                                    _int]),
    'nb200_session_wait': (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp]),
    'nb200_session_set_returns': (_int, [_vp, _int]),
    'nb200_session_wait_index': (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp]),
    'nb200_session_materialize': (_int, [_vp, _int, _u64, _u32, _int, _vp,
                                         _i64, _vp]),
}


def lib():
    """Return the loaded library; raise if it does not exist."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NautilusB200Error(
                'libnautilus_b200.so is not built; run '
                '`python -c "import __graft_entry__ as g; g.build()"`. '
                'nautilus_b200 has no CPU fallback.')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)       # AttributeError if missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise NautilusB200Error(lib().nb200_last_error().decode())
