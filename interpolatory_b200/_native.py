"""ctypes binding of libmemci_b200.so (include/memci_b200.h).

There is no CPU fallback: if the library has not been built, or no sm_100a device is
visible, the first use raises -- loudly -- instead of computing anything on the host.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MEMCI_B200_LIB") or os.path.join(_HERE, "_lib", "libmemci_b200.so")   # env override: A/B builds while tuning
_lib = None


class MemciError(Exception):
    pass


class PairParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("me_mode", "block_size", "region", "sub_region", "steps",
                                       "min_block_size", "filter_mode", "filter_size", "bidir")]


_P = C.c_void_p
_I = C.c_int
_PROTOS = {
    "memci_version": ([], _I),
    "memci_device_count": ([], _I),
    "memci_ctx_create": ([_I, _I, _I, _P, C.POINTER(_P)], _I),
    "memci_ctx_destroy": ([_P], _I),
    "memci_last_error": ([_P], C.c_char_p),
    "memci_sync": ([_P], _I),
    "memci_stream": ([_P], _P),
    "memci_host_alloc": ([C.c_size_t, C.POINTER(_P)], _I),
    "memci_host_free": ([_P], _I),
    "memci_upload_frame": ([_P, _I, _P], _I),
    "memci_set_frame_device": ([_P, _I, _P], _I),
    "memci_download_frame": ([_P, _I, _P], _I),
    "memci_frame_device_ptr": ([_P, _I], _P),
    "memci_stager_create": ([_I, _I, _I, _I, C.POINTER(_P)], _I),
    "memci_stager_destroy": ([_P], _I),
    "memci_stager_upload": ([_P, _I, _P], _I),
    "memci_frame_from_stager": ([_P, _I, _P, _I], _I),
    "memci_download_frame_via": ([_P, _I, _P, _P], _I),
    "memci_stager_sync": ([_P], _I),
    "memci_me_fs": ([_P, _I, _I, _I, _I, _I], _I),
    "memci_me_fs_pair": ([_P, _I, _I, _I, _I, _I, _I], _I),
    "memci_me_tss": ([_P, _I, _I, _I, _I, _I], _I),
    "memci_me_hbma": ([_P] + [_I] * 9, _I),
    "memci_download_pyramid_level": ([_P, _I, _I, _P], _I),
    "memci_smooth": ([_P, _I, _I, _I, _I], _I),
    "memci_mv_info": ([_P, _I] + [C.POINTER(_I)] * 6, _I),
    "memci_download_mv_blocks": ([_P, _I, _P, _P], _I),
    "memci_download_mv_dense": ([_P, _I, _P], _I),
    "memci_upload_mv_dense": ([_P, _I, _P], _I),
    "memci_upload_mv_blocks": ([_P, _I, _I, _P, _I, _P], _I),
    "memci_mci_unidir": ([_P, _I, _I, _I, C.c_float, _I], _I),
    "memci_mci_bidir": ([_P, _I, _I, _I, _I, C.c_float, _I], _I),
    "memci_mci_bidir2": ([_P, _I, _I, _I, _I, C.c_float, C.c_float, _I, _I, _P, _I], _I),
    "memci_bidir2_stats": ([_P, C.POINTER(_I), C.POINTER(_I)], _I),
    "memci_bidir2_out_of_range_total": ([_P, C.POINTER(_I), _I], _I),
    "memci_quality": ([_P, _I, _I, C.POINTER(C.c_double), C.POINTER(C.c_double)], _I),
    "memci_me_fs_rows": ([_P, _I, _I, _I, _I, _I, _I, _I], _I),
    "memci_mci_rows": ([_P, _I, _I, _I, _I, C.c_float, _I, _I, _I, _I], _I),
    "memci_upload_frame_rows": ([_P, _I, _P, _I, _I], _I),
    "memci_download_frame_rows": ([_P, _I, _P, _I, _I], _I),
    "memci_frame_mark": ([_P, _I], _I),
    "memci_mv_alloc": ([_P, _I, _I], _I),
    "memci_mv_device_ptrs": ([_P, _I, C.POINTER(_P), C.POINTER(_P)], _I),
    "memci_estimate_pair": ([_P, C.POINTER(PairParams), _I, _I, _I, _I], _I),
    "memci_launch_count": ([_P], C.c_longlong),
    "memci_reset_launch_count": ([_P], _I),
    "memci_fs_stats": ([_P, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(_I)], _I),
    "memci_fs_prefilter_stats": ([_P, C.POINTER(_I), C.POINTER(_I)], _I),
    "memci_mci_stats": ([_P, C.POINTER(_I)], _I),
    "memci_enable_kernel_timing": ([_P, _I], _I),
    "memci_last_kernel_ms": ([_P, C.POINTER(C.c_float), C.POINTER(C.c_float)], _I),
    "memci_profile_start": ([_P, _I], _I),
    "memci_profile_collect": ([_P, C.POINTER(C.c_double), C.POINTER(_I), C.POINTER(C.c_double), C.POINTER(_I)], _I),
    "memci_profile_collect_kinds": ([_P, C.POINTER(C.c_double), C.POINTER(_I), C.POINTER(C.c_double)], _I),
    "memci_microbench": ([_P, _I, _I, C.POINTER(C.c_float), C.POINTER(C.c_double)], _I),
}
EXPORTS = tuple(sorted(_PROTOS))


def lib():
    """Load the shared library (building nothing: see interpolatory_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MemciError(
                f"{LIB_PATH} is missing: run `python -m interpolatory_b200.build` (nvcc, sm_100a). "
                "The MEMCI path has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (args, res) in _PROTOS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def check(ctx, rc):
    if rc != 0:
        msg = lib().memci_last_error(ctx)
        raise MemciError(f"libmemci_b200 error {rc}: {msg.decode() if msg else '?'}")
