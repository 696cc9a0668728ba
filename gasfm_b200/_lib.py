"""ctypes binding of ``libgasfm_b200.so`` (the C ABI declared in ``include/gasfm_b200.h``).

There is deliberately no fallback: if the shared library is missing or an entry point
fails, a ``RuntimeError`` is raised.  The library is built in-tree by
``__graft_entry__.build()`` / ``make -C gasfm_b200/csrc``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GASFM_B200_LIB", os.path.join(_HERE, "lib", "libgasfm_b200.so"))  # override: A/B builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gasfm_b200.h")

_c = ctypes
_P = _c.c_void_p
_I, _L, _F, _SZ = _c.c_int, _c.c_int64, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); mirrors include/gasfm_b200.h one to one
SIGNATURES = {
    "gasfm_abi_version": (_I, []),
    "gasfm_last_error": (_c.c_char_p, []),
    "gasfm_m2sparse_count": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "gasfm_m2sparse_ws_bytes": (_SZ, [_I, _I]),
    "gasfm_m2sparse_fill": (_I, [_P, _P, _P, _I, _I, _L, _P, _P, _P, _P]),
    "gasfm_csr_build_ws_bytes": (_SZ, [_L, _I]),
    "gasfm_csr_build": (_I, [_P, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gasfm_plan_chunks": (_I, [_P, _I, _I, _P, _P, _I, _P]),
    "gasfm_gat_ws_bytes": (_SZ, [_I, _I, _I]),
    "gasfm_gat_edge_fwd": (_I, [_P, _L, _P, _L, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _F, _I,
                                _P, _P, _P, _P, _P]),
    "gasfm_gat_bwd_ws_bytes": (_SZ, [_L, _I, _I, _I, _I]),
    "gasfm_gat_edge_bwd": (_I, [_P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _F,
                                _P, _L, _P, _P, _P, _P]),
    "gasfm_gat_edge_fwd_bf16": (_I, [_P, _L, _P, _L, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _F, _I,
                                     _P, _P, _P, _P, _P]),
    "gasfm_gat_edge_bwd_bf16": (_I, [_P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _F,
                                     _P, _L, _P, _P, _P, _P]),
    "gasfm_seg_sum": (_I, [_P, _L, _I, _P, _P, _I, _I, _P, _P, _I, _F, _I, _P, _P, _P]),
    "gasfm_seg_sum_ws_bytes": (_SZ, [_I, _I]),
    "gasfm_seg_bcast": (_I, [_P, _I, _P, _P, _L, _F, _I, _P, _P]),
    "gasfm_ln_relu_fwd": (_I, [_P, _L, _I, _P, _P, _F, _P, _P, _P, _P]),
    "gasfm_ln_relu_bwd_ws_bytes": (_SZ, [_L, _I]),
    "gasfm_col_sum_ws_bytes": (_SZ, [_L, _I]),
    "gasfm_col_sum": (_I, [_P, _L, _L, _I, _P, _P, _P]),
    "gasfm_ln_relu_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _P, _P]),
    "gasfm_edge_update_fwd": (_I, [_P, _L, _P, _I, _P, _P, _P, _P, _P, _L, _P, _P, _L, _I, _F, _F, _P, _P]),
    "gasfm_split_tf32": (_I, [_P, _P, _P, _L, _P]),
    "gasfm_linear_tf32x3_cat": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _L, _L, _I, _I, _P, _P]),
    "gasfm_split_f16": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "gasfm_debug_set_gemm_trace": (_I, [_P]),
    "gasfm_linear_f16x2_supported": (_I, [_L, _I, _I, _L, _L]),
    "gasfm_linear_f16x2": (_I, [_P, _L, _P, _P, _P, _P, _P, _L, _L, _I, _I, _I, _I, _P, _P]),
    "gasfm_linear_f16x2_cat_supported": (_I, [_L, _I, _I, _I, _L]),
    "gasfm_linear_f16x2_cat": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _L, _L, _I, _P, _P]),
    "gasfm_linear_f16x2_cat_rowmax": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _L, _L, _I, _P, _P]),
    "gasfm_gat_edge_bwd_rowmax_supported": (_I, [_I, _I]),
    "gasfm_gat_edge_bwd_rowmax": (_I, [_P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _F,
                                       _P, _L, _P, _P, _P, _P, _P]),
    "gasfm_update_bwd_views_ws_bytes": (_SZ, [_I, _I]),
    "gasfm_update_bwd_views": (_I, [_P, _L, _I, _P, _P, _I, _F, _P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P]),
    "gasfm_x0_bwd_rowmax": (_I, [_P, _L, _I, _P, _P, _I, _F, _P, _P, _P, _P, _P]),
    "gasfm_linear_f16x2_ln": (_I, [_P, _L, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _L, _L, _I, _I, _I, _P, _P]),
    "gasfm_linear_f16x2_ln_y_supported": (_I, [_L, _I, _I, _L, _L]),
    "gasfm_linear_f16x2_ln_y": (_I, [_P, _L, _P, _P, _F, _P, _P, _P, _L, _P, _P, _P, _P, _P, _L, _L, _I, _I, _I, _P, _P]),
    "gasfm_wgrad_f16x2_supported": (_I, [_L, _I, _I, _L, _L]),
    "gasfm_wgrad_f16x2_ws_bytes": (_SZ, [_I, _I]),
    "gasfm_wgrad_f16x2": (_I, [_P, _L, _P, _L, _P, _P, _L, _I, _I, _P, _P, _P, _P]),
    "gasfm_wgrad_f16x2_multi": (_I, [_P, _P, _I, _P, _L, _P, _P, _L, _I, _I, _P, _P, _P, _P]),
    "gasfm_linear_tf32x3_supported": (_I, [_L, _I, _I, _L, _L]),
    "gasfm_linear_tf32x3": (_I, [_P, _L, _P, _P, _P, _P, _L, _L, _I, _I, _I, _P]),
    "gasfm_wgrad_tf32x3_supported": (_I, [_L, _I, _I, _L, _L]),
    "gasfm_wgrad_tf32x3_ws_bytes": (_SZ, [_I, _I]),
    "gasfm_wgrad_tf32x3": (_I, [_P, _L, _P, _L, _L, _I, _I, _P, _P, _P, _P]),
    "gasfm_wgrad_small_supported": (_I, [_I, _I, _L, _L]),
    "gasfm_wgrad_small_ws_bytes": (_SZ, [_I, _I]),
    "gasfm_wgrad_small": (_I, [_P, _L, _P, _L, _L, _I, _I, _P, _P, _P, _P]),
    "gasfm_x0_bwd_ws_bytes": (_SZ, [_L, _I]),
    "gasfm_x0_bwd": (_I, [_P, _L, _I, _P, _P, _I, _F, _P, _P, _P, _P]),
    "gasfm_esfm_loss_ws_bytes": (_SZ, [_L]),
    "gasfm_esfm_loss_fwd": (_I, [_P, _P, _L, _P, _P, _P, _L, _F, _I, _F, _P, _P, _P]),
    "gasfm_esfm_loss_bwd": (_I, [_P, _P, _L, _P, _P, _P, _L, _F, _I, _F, _P, _P, _I, _P, _P]),
    "gasfm_reproj_error": (_I, [_P, _P, _L, _P, _P, _P, _L, _P, _P, _P]),
    "gasfm_peer_alloc": (_I, [_SZ, _c.POINTER(_P)]),
    "gasfm_peer_free": (_I, [_P]),
    "gasfm_peer_export": (_I, [_P, _P]),
    "gasfm_peer_import": (_I, [_P, _c.POINTER(_P)]),
    "gasfm_peer_close": (_I, [_P]),
    "gasfm_peer_buffer_bytes": (_SZ, [_I, _L]),
    "gasfm_peer_flags_bytes": (_SZ, [_I]),
    "gasfm_peer_comm_create": (_I, [_I, _I, _P, _P, _L, _c.c_double, _c.POINTER(_P)]),
    "gasfm_peer_comm_destroy": (_I, [_P]),
    "gasfm_peer_comm_error": (_I, [_P, _c.POINTER(_I)]),
    "gasfm_peer_allreduce_sum": (_I, [_P, _P, _P, _L, _F, _P]),
    "gasfm_peer_lse_merge": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "gasfm_lse_merge_gathered": (_I, [_P, _I, _L, _P, _I, _I, _I, _P, _P, _P, _P]),
    "gasfm_sum_gathered": (_I, [_P, _I, _L, _L, _F, _P, _P]),
    "gasfm_csr_build_host": (_I, [_P, _L, _I, _I, _P, _P, _P]),
    "gasfm_gat_edge_fwd_host": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _I, _F, _P]),
}

_lib = None
launch_count = 0  # number of C-ABI compute calls made (bench.py reports kernel launches from this)


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"gasfm_b200: {LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C gasfm_b200/csrc`. There is no CPU or PyTorch fallback for the CUDA path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().gasfm_last_error().decode("utf-8", "replace")


def call(name, *args):
    """Invoke an int-returning entry point; raise on a non-zero return code."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    launch_count += 1
    if rc != 0:
        msg = last_error()
        if "out of memory" in msg.lower():
            import torch

            raise torch.cuda.OutOfMemoryError(f"{name}: {msg}")
        raise RuntimeError(f"{name} failed (code {rc}): {msg}")


def call_setup(name, *args):
    """Like ``call`` for entry points that launch no kernel (allocation, IPC set-up, status reads)."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {last_error()}")


def size_query(name, *args):
    return int(getattr(load(), name)(*args))


def ptr(t):
    """Device (or host) address of a tensor, or NULL."""
    return None if t is None else t.data_ptr()


class device_guard:
    """Make ``dev`` the current CUDA device for the duration of a C-ABI call.  Unlike ``torch.cuda.device`` it
    does nothing (one integer compare) when the device is already current -- the common case, ~800 times a step."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev):
        self.idx = dev.index
        self.prev = None

    def __enter__(self):
        import torch

        cur = torch.cuda.current_device()
        if self.idx is not None and cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            import torch

            torch.cuda.set_device(self.prev)
        return False


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream
