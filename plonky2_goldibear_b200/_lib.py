"""ctypes binding of libp2g_cuda.so (include/p2g.h).  Loading fails loudly when the library has not been
built (python -c 'import __graft_entry__ as g; g.build()') -- there is no fallback implementation."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LOCK = threading.Lock()

FLAG_DEVICE_INPUT = 1
FLAG_DEVICE_OUTPUT = 2
FLAG_NO_LEAVES_OUT = 4
FLAG_KEEP_ON_DEVICE = 8
FLAG_LEAVES_ONLY = 16
FLAG_LEAVES_READY = 32

ERR_NAMES = {0: "OK", -1: "INVALID", -2: "NO_DEVICE", -3: "CUDA", -4: "OOM", -5: "UNSUPPORTED"}


class P2GError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("p2g error %s (%d): %s" % (ERR_NAMES.get(code, "?"), code, message))
        self.code = code


class CommitDesc(ctypes.Structure):
    _fields_ = [
        ("field", ctypes.c_int32), ("hash", ctypes.c_int32), ("repr", ctypes.c_int32), ("input_kind", ctypes.c_int32),
        ("log_n", ctypes.c_uint32), ("num_cols", ctypes.c_uint32), ("rate_bits", ctypes.c_uint32),
        ("cap_height", ctypes.c_uint32), ("num_salt_cols", ctypes.c_uint32), ("flags", ctypes.c_uint32),
        ("col_ptrs", ctypes.POINTER(ctypes.c_void_p)), ("salt_ptrs", ctypes.POINTER(ctypes.c_void_p)),
        ("coeffs_out", ctypes.POINTER(ctypes.c_void_p)), ("leaves_out", ctypes.c_void_p),
        ("leaf_pitch", ctypes.c_uint64), ("digests_out", ctypes.c_void_p), ("cap_out", ctypes.c_void_p),
        ("shard_index", ctypes.c_uint32), ("shard_count", ctypes.c_uint32),
        ("stage_seconds", ctypes.c_double * 4), ("h2d_seconds", ctypes.c_double), ("d2h_seconds", ctypes.c_double),
        ("kernel_seconds", ctypes.c_double * 4), ("batch_out", ctypes.POINTER(ctypes.c_void_p)),
    ]


class MerkleDesc(ctypes.Structure):
    _fields_ = [
        ("field", ctypes.c_int32), ("hash", ctypes.c_int32), ("repr", ctypes.c_int32), ("flags", ctypes.c_uint32),
        ("num_leaves", ctypes.c_uint64), ("leaf_len", ctypes.c_uint64), ("leaf_pitch", ctypes.c_uint64),
        ("cap_height", ctypes.c_uint32), ("leaves", ctypes.c_void_p), ("digests_out", ctypes.c_void_p),
        ("cap_out", ctypes.c_void_p), ("seconds", ctypes.c_double),
    ]


EXPORTS = [
    "p2g_abi_version", "p2g_init", "p2g_destroy", "p2g_last_error", "p2g_host_alloc", "p2g_host_free", "p2g_commit",
    "p2g_merkle", "p2g_merkle_finish_cap", "p2g_permute", "p2g_transform", "p2g_workspace_bytes",
    "p2g_release_workspace", "p2g_stream", "p2g_kernel_launches", "p2g_batch_get_rows", "p2g_batch_prove",
    "p2g_batch_download", "p2g_batch_shape", "p2g_batch_free", "p2g_dev_alloc", "p2g_dev_free", "p2g_ipc_export",
    "p2g_ipc_import", "p2g_ipc_close", "p2g_intt_allgather", "p2g_lde_columns", "p2g_memcpy_async",
    "p2g_fri_begin", "p2g_fri_commit_layer", "p2g_fri_fold", "p2g_fri_finish", "p2g_fri_shape", "p2g_fri_free", "p2g_fri_pow",
    "p2g_num_devices", "p2g_sub_context",
]


def library_path() -> str:
    # P2G_LIBRARY lets kernel experiments load an alternative build of the same CUDA library
    return os.environ.get("P2G_LIBRARY") or os.path.join(_HERE, "libp2g_cuda.so")


def load_library() -> ctypes.CDLL:
    global _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). The commit path has no CPU fallback." % path)
        lib = ctypes.CDLL(path)
        vp, i32, u32, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint64
        lib.p2g_abi_version.restype = i32
        lib.p2g_init.argtypes = [ctypes.POINTER(i32), i32, ctypes.POINTER(vp)]
        lib.p2g_init.restype = i32
        lib.p2g_destroy.argtypes = [vp]
        lib.p2g_destroy.restype = None
        lib.p2g_last_error.argtypes = [vp]
        lib.p2g_last_error.restype = ctypes.c_char_p
        lib.p2g_host_alloc.argtypes = [vp, ctypes.c_size_t]
        lib.p2g_host_alloc.restype = vp
        lib.p2g_host_free.argtypes = [vp, vp]
        lib.p2g_host_free.restype = None
        lib.p2g_commit.argtypes = [vp, ctypes.POINTER(CommitDesc)]
        lib.p2g_commit.restype = i32
        lib.p2g_merkle.argtypes = [vp, ctypes.POINTER(MerkleDesc)]
        lib.p2g_merkle.restype = i32
        lib.p2g_merkle_finish_cap.argtypes = [vp, i32, i32, i32, vp, u32, u32, vp, vp]
        lib.p2g_merkle_finish_cap.restype = i32
        lib.p2g_permute.argtypes = [vp, i32, i32, vp, u64]
        lib.p2g_permute.restype = i32
        lib.p2g_transform.argtypes = [vp, i32, i32, i32, u32, u32, u32, u32, vp, vp]
        lib.p2g_transform.restype = i32
        lib.p2g_workspace_bytes.argtypes = [vp]
        lib.p2g_workspace_bytes.restype = u64
        lib.p2g_release_workspace.argtypes = [vp]
        lib.p2g_release_workspace.restype = None
        lib.p2g_stream.argtypes = [vp]
        lib.p2g_stream.restype = vp
        lib.p2g_kernel_launches.argtypes = [vp]
        lib.p2g_kernel_launches.restype = u64
        lib.p2g_batch_get_rows.argtypes = [vp, vp, vp, u64, vp]
        lib.p2g_batch_get_rows.restype = i32
        lib.p2g_batch_prove.argtypes = [vp, vp, vp, u64, vp]
        lib.p2g_batch_prove.restype = i32
        lib.p2g_batch_download.argtypes = [vp, vp, i32, vp]
        lib.p2g_batch_download.restype = i32
        lib.p2g_batch_shape.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.POINTER(u32), ctypes.POINTER(u64)]
        lib.p2g_batch_shape.restype = i32
        lib.p2g_batch_free.argtypes = [vp, vp]
        lib.p2g_batch_free.restype = None
        lib.p2g_dev_alloc.argtypes = [vp, ctypes.c_size_t]
        lib.p2g_dev_alloc.restype = vp
        lib.p2g_dev_free.argtypes = [vp, vp]
        lib.p2g_dev_free.restype = None
        lib.p2g_ipc_export.argtypes = [vp, vp, ctypes.c_char_p]
        lib.p2g_ipc_export.restype = i32
        lib.p2g_ipc_import.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]
        lib.p2g_ipc_import.restype = i32
        lib.p2g_ipc_close.argtypes = [vp, vp]
        lib.p2g_ipc_close.restype = i32
        lib.p2g_intt_allgather.argtypes = [vp, i32, i32, u32, u32, u32, vp, ctypes.POINTER(vp), u32]
        lib.p2g_intt_allgather.restype = i32
        lib.p2g_lde_columns.argtypes = [vp, i32, i32, u32, u32, u32, u32, u32, u32, u32, vp, vp, u64, i32]
        lib.p2g_lde_columns.restype = i32
        lib.p2g_memcpy_async.argtypes = [vp, vp, vp, ctypes.c_size_t, i32]
        lib.p2g_memcpy_async.restype = i32
        lib.p2g_fri_begin.argtypes = [vp, i32, i32, i32, u32, u32, vp, vp, u32, ctypes.POINTER(vp)]
        lib.p2g_fri_begin.restype = i32
        lib.p2g_fri_commit_layer.argtypes = [vp, vp, u32, u32, vp, ctypes.POINTER(vp)]
        lib.p2g_fri_commit_layer.restype = i32
        lib.p2g_fri_fold.argtypes = [vp, vp, vp]
        lib.p2g_fri_fold.restype = i32
        lib.p2g_fri_finish.argtypes = [vp, vp, u32, vp, ctypes.POINTER(u64)]
        lib.p2g_fri_finish.restype = i32
        lib.p2g_fri_shape.argtypes = [vp, ctypes.POINTER(u32), ctypes.POINTER(u64)]
        lib.p2g_fri_shape.restype = i32
        lib.p2g_fri_pow.argtypes = [vp, i32, i32, vp, vp, u32, u32, ctypes.POINTER(u64)]
        lib.p2g_fri_pow.restype = i32
        lib.p2g_fri_free.argtypes = [vp, vp]
        lib.p2g_fri_free.restype = None
        lib.p2g_num_devices.argtypes = [vp]
        lib.p2g_num_devices.restype = i32
        lib.p2g_sub_context.argtypes = [vp, i32]
        lib.p2g_sub_context.restype = vp
        _LIB = lib
        return lib


class Context:
    """One GPU context (p2g_init): a stream, resident tables and a reusable device workspace."""

    def __init__(self, device=0):
        """device: one device index, or a list of 2 / 4 / 8 indices for a multi-device context (one process drives them
        all; p2g_commit shards the job by LDE coset, include/p2g.h)."""
        self.lib = load_library()
        self.handle = ctypes.c_void_p()
        devices = [int(d) for d in device] if isinstance(device, (list, tuple)) else [int(device)]
        self.devices = devices
        device = devices[0]
        dev = (ctypes.c_int32 * len(devices))(*devices)
        rc = self.lib.p2g_init(dev, len(devices), ctypes.byref(self.handle))
        if rc != 0:
            msg = self.lib.p2g_last_error(None)
            raise P2GError(rc, msg.decode() if msg else "p2g_init failed")
        self.device = device
        self._pinned = []

    def check(self, rc: int):
        if rc != 0:
            msg = self.lib.p2g_last_error(self.handle)
            raise P2GError(rc, msg.decode() if msg else "")

    def kernel_launches(self) -> int:
        return int(self.lib.p2g_kernel_launches(self.handle))

    def workspace_bytes(self) -> int:
        return int(self.lib.p2g_workspace_bytes(self.handle))

    def release_workspace(self):
        self.lib.p2g_release_workspace(self.handle)

    def pinned_empty(self, shape, dtype):
        """numpy array backed by pinned host memory from p2g_host_alloc."""
        import numpy as np
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = max(count * dtype.itemsize, 1)
        ptr = self.lib.p2g_host_alloc(self.handle, nbytes)
        if not ptr:
            self.check(-4)
        buf = (ctypes.c_ubyte * nbytes).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
        self._pinned.append(ptr)
        return arr

    def close(self):
        if self.handle:
            for p in self._pinned:
                self.lib.p2g_host_free(self.handle, p)
            self._pinned = []
            self.lib.p2g_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_DEFAULT = {}


def default_context(device: int = 0) -> Context:
    if device not in _DEFAULT:
        _DEFAULT[device] = Context(device)
    return _DEFAULT[device]
