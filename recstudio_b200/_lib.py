"""ctypes binding of librsb200.so (the C ABI declared in include/rsb200.h).

PyTorch is used only as the owner of device memory and streams: tensors are
passed as raw ``data_ptr()`` values plus sizes, the stream as the raw
``cudaStream_t``.  There is NO fallback: if the shared library is missing or no
CUDA device is usable every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librsb200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "rsb200.h")

# enums of include/rsb200.h
LOSS_BPR, LOSS_SSM, LOSS_FULL = 0, 1, 2
SCORE_IP, SCORE_EUCLID = 0, 1
PHASE_COUNT, PHASE_SCAN, PHASE_FWD, PHASE_SCATTER, PHASE_ALL = 1, 2, 4, 8, 15
SINK_COMPACT, SINK_DENSE, SINK_APPLY = 0, 1, 2
SHARD_PREP, SHARD_FWD, SHARD_FINISH, SHARD_SCATTER, SHARD_PREP_NEG, SHARD_PREP_POS = 1, 2, 4, 8, 16, 32


class Rsb200Error(RuntimeError):
    pass


_CTYPE = {"int64_t": C.c_int64, "int32_t": C.c_int32, "float": C.c_float, "uint64_t": C.c_uint64,
          "uint32_t": C.c_uint32, "size_t": C.c_size_t}


def _struct_fields(name: str):
    """Parse ``typedef struct <name> { ... } <name>;`` out of rsb200.h so the ctypes
    mirror can never drift from the C declaration."""
    src = open(HEADER).read()
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, re.S)
    if not m:
        raise Rsb200Error("struct %s not found in %s" % (name, HEADER))
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        mm = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.*)$", stmt, re.S)
        ctype, is_ptr, names = mm.group(2), mm.group(3) == "*", mm.group(4)
        for nm in names.split(","):
            nm = nm.strip()
            ptr = is_ptr or nm.startswith("*")
            nm = nm.lstrip("* ")
            fields.append((nm, C.c_void_p if ptr else _CTYPE[ctype]))
    return fields


class PairArgs(C.Structure):
    _fields_ = _struct_fields("rsb200_pair_args")


class PairSizes(C.Structure):
    _fields_ = _struct_fields("rsb200_pair_sizes")


class ShardArgs(C.Structure):
    _fields_ = _struct_fields("rsb200_shard_args")


_lib = None
_lock = threading.Lock()


def declared_symbols():
    """Every function name include/rsb200.h declares."""
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(rsb200_\w+)\s*\(", src)))


def lib():
    """Load librsb200.so (built in-tree by recstudio_b200.build / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise Rsb200Error(
                "librsb200.so is missing (%s). Build it with `python -m recstudio_b200.build` "
                "or __graft_entry__.build(); there is no CPU / PyTorch fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.rsb200_last_error.restype = C.c_char_p
        L.rsb200_version.restype = C.c_int32
        L.rsb200_sizeof_pair_args.restype = C.c_size_t
        if L.rsb200_sizeof_pair_args() != C.sizeof(PairArgs):
            raise Rsb200Error('rsb200_pair_args layout mismatch between rsb200.h and librsb200.so: rebuild')
        L.rsb200_sizeof_shard_args.restype = C.c_size_t
        if L.rsb200_sizeof_shard_args() != C.sizeof(ShardArgs):
            raise Rsb200Error('rsb200_shard_args layout mismatch between rsb200.h and librsb200.so: rebuild')
        L.rsb200_scan_tmp_elems.restype = C.c_int64
        L.rsb200_scan_tmp_elems.argtypes = [C.c_int64]
        L.rsb200_launch_count.restype = C.c_uint64
        L.rsb200_bin_heavy_elems.restype = C.c_int64
        L.rsb200_philox_counter_offset.restype = C.c_int64
        L.rsb200_philox_counter_offset.argtypes = [C.c_int64, C.c_int32, C.c_int32]
        L.rsb200_topk_workspace_bytes.restype = C.c_size_t
        L.rsb200_topk_workspace_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        L.rsb200_fullsoftmax_workspace_bytes.restype = C.c_size_t
        L.rsb200_fullsoftmax_workspace_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int64]
        L.rsb200_index_workspace_bytes.restype = C.c_size_t
        L.rsb200_index_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
        v, i64, i32, u64, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_float
        sigs = {
            "rsb200_device_info": [v, v, v, v],
            "rsb200_sample_uniform": [u64, u64, i64, i64, i64, i32, i32, v, v, v],
            "rsb200_sample_uniform_dev": [v, i64, i64, i64, i32, i32, v, v, v],
            "rsb200_popular_build_guide": [v, i64, i32, v, v],
            "rsb200_popular_build_guide_range": [v, i64, i32, i64, i64, v, v],
            "rsb200_bin_shift": [i64, i64, i64],
            "rsb200_uniform_owner_range": [i64, i64, i64, v, v, v],
            "rsb200_sample_popular": [u64, u64, v, v, i64, i64, i64, i32, i32, v, i32, v, v, v, v],
            "rsb200_popular_logq": [v, i64, v, i64, v, v],
            "rsb200_masked_workspace_elems": [i64, i64, v, v],
            "rsb200_sample_uniform_masked": [u64, u64, i64, v, i64, i64, i64, i32, i32, v, v, v, v, v],
            "rsb200_kmeans_assign": [v, i64, i64, i64, v, i64, v, v, v, v],
            "rsb200_kmeans_update": [v, i64, i64, i64, v, i64, v, v, v],
            "rsb200_index_build": [v, i64, i64, v, v, v, C.c_size_t, v],
            "rsb200_segment_cdf": [v, v, v, i64, v, v, v],
            "rsb200_segment_search": [v, v, i64, v, v, v, i64, v, v, v, v],
            "rsb200_pair_workspace_sizes": [i64, i64, i64, i64, i64, C.POINTER(PairSizes)],
            "rsb200_pair_step": [C.POINTER(PairArgs), i32, v],
            "rsb200_pair_draw_count": [C.POINTER(PairArgs), v, u64, u64, i32, i32, v, v],
            "rsb200_shard_step": [C.POINTER(ShardArgs), i32, v],
            "rsb200_gather_rows": [v, i64, i64, v, i64, v, v],
            "rsb200_scatter_add_rows": [v, i64, i64, v, i64, v, v],
            "rsb200_rows_update": [i32, v, v, v, i64, i64, v, v, v, i64, i64, f32, f32, f32, f32, v],
            "rsb200_attn_fwd": [v, v, v, v, i64, i64, i64, i64, i32, f32, u64, v, v, v, v],
            "rsb200_attn_bwd": [v, v, v, v, v, v, v, i64, i64, i64, i64, i32, f32, u64, v, v, v, v, v],
            "rsb200_tc_gemm_test": [v, v, v, i64, i64, v, v],
            "rsb200_rows_coalesce": [v, v, i64, i64, i64, i32, v, v, i32, i32, v, v, v, v, v, i64, v, i64, v, v],
            "rsb200_score_ids": [i32, v, v, i64, i64, v, i64, i64, v, v],
            "rsb200_score_dense": [i32, v, v, i64, i64, i64, v, v],
            "rsb200_score_dense_bwd": [i32, v, v, v, i64, i64, i64, v, v, v],
            "rsb200_pair_loss": [i32, v, v, v, v, i64, i64, v, v, v, v, v],
            "rsb200_topk_full": [i32, v, v, i64, i64, i64, i64, v, i64, v, v, v, C.c_size_t, v],
            "rsb200_fullsoftmax_fwd_bwd": [v, v, v, i64, i64, i64, v, v, v, v, C.c_size_t, v],
        }
        for name, argt in sigs.items():
            fn = getattr(L, name)
            fn.restype = C.c_int32
            fn.argtypes = argt
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().rsb200_last_error()
        raise Rsb200Error("%s failed (rc=%d): %s" % (what or "librsb200 call", rc, msg.decode() if msg else "?"))


def require_cuda():
    """The product path has no CPU fallback: fail loudly without a CUDA device."""
    import torch
    if not torch.cuda.is_available():
        raise Rsb200Error("recstudio_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


_dev_info = {}


def device_info(device_index: int):
    """(sm_count, max_threads_per_sm) of a device; fixes torch's Philox element mapping."""
    import torch
    if device_index not in _dev_info:
        require_cuda()
        with torch.cuda.device(device_index):
            sm, mt, maj, mnr = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            check(lib().rsb200_device_info(C.byref(sm), C.byref(mt), C.byref(maj), C.byref(mnr)), "device_info")
        _dev_info[device_index] = (sm.value, mt.value, maj.value, mnr.value)
    return _dev_info[device_index]


def ptr(t):
    """data_ptr of a tensor or 0 for None."""
    return 0 if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
