"""ctypes binding of librag_b200.so (the C ABI declared in include/rag_b200.h).

There is deliberately no fallback: if the library is missing, or there is no
sm_100 device when a store is created, callers get an exception -- never a CPU
code path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librag_b200.so")

DTYPE_F32, DTYPE_BF16 = 0, 1
SPACE_L2, SPACE_COSINE, SPACE_IP = 0, 1, 2
SPACES = {"l2": SPACE_L2, "cosine": SPACE_COSINE, "ip": SPACE_IP}
DTYPES = {"f32": DTYPE_F32, "fp32": DTYPE_F32, "float32": DTYPE_F32,
          "bf16": DTYPE_BF16, "bfloat16": DTYPE_BF16}
QUERY_AUTO, QUERY_FORCE_STREAM, QUERY_FORCE_TENSOR = 0, 1, 2
EINVAL, ECUDA, ENOMEM, ENODEV = -1, -2, -3, -4
EMPTY_KEY = 0xFFFFFFFFFFFFFFFF
MAX_K = 1024
MAX_MASK_SLOTS = 16
EXCHANGE_HANDLE_BYTES = 64
STORE_NO_RERANK = 1
ABI_VERSION = 3
F32_SHADOW_AUTO, F32_SHADOW_HI, F32_SHADOW_HILO = 0, 1, 2

_p = C.c_void_p
_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); must list every symbol include/rag_b200.h declares
SIGNATURES = {
    "rag_last_error": (C.c_char_p, []),
    "rag_abi_version": (C.c_int, []),
    "rag_device_count": (C.c_int, []),
    "rag_store_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(_p)]),
    "rag_store_create_ex": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(_p)]),
    "rag_store_destroy": (C.c_int, [_p]),
    "rag_store_reserve": (C.c_int, [_p, C.c_int64]),
    "rag_store_upsert": (C.c_int, [_p, C.c_int64, _p, _p, _p]),
    "rag_store_flush": (C.c_int, [_p]),
    "rag_store_upsert_dev": (C.c_int, [_p, C.c_int64, _p, _p, _p]),
    "rag_store_delete": (C.c_int, [_p, C.c_int64, _p]),
    "rag_store_count": (C.c_int64, [_p]),
    "rag_store_rows": (C.c_int64, [_p]),
    "rag_store_capacity": (C.c_int64, [_p]),
    "rag_store_dim": (C.c_int, [_p]),
    "rag_store_dtype": (C.c_int, [_p]),
    "rag_store_space": (C.c_int, [_p]),
    "rag_store_device": (C.c_int, [_p]),
    "rag_store_has_rerank": (C.c_int, [_p]),
    "rag_store_is_live": (C.c_int, [_p, C.c_int64]),
    "rag_store_kernel_launches": (C.c_int64, [_p]),
    "rag_store_fetch": (C.c_int, [_p, C.c_int64, _p, _p]),
    "rag_store_fetch_exact": (C.c_int, [_p, C.c_int64, _p, _p]),
    "rag_store_patch_mask": (C.c_int, [_p, C.c_int, C.c_int64, _p, _p]),
    "rag_store_set_mask": (C.c_int, [_p, C.c_int, _p, C.c_int64]),
    "rag_store_clear_mask": (C.c_int, [_p, C.c_int]),
    "rag_store_query": (C.c_int, [_p, C.c_int, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p]),
    "rag_store_query_dev": (C.c_int, [_p, C.c_int, _p, C.c_int, C.c_int, C.c_int, C.c_uint32, _p, _p, _p, _p, _p]),
    "rag_merge_keys_dev": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p, _p, _p]),
    "rag_exchange_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(_p)]),
    "rag_exchange_handle": (C.c_int, [_p, _p]),
    "rag_exchange_connect": (C.c_int, [_p, _p]),
    "rag_exchange_status": (C.c_int, [_p, _i32p]),
    "rag_exchange_destroy": (C.c_int, [_p]),
    "rag_store_fused_ok": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int]),
    "rag_store_query_fused_dev": (C.c_int, [_p, _p, C.c_int, _p, C.c_int, C.c_int, C.c_int, C.c_uint32, _p, _p, _p, _p]),
    "rag_store_query_fused": (C.c_int, [_p, _p, C.c_int, _p, C.c_int, C.c_int, C.c_int, C.c_uint32, _p, _p, _p]),
    "rag_store_query_submit": (C.c_int, [_p, _p, C.c_int, _p, C.c_int, C.c_int, C.c_int, C.c_uint32, _i32p]),
    "rag_store_query_wait": (C.c_int, [_p, C.c_int, _p, _p, _p]),
    "rag_sharded_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _p, C.c_int64, C.c_int, C.POINTER(_p)]),
    "rag_sharded_destroy": (C.c_int, [_p]),
    "rag_sharded_shards": (C.c_int, [_p]),
    "rag_sharded_shard": (_p, [_p, C.c_int]),
    "rag_sharded_fused": (C.c_int, [_p]),
    "rag_sharded_count": (C.c_int64, [_p]),
    "rag_sharded_rows": (C.c_int64, [_p]),
    "rag_sharded_is_live": (C.c_int, [_p, C.c_int64]),
    "rag_sharded_reserve": (C.c_int, [_p, C.c_int64]),
    "rag_sharded_flush": (C.c_int, [_p]),
    "rag_sharded_upsert": (C.c_int, [_p, C.c_int64, _p, _p, _p]),
    "rag_sharded_delete": (C.c_int, [_p, C.c_int64, _p]),
    "rag_sharded_fetch": (C.c_int, [_p, C.c_int64, _p, _p, C.c_int]),
    "rag_sharded_set_mask": (C.c_int, [_p, C.c_int, _p, C.c_int64]),
    "rag_sharded_patch_mask": (C.c_int, [_p, C.c_int, C.c_int64, _p, _p]),
    "rag_sharded_clear_mask": (C.c_int, [_p, C.c_int]),
    "rag_sharded_query": (C.c_int, [_p, C.c_int, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p]),
    "rag_sharded_last_query_info": (C.c_int, [_p, _f32p, _i32p, _i32p, _i32p]),
    "rag_key_pack": (C.c_uint64, [C.c_float, C.c_uint32]),
    "rag_key_dist": (C.c_float, [C.c_uint64]),
    "rag_key_row": (C.c_uint32, [C.c_uint64]),
    "rag_debug_tensor_stats": (C.c_int, [_u64p, C.c_int]),
    "rag_store_last_upsert_ms": (C.c_float, [_p]),
    "rag_store_last_query_info": (C.c_int, [_p, _f32p, _i32p, _i32p]),
    "rag_debug_ring_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, _i32p, _i32p]),
    "rag_store_set_f32_shadow": (C.c_int, [_p, C.c_int]),
    "rag_store_f32_tensor_info": (C.c_int, [_p, _i32p, _i64p, _i64p]),
}

_lib = None


class EngineError(RuntimeError):
    """CUDA / engine failure (RAG_ECUDA, RAG_ENOMEM, RAG_ENODEV)."""


def load():
    """Load the engine library, binding every declared symbol.  Raises if the
    library has not been built -- there is no Python/CPU substitute."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `python local-rag-system_b200/build.py` "
            "(or __graft_entry__.build()).  The engine has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.rag_abi_version() != ABI_VERSION:
        raise EngineError(f"ABI version mismatch: library reports {lib.rag_abi_version()}")
    _lib = lib
    return lib


def last_error() -> str:
    return (load().rag_last_error() or b"").decode("utf-8", "replace")


def check(rc: int):
    """Translate a C-ABI status into the reference's exception convention
    (SURVEY.md 8b): ValueError for bad input, RuntimeError otherwise."""
    if rc == 0:
        return
    msg = last_error()
    if rc == EINVAL:
        raise ValueError(msg)
    raise EngineError(f"[rag_b200 {rc}] {msg}")
