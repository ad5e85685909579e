"""ctypes binding of libgdr_b200.so (include/gdr_b200.h).  There is no CPU fallback: every call
raises if the CUDA library is missing or reports an error."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_uint32, c_void_p

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libgdr_b200.so")

GDR_OK = 0
DTYPE_F32, DTYPE_BF16 = 0, 1
ACT = {"none": 0, None: 0, "tanh": 1, "sigmoid": 2}
Q_PER_BEAM, FORCE_SIMT, FORCE_UMMA = 1, 2, 4
SKIP_INVERT, SKIP_SCORE, SKIP_TOPK = 256, 512, 1024
OPTIONS = {"umma_ctas": 1, "umma_min_group": 2, "launch_priorities": 3, "fused_groups": 4, "topk_groups": 5, "topk_wide": 6, "umma_ctas_per_sm": 7}

# every symbol include/gdr_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "gdr_abi_version": (c_int32, []),
    "gdr_last_error": (c_char_p, []),
    "gdr_store_create": (c_int32, [POINTER(c_void_p), c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_int32]),
    "gdr_store_create_shard": (c_int32, [POINTER(c_void_p), c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_int64, c_int32,
                                          c_int32, c_int32, c_int64]),
    "gdr_store_p2p_init": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "gdr_store_p2p_attach": (c_int32, [c_void_p, c_void_p]),
    "gdr_store_p2p_attach_local": (c_int32, [c_void_p, POINTER(c_void_p)]),
    "gdr_xchg_bytes": (c_int64, [c_int32, c_int32, POINTER(c_int64), c_int32]),
    "gdr_xchg_create": (c_int32, [POINTER(c_void_p), c_void_p, c_int32, c_int32, c_int32, POINTER(c_int64), c_int32, c_void_p]),
    "gdr_xchg_attach": (c_int32, [c_void_p, c_void_p]),
    "gdr_xchg_attach_local": (c_int32, [c_void_p, POINTER(c_void_p)]),
    "gdr_xchg_all_gather": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p]),
    "gdr_xchg_part_offset": (c_int64, [c_void_p, c_int32, c_int32]),
    "gdr_xchg_destroy": (c_int32, [c_void_p]),
    "gdr_partition_create": (c_int32, [POINTER(c_void_p), c_int32, c_int32, c_int32]),
    "gdr_partition_sms": (c_int32, [c_void_p, POINTER(c_int32)]),
    "gdr_partition_stream": (c_void_p, [c_void_p, c_int32, c_int32]),
    "gdr_partition_destroy": (c_int32, [c_void_p]),
    "gdr_store_destroy": (c_int32, [c_void_p]),
    "gdr_score_topk": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_float), c_int32, c_int32, c_int32,
                                 c_int32, c_int32, c_uint32, c_void_p, c_void_p, c_void_p]),
    "gdr_score_fused": (c_int32, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p]),
    "gdr_store_reserve": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_uint32, c_void_p]),
    "gdr_store_set_option": (c_int32, [c_void_p, c_int32, c_int32]),
    "gdr_store_last_stats": (c_int32, [c_void_p, POINTER(c_int64), c_void_p]),
    "gdr_store_set_profiling": (c_int32, [c_void_p, c_int32]),
    "gdr_store_last_phase_ms": (c_int32, [c_void_p, POINTER(c_float)]),
    "gdr_cluster_centroids": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "gdr_similarity": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "gdr_merge_topk": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "gdr_contrastive_loss": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_float,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "gdr_trie_create": (c_int32, [POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_int32, c_int32]),
    "gdr_trie_destroy": (c_int32, [c_void_p]),
    "gdr_trie_set_child_order": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "gdr_trie_node_embeddings": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "gdr_tree_match": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "gdr_tree_mask": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int64, c_int32, c_int32,
                                c_int32, c_void_p]),
    "gdr_beam_step": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                c_int32, c_void_p, c_void_p, c_void_p]),
    "gdr_position_mask": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p]),
}


class GdrError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libgdr_b200 error {status}: {message}")
        self.status = status


def lib():
    """Load the shared library (once).  Raises if it has not been built: the product has no other path."""
    global _LIB
    if _LIB is None:
        import torch  # noqa: F401  (loads libcudart.so.12 into the process before our library resolves it)
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing — build it with `python -m gdr_b200._build` "
                               "(or __graft_entry__.build()); gdr_b200 has no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _LIB = handle
    return _LIB


def check(status):
    if status != GDR_OK:
        raise GdrError(status, lib().gdr_last_error().decode(errors="replace"))


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)
