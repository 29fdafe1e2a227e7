"""ctypes binding of libsnpgenie_cuda.so (see include/snpgenie_cuda.h).

The library is built in-tree by __graft_entry__.build() / snpgenie_b200/csrc/Makefile.
There is no fallback: if the shared object is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsnpgenie_cuda.so")

OK, E_ARG, E_STATE, E_INPUT, E_CUDA, E_NOMEM = 0, -1, -2, -3, -4, -5
CODE_UNDEF, CODE_OTHER, HIST_BINS = 64, 65, 66
OPT_KEEP_CODES, OPT_PROFILE, OPT_SITE_COUNTS, OPT_FUSED_KERNEL, OPT_BOOTSTRAP_KERNEL, OPT_BOOT_SLOTS, OPT_INGEST_ASYNC, OPT_FIRST_DEFINED = 1, 2, 3, 4, 5, 6, 7, 8
BOOT_GATHER, BOOT_WEIGHTS, BOOT_CLASSES = 0, 1, 2
SLOTS_CORRECTED, SLOTS_FAITHFUL = 0, 1
PEER_HANDLE_BYTES = 128
FUSED_TMA_TILES, FUSED_LDG_SPANS, FUSED_TMA_SPANS = 0, 1, 2
KERNEL_KINDS = ("pack", "hist", "stats", "sums", "bootstrap", "moments", "window", "other", "fused", "sites")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(snpgenie_b200 has no CPU fallback)"
    )

lib = C.CDLL(LIB_PATH)

_u8p, _u32p, _u64p, _i64p, _i8p, _f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int64, C.c_int8, C.c_double))
_ctx = C.c_void_p


class Job(C.Structure):
    """snpg_job (include/snpgenie_cuda.h): what a whole job computes and where its outputs go."""
    _fields_ = [("struct_size", C.c_uint32), ("reserved", C.c_uint32), ("n_bootstraps", C.c_int64), ("window", C.c_int64),
                ("step", C.c_int64), ("n_window_bootstraps", C.c_int64), ("prod_sums", C.c_void_p), ("prod_se", C.c_void_p),
                ("prod_undefined", C.c_void_p), ("poly", C.c_void_p), ("n_defined", C.c_void_p), ("comparisons", C.c_void_p),
                ("stats4", C.c_void_p), ("conserved", C.c_void_p), ("win_sums", C.c_void_p), ("win_se", C.c_void_p)]


class RWindows(C.Structure):
    """snpg_rwindows (include/snpgenie_cuda.h): the arguments of SNPGenie_sliding_windows.R."""
    _fields_ = [("struct_size", C.c_uint32), ("numerator", C.c_int32), ("denominator", C.c_int32), ("jukes_cantor", C.c_int32),
                ("window", C.c_int64), ("step", C.c_int64), ("n_bootstraps", C.c_int64), ("min_defined", C.c_int64)]


RWIN_COLS = 24
RWIN_NAMES = ("sw_start", "sw_center", "sw_end", "sw_num_replicates", "sw_N_diffs", "sw_S_diffs", "sw_N_sites", "sw_S_sites", "sw_dN", "sw_dS",
              "sw_dNdS", "sw_dN_m_dS", "sw_boot_dN_SE", "sw_boot_dS_SE", "sw_boot_dN_over_dS_SE", "sw_boot_dN_over_dS_P", "sw_boot_dN_m_dS_SE",
              "sw_boot_dN_m_dS_P", "sw_boot_dN_gt_dS_count", "sw_boot_dN_eq_dS_count", "sw_boot_dN_lt_dS_count", "sw_ASL_dN_gt_dS_P",
              "sw_ASL_dN_lt_dS_P", "sw_ASL_dNdS_P")

# every exported symbol of include/snpgenie_cuda.h with its signature
SIGNATURES = {
    "snpg_create": (C.c_int, [C.POINTER(_ctx), C.c_int, C.c_uint64]),
    "snpg_destroy": (None, [_ctx]),
    "snpg_last_error": (C.c_char_p, [_ctx]),
    "snpg_version": (C.c_char_p, []),
    "snpg_get_tables": (C.c_int, [_f64p, _f64p]),
    "snpg_set_products": (C.c_int, [_ctx, C.c_int, _i64p, _i64p, _i64p, _i8p, C.c_int64]),
    "snpg_n_products": (C.c_int, [_ctx]),
    "snpg_n_codons": (C.c_int, [_ctx, C.c_int, _i64p]),
    "snpg_set_shard": (C.c_int, [_ctx, C.c_int, C.c_int]),
    "snpg_shard_plan": (C.c_int, [C.c_int64, C.c_int, C.c_int, _i64p, _i64p]),
    "snpg_shard_range": (C.c_int, [_ctx, _i64p, _i64p, _i64p]),
    "snpg_load_group": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "snpg_load_group_device": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "snpg_read_fasta": (C.c_int, [_ctx, C.c_int, C.c_char_p, C.POINTER(C.c_void_p), _u64p, _u64p]),
    "snpg_fasta_wait": (C.c_int, [_ctx, C.c_int]),
    "snpg_codon_strings": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int64, _u8p]),
    "snpg_drop_group": (C.c_int, [_ctx, C.c_int]),
    "snpg_group_size": (C.c_int, [_ctx, C.c_int, _u64p]),
    "snpg_set_option": (C.c_int, [_ctx, C.c_int, C.c_int64]),
    "snpg_get_codon_indices": (C.c_int, [_ctx, C.c_int, C.c_int, _u8p]),
    "snpg_get_site_counts": (C.c_int, [_ctx, C.c_int, _u32p]),
    "snpg_get_hist": (C.c_int, [_ctx, C.c_int, C.c_int, _u32p, _u8p]),
    "snpg_within": (C.c_int, [_ctx, C.c_int, C.c_int, _u8p, _u32p, _u64p, _f64p, _u8p]),
    "snpg_between": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, _u8p, _u32p, _u32p, _u64p, _f64p, _u8p]),
    "snpg_bootstrap_product": (C.c_int, [_ctx, _f64p, _u8p, C.c_int64, C.c_int64, _f64p, _f64p]),
    "snpg_windows": (C.c_int, [_ctx, _f64p, _u8p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _f64p, _f64p, _i64p]),
    "snpg_sliding_windows_r": (C.c_int, [_ctx, _f64p, _u32p, _i64p, C.c_int64, C.POINTER(RWindows), _i64p, _f64p, C.c_int64]),
    "snpg_pooled_diffs": (C.c_int, [_ctx, _f64p, _f64p, C.c_int64, _f64p, _f64p]),
    "snpg_within_all": (C.c_int, [_ctx, C.c_int, C.c_int64, _f64p, _f64p]),
    "snpg_between_all": (C.c_int, [_ctx, C.POINTER(C.c_int), C.c_int, C.c_int64, C.c_int64, C.c_int64, _f64p, _f64p]),
    "snpg_between_job": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(C.c_uint64), C.c_int64,
                                  C.c_int64, C.c_int64, _f64p, _f64p]),
    "snpg_within_job": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, _f64p, _f64p]),
    "snpg_within_job_ex": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(Job)]),
    "snpg_window_plan": (C.c_int, [_ctx, C.c_int64, C.c_int64, _i64p]),
    "snpg_create_multi": (C.c_int, [C.POINTER(_ctx), C.c_int, C.POINTER(C.c_int), C.c_uint64]),
    "snpg_team_size": (C.c_int, [_ctx]),
    "snpg_peer_export": (C.c_int, [_ctx, C.c_int64, C.c_void_p]),
    "snpg_peer_connect": (C.c_int, [_ctx, C.c_void_p, C.c_int]),
    "snpg_peer_connected": (C.c_int, [_ctx]),
    "snpg_within_stats_device": (C.c_int, [_ctx, C.c_int, C.c_void_p]),
    "snpg_bootstrap_all_device": (C.c_int, [_ctx, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "snpg_product_sums_device": (C.c_int, [_ctx, C.c_void_p, _f64p]),
    "snpg_rep_moments_device": (C.c_int, [_ctx, C.c_void_p, C.c_int64, C.c_int64, _f64p]),
    "snpg_test_philox": (C.c_int, [_ctx, _u32p, _u32p, _u32p]),
    "snpg_staged_columns": (C.c_int, [_ctx, _i64p, _i64p]),
    "snpg_probe_peaks": (C.c_int, [_ctx, _f64p]),
    "snpg_launch_count": (C.c_int64, [_ctx]),
    "snpg_tile_launch_count": (C.c_int64, [_ctx]),
    "snpg_span_tma_launch_count": (C.c_int64, [_ctx]),
    "snpg_graph_replays": (C.c_int64, [_ctx]),
    "snpg_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "snpg_get_profile": (C.c_int, [_ctx, _f64p, _i64p, C.c_int]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class SnpgError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libsnpgenie_cuda error {code}: {message}")
        self.code = code
