"""ctypes binding of libdinox_b200.so (C ABI in include/dinox_b200.h) and its in-tree build.

The shared library is built with nvcc for sm_100a only and lives next to this file so that it
travels with the source tree.  There is no fallback: if the library is missing or the device
is not a B200, the compute API raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# Experiment variants (tools/ only): DINOX_LIB_TAG=<tag> loads / builds libdinox_b200_<tag>.so, compiled
# with the extra -D flags given to build(extra_flags=...).  The product library has no tag.
_TAG = os.environ.get("DINOX_LIB_TAG", "")
LIB_PATH = os.path.join(HERE, f"libdinox_b200{'_' + _TAG if _TAG else ''}.so")
BUILD_DIR = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--use_fast_math", "-Xptxas", "-v",
    "-DDINOX_BUILD",
]
# --use_fast_math would turn expf/logf/division into approximations everywhere; the kernels pick
# intrinsics explicitly instead, so it is NOT enabled (kept out of NVCC_FLAGS below).
NVCC_FLAGS.remove("--use_fast_math")


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target: str, deps: List[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags: Optional[List[str]] = None) -> str:
    """Compile every .cu under csrc/ for sm_100a and link libdinox_b200.so (in-tree)."""
    extra_flags = list(extra_flags or [])
    os.makedirs(BUILD_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "dinox_b200.h"))
    srcs = sources()
    objs = [os.path.join(BUILD_DIR, os.path.basename(s)[:-3] + ".o") for s in srcs]

    def compile_one(args):
        src, obj = args
        if not force and not _stale(obj, [src] + headers):
            return ""
        cmd = [_nvcc()] + NVCC_FLAGS + extra_flags + ["-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        logs = list(ex.map(compile_one, zip(srcs, objs)))
    if verbose:
        for l in logs:
            if l:
                print(l)
    if force or _stale(LIB_PATH, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


_lib: Optional[ctypes.CDLL] = None


class DinoxError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the shared library (building is explicit: ``python -m dinox_b200._ext`` or
    ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DinoxError(
                f"{LIB_PATH} not found: build it with `python -m dinox_b200._ext` "
                "(nvcc, sm_100a). dinox_b200 has no CPU or eager fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


c_void_p, c_int, c_i64, c_f32, c_size = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); kept in sync with include/dinox_b200.h (tests/test_abi.py checks it)
SIGNATURES = {
    "dinox_version": (c_int, []),
    "dinox_last_error_string": (ctypes.c_char_p, []),
    "dinox_device_check": (c_int, []),
    "dinox_launch_count": (c_i64, []),
    "dinox_launch_count_reset": (None, []),
    "dinox_trace_begin": (c_int, [c_void_p, c_int]),
    "dinox_trace_end": (c_int, []),
    "dinox_trace_name": (ctypes.c_char_p, [c_int]),
    "dinox_trace_stream": (ctypes.c_uint64, [c_int]),
    "dinox_ema_plan_create": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dinox_ema_plan_destroy": (c_int, [c_void_p]),
    "dinox_ema_plan_numel": (c_i64, [c_void_p]),
    "dinox_ema_apply": (c_int, [c_void_p, c_f32, c_f32, c_void_p]),
    "dinox_rows_lse": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_f32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dinox_cols_lse": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_f32, c_void_p, c_void_p, c_void_p]),
    "dinox_lse_combine": (c_int, [c_void_p, c_int, c_i64, c_f32, c_void_p, c_void_p]),
    "dinox_cols_sum": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_void_p]),
    "dinox_cols_sum_axpy": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_f32, c_void_p, c_void_p, c_int, c_void_p]),
    "dinox_cols_sum_chunked": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_i64, c_void_p, c_void_p]),
    "dinox_center_ema": (c_int, [c_void_p, c_void_p, c_f32, c_f32, c_i64, c_void_p]),
    "dinox_axpb": (c_int, [c_void_p, c_f32, c_f32, c_void_p, c_i64, c_void_p]),
    "dinox_axpby": (c_int, [c_void_p, c_f32, c_void_p, c_void_p, c_f32, c_void_p, c_i64, c_void_p]),
    "dinox_ce_workspace_bytes": (c_size, [c_i64, c_i64]),
    "dinox_ce_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_i64, c_int, c_int, c_i64, c_i64, c_i64,
                             c_f32, c_f32, c_void_p, c_void_p, c_void_p, c_void_p, c_f32, c_int,
                             c_void_p, c_void_p, c_void_p]),
    "dinox_ce_onepass_max_views": (c_int, []),
    "dinox_ce_onepass_workspace_bytes": (c_size, [c_i64, c_int, c_int, c_i64]),
    "dinox_ce_fwd_onepass": (c_int, [c_void_p, c_int, c_void_p, c_int, c_i64, c_int, c_int, c_i64, c_i64, c_i64,
                                     c_f32, c_f32, c_void_p, c_void_p, c_f32, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dinox_ce_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_i64, c_int, c_int, c_i64, c_i64, c_i64,
                             c_f32, c_f32, c_void_p, c_void_p, c_void_p, c_void_p, c_f32, c_int,
                             c_void_p, c_void_p, c_i64, c_void_p]),
    "dinox_gemm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int,
                                c_int, c_int, c_f32, c_void_p, c_void_p, c_int, c_void_p]),
    "dinox_gemm_bf16_splitk": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                       c_int, c_int, c_int, c_f32, c_void_p, c_int, c_void_p]),
    "dinox_gemm_splitk_plan": (c_int, [c_i64, c_i64, c_i64]),
    "dinox_gemm_bf16_reduce_scatter": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                               c_i64, c_int, c_int, c_f32, c_void_p, c_void_p, c_i64, c_f32, c_void_p]),
    "dinox_debug_max_active_clusters": (c_int, [c_int]),
    "dinox_head_stats_workspace_bytes": (c_size, [c_i64, c_i64]),
    "dinox_head_stats": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_f32, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "dinox_head_grad_workspace_bytes": (c_size, [c_i64, c_i64]),
    "dinox_head_grad_db2_rows": (c_i64, [c_i64]),
    "dinox_head_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                c_f32, c_f32, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_i64, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "dinox_head_teacher_workspace_bytes": (c_size, [c_i64, c_i64]),
    "dinox_head_teacher_granules_per_tile": (c_int, []),
    "dinox_head_teacher_tile_cols": (c_int, []),
    "dinox_head_teacher": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_f32, c_void_p, c_void_p, c_i64,
                                   c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dinox_head_grad2_workspace_bytes": (c_size, [c_i64, c_i64]),
    "dinox_head_grad2_db2_rows": (c_i64, [c_i64]),
    "dinox_head_grad2": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_f32, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i64, c_void_p, c_i64,
                                 c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "dinox_gemm_bf16_batched": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                        c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_f32, c_void_p, c_void_p]),
    "dinox_normalize_tokens": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_void_p, c_void_p, c_void_p]),
    "dinox_normalize_tokens_bwd": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_void_p, c_void_p,
                                           c_void_p, c_f32, c_void_p, c_i64, c_i64, c_void_p]),
    "dinox_gram_diff_workspace_bytes": (c_size, [c_i64, c_i64]),
    "dinox_gram_diff": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p, c_i64, c_f32, c_void_p, c_void_p, c_void_p]),
    "dinox_gemm_bf16_balanced_workspace_bytes": (c_size, [c_i64, c_i64]),
    "dinox_plan_ordered_split": (c_int, [c_i64, c_i64, c_i64, c_void_p, c_void_p]),
    "dinox_debug_walk_ordered": (c_i64, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_i64]),
    "dinox_gemm_bf16_balanced": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int,
                                         c_int, c_f32, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "dinox_gather_cast_bf16_2": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_int, c_i64,
                                         c_void_p, c_i64, c_void_p]),
    "dinox_gelu_bwd_gather": (c_int, [c_void_p, c_i64, c_int, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    "dinox_segment_cols_sum_chunks": (c_i64, [c_i64, c_i64]),
    "dinox_segment_cols_sum_workspace_bytes": (c_size, [c_i64, c_i64, c_i64]),
    "dinox_segment_cols_sum": (c_int, [c_void_p, c_int, c_i64, c_i64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p]),
    "dinox_gather_cast_bf16": (c_int, [c_void_p, c_int, c_i64, c_void_p, c_i64, c_i64, c_void_p, c_i64, c_void_p]),
    "dinox_gather_f32": (c_int, [c_void_p, c_void_p, c_i64, c_f32, c_void_p, c_void_p]),
    "dinox_scatter_rows_f32": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_i64, c_void_p, c_i64, c_void_p]),
    "dinox_gather_rows_f32": (c_int, [c_void_p, c_int, c_i64, c_void_p, c_i64, c_i64, c_void_p, c_i64, c_void_p]),
    "dinox_scatter_add_rows_f32": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_i64, c_void_p, c_i64, c_void_p]),
    "dinox_gelu_fwd": (c_int, [c_void_p, c_i64, c_void_p, c_void_p]),
    "dinox_gelu_bwd_workspace_bytes": (c_size, [c_i64, c_i64]),
    "dinox_gelu_bwd": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dinox_gemv_bf16": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_i64, c_f32, c_void_p, c_f32, c_void_p, c_void_p]),
    "dinox_gemv_bf16_multi": (c_int, [c_void_p, c_i64, c_void_p, c_int, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_f32,
                                      c_void_p, c_void_p]),
    "dinox_gemv_bf16_multi_ema": (c_int, [c_void_p, c_i64, c_void_p, c_int, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_f32,
                                          c_void_p, c_void_p, c_void_p]),
    "dinox_sum_slabs": (c_int, [c_void_p, c_int, c_i64, c_i64, c_void_p, c_f32, c_void_p, c_int, c_void_p]),
    "dinox_gather_sum_rows": (c_int, [c_void_p, c_i64, c_int, c_i64, c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_f32,
                                      c_void_p, c_i64, c_int, c_void_p]),
    "dinox_head_offsets": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_f32, c_f32, c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "dinox_entry_weights": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_i64, c_f32, c_void_p, c_void_p]),
    "dinox_fill_f32": (c_int, [c_void_p, c_i64, c_f32, c_void_p]),
    "dinox_scalar_combine": (c_int, [c_void_p, c_void_p, c_int, c_f32, c_void_p, c_void_p, c_void_p]),
    "dinox_scalar_fanout": (c_int, [c_void_p, c_void_p, c_int, c_f32, c_void_p, c_void_p]),
    "dinox_split_bf16": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_void_p, c_i64, c_void_p]),
    "dinox_gelu_fwd_f32": (c_int, [c_void_p, c_i64, c_void_p, c_void_p]),
    "dinox_gelu_bwd_f32_workspace_bytes": (c_size, [c_i64, c_i64]),
    "dinox_gelu_bwd_f32": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dinox_normalize_tokens_f32": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_void_p, c_void_p, c_void_p]),
    "dinox_sqdiff_workspace_bytes": (c_size, []),
    "dinox_sqdiff_f32": (c_int, [c_void_p, c_void_p, c_i64, c_f32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dinox_adamw_plan_create": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dinox_adamw_plan_destroy": (c_int, [c_void_p]),
    "dinox_adamw_step": (c_int, [c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                 ctypes.c_double, ctypes.c_double, c_f32,
                                 c_void_p, c_void_p]),
    "dinox_koleo_candidates": (c_int, []),
    "dinox_koleo_rownorm": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_void_p, c_i64, c_void_p]),
    "dinox_koleo_fwd": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_void_p, c_i64, c_f32, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "dinox_koleo_bwd": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_f32, c_void_p, c_void_p,
                                c_i64, c_void_p]),
}


def _declare(l: ctypes.CDLL) -> None:
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().dinox_last_error_string().decode("utf-8", "replace")
        raise DinoxError(f"{what or 'dinox call'} failed (code {rc}): {msg}")


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args), name)


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", p)
