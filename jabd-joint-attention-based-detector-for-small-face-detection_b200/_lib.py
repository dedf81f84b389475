"""ctypes binding of ``libjabd_b200.so`` (C-ABI declared in ``include/jabd_b200.h``).

The library is the product: if it is missing or does not export a declared
symbol the import of any operator module fails loudly -- there is no Python or
CPU fallback anywhere in this package.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# JABD_B200_LIB selects another build of the same library (kernel-variant experiments); never a different backend
SO_PATH = os.environ.get("JABD_B200_LIB") or os.path.join(_HERE, "libjabd_b200.so")
CSRC = os.path.join(_HERE, "csrc")

c_int, c_i64, c_f32, c_f64, c_sz, c_vp = (ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double,
                                          ctypes.c_size_t, ctypes.c_void_p)

class AssignBatch(ctypes.Structure):
    """``jabd_assign_batch_t`` of include/jabd_b200.h (one batch of ``jabd_assign_batches``)."""
    _fields_ = [("gt", c_vp), ("gt_off", c_vp), ("B", c_int), ("sumG", c_i64), ("loc_t", c_vp), ("conf_t", c_vp),
                ("landm_t", c_vp), ("workspace", c_vp), ("workspace_bytes", c_sz)]


class DetectBatch(ctypes.Structure):
    """``jabd_detect_batch_t`` of include/jabd_b200.h (one batch of ``jabd_detect_batches``)."""
    _fields_ = [("loc", c_vp), ("conf", c_vp), ("landm", c_vp), ("B", c_int), ("dets", c_vp), ("counts", c_vp), ("keep_idx", c_vp),
                ("workspace", c_vp), ("workspace_bytes", c_sz)]


# name -> (restype, argtypes); mirrors include/jabd_b200.h one to one
SIGNATURES = {
    "jabd_version": (c_int, []),
    "jabd_last_error": (ctypes.c_char_p, []),
    "jabd_device_info": (c_int, [c_vp, c_vp, c_vp]),
    "jabd_priors_count": (c_i64, [c_vp, c_vp, c_int, c_int, c_int]),
    "jabd_priors": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_i64, c_vp]),
    "jabd_point_form": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "jabd_jaccard": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "jabd_intersect": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "jabd_encode": (c_int, [c_vp, c_vp, c_i64, c_f32, c_f32, c_vp, c_vp]),
    "jabd_encode_landm": (c_int, [c_vp, c_vp, c_i64, c_f32, c_vp, c_vp]),
    "jabd_decode": (c_int, [c_vp, c_vp, c_i64, c_int, c_f32, c_f32, c_vp, c_vp]),
    "jabd_decode_landm": (c_int, [c_vp, c_vp, c_i64, c_int, c_f32, c_vp, c_vp]),
    "jabd_assign_workspace_bytes": (c_sz, [c_int, c_i64, c_i64]),
    "jabd_assign": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_i64, c_f32, c_f32, c_f32, c_int, c_int, c_int,
                            c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "jabd_assign_batches": (c_int, [c_vp, c_i64, c_vp, c_int, c_f32, c_f32, c_f32, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "jabd_assign_match": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_sz, c_vp]),
    "jabd_assign_encode": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_i64, c_f32, c_f32, c_f32, c_int, c_int,
                                   c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "jabd_pack_gt_rows": (c_i64, [c_vp, c_vp, c_int, c_vp, c_i64, c_vp]),
    "jabd_assign_host_scratch_bytes": (c_sz, [c_int, c_i64, c_i64, c_int]),
    "jabd_assign_host_out_offsets": (c_int, [c_int, c_i64, c_int, c_vp]),
    "jabd_assign_host": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_f32, c_f32, c_f32, c_int, c_int, c_int,
                                 c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "jabd_correct_boxes": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    "jabd_multibox_loss_workspace_bytes": (c_sz, [c_int]),
    "jabd_multibox_loss_forward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "jabd_multibox_loss_backward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "jabd_bbox_overlaps_family": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp]),
    "jabd_iou_loss_forward": (c_int, [c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_int, c_int, c_vp, c_vp, c_vp]),
    "jabd_iou_loss_backward": (c_int, [c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_int, c_int, c_vp, c_vp, c_vp]),
    "jabd_multibox_loss_forward_ex": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_int, c_vp, c_f32, c_f32,
                                              c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "jabd_multibox_loss_backward_ex": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_f32,
                                               c_f32, c_vp, c_vp, c_vp, c_vp]),
    "jabd_bbox_overlaps_f64": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "jabd_img_pr_info": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_vp, c_vp, c_sz, c_vp]),
    "jabd_norm_score": (c_int, [c_vp, c_i64, c_vp, c_sz, c_vp]),
    "jabd_wider_eval_workspace_bytes": (c_sz, [c_int, c_i64, c_i64, c_int]),
    "jabd_wider_eval": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_f64, c_int, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "jabd_topk_workspace_bytes": (c_sz, [c_int, c_i64, c_int]),
    "jabd_topk": (c_int, [c_vp, c_i64, c_i64, c_int, c_i64, c_f32, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "jabd_nms_workspace_bytes": (c_sz, [c_int, c_i64, c_int]),
    "jabd_nms_stats_offset": (c_sz, [c_int, c_int]),
    "jabd_nms": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_int, c_i64, c_f32, c_int, c_int, c_f64, c_int,
                         c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "jabd_diounms": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_int, c_i64, c_int, c_f64, c_f32, c_int, c_vp, c_vp, c_vp,
                             c_sz, c_vp]),
    "jabd_detect_workspace_bytes": (c_sz, [c_int, c_i64, c_int]),
    "jabd_detect": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_f32, c_f32, c_f32, c_int, c_int, c_f64, c_int, c_int,
                            c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "jabd_detect_batches": (c_int, [c_vp, c_i64, c_vp, c_int, c_f32, c_f32, c_f32, c_int, c_int, c_f64, c_int, c_int, c_vp, c_int, c_vp]),
    "jabd_detect_host_scratch_bytes": (c_sz, [c_int, c_i64, c_int, c_int]),
    "jabd_detect_host": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_f32, c_f32, c_f32, c_int, c_int, c_f64,
                                 c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "jabd_detect_host_async": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_f32, c_f32, c_f32, c_int, c_int, c_f64,
                                       c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "jabd_p2p_alloc": (c_int, [c_sz, c_vp, c_vp]),
    "jabd_p2p_open": (c_int, [c_vp, c_vp]),
    "jabd_p2p_close": (c_int, [c_vp]),
    "jabd_p2p_free": (c_int, [c_vp]),
    "jabd_p2p_allgather": (c_int, [c_vp, c_sz, c_vp, c_sz, c_vp, c_vp, c_vp, c_int, c_int, ctypes.c_uint64, ctypes.c_uint64, c_vp,
                                   c_f64, c_vp, c_vp]),
    "jabd_p2p_wait": (c_int, [c_vp, c_int, ctypes.c_uint64, c_f64, c_vp, c_vp]),
}

# libjabd_b200_selftest.so (include/jabd_b200_selftest.h): test / bench hooks, deliberately not in the product library
SELFTEST_SO_PATH = os.path.join(_HERE, "libjabd_b200_selftest.so")
SELFTEST_SIGNATURES = {
    "jabd_selftest_div": (c_int, [ctypes.c_uint64, ctypes.c_uint64, c_vp, c_vp, c_vp]),
    "jabd_selftest_fp32_probe": (c_int, [c_int, c_int, c_vp, c_vp]),
    "jabd_selftest_last_error": (ctypes.c_char_p, []),
}

ERRORS = {-1: ValueError, -2: ValueError, -3: ValueError, -4: RuntimeError, -5: RuntimeError}

_lib = None
_selftest = None


def build(force=False, verbose=False):
    """Compile ``libjabd_b200.so`` for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", "Makefile"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "jabd_b200.h"))
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "jabd_b200_selftest.h"))
    outs = (SO_PATH, SELFTEST_SO_PATH)
    stale = force or any(not os.path.exists(o) or any(os.path.getmtime(s) > os.path.getmtime(o) for s in srcs) for o in outs)
    if stale:
        cmd = ["make", "-C", CSRC, "-j4"] + ([] if verbose else ["-s"])
        if force:
            subprocess.check_call(["make", "-C", CSRC, "-s", "clean"])
        subprocess.check_call(cmd)
    return SO_PATH


def lib():
    """The loaded library; raises ``RuntimeError`` if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "libjabd_b200.so is missing at %s: build it with `python -c \"import __graft_entry__ as g; g.build()\"` "
                "(or `make -C %s`).  This package has no CPU or PyTorch fallback." % (SO_PATH, CSRC))
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def selftest_lib():
    """The test / bench hook library (never needed by the operators)."""
    global _selftest
    if _selftest is None:
        if not os.path.exists(SELFTEST_SO_PATH):
            raise RuntimeError("libjabd_b200_selftest.so is missing at %s: run `make -C %s`" % (SELFTEST_SO_PATH, CSRC))
        L = ctypes.CDLL(SELFTEST_SO_PATH)
        for name, (res, args) in SELFTEST_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _selftest = L
    return _selftest


def selftest_call(name, *args):
    L = selftest_lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise ERRORS.get(rc, RuntimeError)("%s failed (%d): %s" % (name, rc, L.jabd_selftest_last_error().decode("utf-8", "replace")))


def last_error():
    return lib().jabd_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise ERRORS.get(rc, RuntimeError)("%s failed (%d): %s" % (what or "libjabd_b200", rc, last_error()))


def call(name, *args):
    """Call an int-returning entry point and raise on a negative code."""
    check(getattr(lib(), name)(*args), name)
