"""ctypes binding of libcds_b200.so (the C ABI declared in include/cds.h)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class CdsError(RuntimeError):
    pass


def lib_path():
    return os.path.join(HERE, "libcds_b200.so")


def _declare(L):
    c_p = C.c_void_p
    i32, i64, f64, f32 = C.c_int, C.c_long, C.c_double, C.c_float
    sz = C.c_size_t
    sig = {
        "cds_last_error": (C.c_char_p, []),
        "cds_version": (i32, []),
        "cds_device_ok": (i32, []),
        "cds_take_launch_count": (i64, []),
        "cds_contrast5": (i32, [c_p, i32, i32, c_p, i32, i32, i32, i32, c_p]),
        "cds_sudoku_contrast": (i32, [c_p, i32, i32, f64, f64, f64, c_p]),
        "cds_cu_contrast5": (i32, [c_p, i32, i32, i32, f64, f64, c_p]),
        "cds_contrast_sep_dev": (i32, [c_p, c_p, c_p, i32, i32, i32, i32, f32, c_p]),
        "cds_getframe": (i32, [i32, i32, c_p, c_p, sz, c_p, c_p, c_p]),
        "cds_syntax_pack_ref": (i64, [i32, c_p, c_p, c_p, c_p, c_p]),
        "cds_getframe_dev": (i32, [i32, i32, i32, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
        "cds_svm_model_create": (i32, [i32, i32, c_p, c_p, f64, f64, i32, i32, C.POINTER(c_p)]),
        "cds_svm_model_load": (i32, [C.c_char_p, C.POINTER(c_p)]),
        "cds_svm_model_free": (None, [c_p]),
        "cds_svm_model_nsv": (i32, [c_p]),
        "cds_svm_model_nfeat": (i32, [c_p]),
        "cds_svm_rbf_predict": (i32, [c_p, c_p, i32, i32, c_p, c_p]),
        "cds_svm_rbf_predict_dev": (i32, [c_p, c_p, i64, c_p, c_p]),
        "cds_camera_motion": (i32, [i32, i32, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
        "cds_mv_vote": (i32, [c_p, c_p, i32, c_p, c_p]),
        "cds_magnitude_vote": (i32, [c_p, i32, c_p]),
        "cds_create": (i32, [c_p, C.POINTER(c_p)]),
        "cds_destroy": (None, [c_p]),
        "cds_acquire_frame_buffers": (i32, [c_p, c_p, c_p, c_p, c_p]),
        "cds_submit_frame": (i32, [c_p, i32, c_p, c_p, c_p, sz]),
        "cds_get_map": (i32, [c_p, i32, c_p, c_p]),
        "cds_flush": (i32, [c_p]),
        "cds_clip_process_dev": (i32, [c_p, i32, i32, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
        "cds_clip_process": (i32, [c_p, i32, i32, c_p, c_p, c_p, c_p, sz, c_p, c_p]),
        "cds_clip_process_maps": (i32, [c_p, i32, i32, c_p, c_p, c_p, c_p, c_p, c_p]),
        "cds_getframe_semantic": (i32, [i32, i32, i32, c_p, c_p, c_p, c_p, c_p, c_p]),
        "cds_getframe_semantic_dev": (i32, [i32, i32, i32, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
        "cds_probe_seq_sum": (i32, [c_p, c_p, i32, c_p, c_p]),
        "cds_probe_run_sum": (i32, [c_p, i32, i32, i32, i32, i32, C.c_double, c_p, c_p]),
        "cds_reset": (i32, [c_p]),
        "cds_measure_peaks": (i32, [c_p, c_p, c_p]),
        "cds_measure_peaks_ex": (i32, [c_p, i32]),
        "cds_seek": (i32, [c_p, i32]),
        "cds_clip_context_dev": (i32, [c_p, i32, i32, c_p, c_p, c_p, c_p, c_p]),
        "cds_clip_context": (i32, [c_p, i32, i32, c_p, c_p, c_p, c_p, sz]),
        "cds_profile_stage_names": (C.c_char_p, []),
        "cds_profile_enable": (i32, [c_p, i32]),
        "cds_profile_read": (i32, [c_p, c_p, c_p, i32]),
        "cds_debug_copy": (i32, [c_p, C.c_char_p, i32, c_p, sz]),
        "cds_probe_mix": (i32, [i32, i32, c_p, c_p]),
        "cds_probe_spin": (i32, [i32, i32, i32, i32, c_p]),
        "cds_probe_umma": (i32, [c_p, i32, c_p, i32, C.c_uint, C.c_uint, C.c_uint, C.c_uint, i32, i32, c_p]),
    }
    missing = []
    for name, (res, args) in sig.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    return missing


def lib():
    """Load libcds_b200.so.  Fails loudly when the CUDA library has not been built."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise CdsError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        L = C.CDLL(p)
        L._cds_missing = _declare(L)   # tests/test_abi.py requires this to be empty
        _LIB = L
    return _LIB


def check(rc, what=""):
    if rc < 0:
        raise CdsError(f"{what} failed ({rc}): {lib().cds_last_error().decode()}")
    return rc


def device_ok():
    return bool(lib().cds_device_ok())


def take_launch_count():
    return int(lib().cds_take_launch_count())
