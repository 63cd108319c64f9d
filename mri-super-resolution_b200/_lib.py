"""ctypes binding of the C ABI in include/b200inr.h (libb200inr.so).

The library is the product path: there is no CPU or PyTorch fallback.  Loading fails loudly when the shared
object has not been built (``python -c "import __graft_entry__ as g; g.build()"``), and every call raises
``RuntimeError`` on a non-zero return code.  Only raw device pointers, sizes and the CUDA stream handle cross the
boundary; torch is used by the callers for allocation and stream management only.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200INR_LIB") or os.path.join(_HERE, "lib", "libb200inr.so")  # (override: kernel tuning builds)

MAX_TAPS = 8
ACT_SINE, ACT_RELU, ACT_GABOR, ACT_TANH = 0, 1, 2, 3
IN_COORDS, IN_FOURIER, IN_FEATURES = 0, 1, 2
DEFAULT_PIPED_BWD = "1"  # raw-coordinate SIRENs train through the pipelined backward (B200INR_PIPED_BWD=0: staged)
NET_RELU_TAIL = 2   # B200INR_NET_RELU_TAIL: last hidden layer Linear + ReLU, ReLU on the output (INR/INR_ERD.py:28-67)
NET_DGRAD_ONLY = 4  # B200INR_NET_DGRAD_ONLY: generic family, frozen network: stash only what the dgrad chain reads
NET_TANH_OUT = 8    # B200INR_NET_TANH_OUT: generic family, out = scale_0 * tanh(final linear) (PN)
NET_STAGED_BWD = 1  # B200INR_NET_STAGED_BWD: the older forward-stash / dgrad / wgrad training path of raw-coordinate SIRENs


class Net(ctypes.Structure):
    """b200inr_net == ctor arguments of Siren (reference INR/SRDWI.py:68-71)."""

    _fields_ = [
        ("in_features", ctypes.c_int32),
        ("hidden_features", ctypes.c_int32),
        ("hidden_layers", ctypes.c_int32),
        ("out_features", ctypes.c_int32),
        ("first_omega_0", ctypes.c_float),
        ("hidden_omega_0", ctypes.c_float),
        ("activation", ctypes.c_int32),
        ("input_mode", ctypes.c_int32),
        ("mapping_size", ctypes.c_int32),
        ("scale_0", ctypes.c_float),
        ("flags", ctypes.c_int32),
    ]


class Grid(ctypes.Structure):
    """b200inr_grid == get_mgrid(shape) rows [row_begin, row_begin + rows) (reference INR/SRDWI.py:12-18)."""

    _fields_ = [("ndim", ctypes.c_int32), ("shape", ctypes.c_int32 * 4), ("row_begin", ctypes.c_int64)]


class AxisTaps(ctypes.Structure):
    _fields_ = [("idx", ctypes.c_int32 * MAX_TAPS), ("w", ctypes.c_float * MAX_TAPS)]


_vp, _i32, _i64, _f32, _f64, _sz = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float,
                                    ctypes.c_double, ctypes.c_size_t)
_P = ctypes.POINTER

# name -> (restype, argtypes); every symbol include/b200inr.h declares
SIGNATURES = {
    "b200inr_version": (ctypes.c_char_p, []),
    "b200inr_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "b200inr_param_count": (ctypes.c_int, [_P(Net), _P(_i64)]),
    "b200inr_param_offsets": (ctypes.c_int, [_P(Net), _P(_i64)]),
    "b200inr_packed_bytes": (ctypes.c_int, [_P(Net), _P(_sz)]),
    "b200inr_pack_weights": (ctypes.c_int, [_P(Net), _vp, _vp, _vp]),
    "b200inr_stash_bytes": (ctypes.c_int, [_P(Net), _i64, _P(_sz)]),
    "b200inr_siren_forward": (ctypes.c_int, [_P(Net), _vp, _vp, _P(Grid), _i64, _vp, ctypes.c_int, _f32, _vp, _vp]),
    "b200inr_siren_forward_pool_loss": (ctypes.c_int, [_P(Net), _vp, _P(Grid), _i64, _vp, _f64, _vp, _vp, _vp, _vp]),
    "b200inr_siren_backward": (ctypes.c_int, [_P(Net), _vp, _vp, _vp, _P(Grid), _i64, _vp, _vp, _vp]),
    "b200inr_siren_backward_input": (ctypes.c_int, [_P(Net), _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "b200inr_siren_backward_coords": (ctypes.c_int, [_P(Net), _vp, _vp, _vp, _P(Grid), _i64, _vp, _vp, _vp, _vp]),
    "b200inr_siren_backward_tanh_out": (ctypes.c_int, [_P(Net), _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "b200inr_pn_effective_params": (ctypes.c_int, [_vp, _i64, _i64, _i32, _f32, _vp, _vp, _vp]),
    "b200inr_pn_fold_grad": (ctypes.c_int, [_vp, _i64, _i64, _i32, _f32, _vp]),
    "b200inr_siren_dgrad": (ctypes.c_int, [_P(Net), _vp, _vp, _i64, _vp, _vp]),
    "b200inr_siren_wgrad": (ctypes.c_int, [_P(Net), _vp, _vp, _P(Grid), _i64, _vp, _vp]),
    "b200inr_mse_loss": (ctypes.c_int, [_vp, _vp, _vp, _i64, _f64, _vp, _vp, _vp]),
    "b200inr_mse_loss_relu_out": (ctypes.c_int, [_vp, _vp, _vp, _i64, _f64, _vp, _vp, _vp]),
    "b200inr_soft_erd": (ctypes.c_int, [_vp, _vp, _i64, _i32, _f64, _f64, _f64, _vp, _vp, _vp]),
    "b200inr_degrade_build_axis_host": (ctypes.c_int, [_i32, ctypes.c_int, _P(AxisTaps), _P(AxisTaps)]),
    "b200inr_degrade_forward": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp]),
    "b200inr_degrade_adjoint": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp]),
    "b200inr_degrade_build_band_host": (ctypes.c_int, [_i32, ctypes.c_int, _P(_f32), _P(_f32)]),
    "b200inr_blurpool_mse": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i64, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200inr_blurpool_mse_slab": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i64, _f64, _vp, _vp, _vp, _vp, _i32, _i32, _vp,
                                                 _vp, _vp, _vp]),
    "b200inr_pool_mse": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i64, _f64, _vp, _vp, _vp]),
    "b200inr_adam_step": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _vp, _vp]),
    "b200inr_optimizer_step": (ctypes.c_int, [_P(Net), _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "b200inr_optimizer_step_peers": (ctypes.c_int, [_P(Net), _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _f32, _f32, _f32,
                                                    _f32, _vp, _vp, _vp, _vp]),
    "b200inr_net_size": (_sz, []),
    "b200inr_param_offset_count": (ctypes.c_int, [_P(Net), _P(_i32)]),
    "b200inr_get_mgrid": (ctypes.c_int, [_P(Grid), _i64, _vp, _vp]),
    "b200inr_input_mapping": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "b200inr_sine_layer_pre": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _f32, _vp, _vp]),
    "b200inr_combinations": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "b200inr_adc_fit": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "b200inr_input_mapping_backward": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "b200inr_selftest_umma": (ctypes.c_int, [ctypes.c_int, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
}

_lib = None


def load():
    """Load libb200inr.so (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"b200inr: {LIB_PATH} is missing - build it with __graft_entry__.build(); there is no fallback path")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.b200inr_net_size() != ctypes.sizeof(Net):  # struct drift between this binding and the built library
        raise RuntimeError(f"b200inr: {LIB_PATH} was built for a {lib.b200inr_net_size()}-byte b200inr_net, this "
                           f"binding passes {ctypes.sizeof(Net)} bytes - rebuild with __graft_entry__.build()")
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().b200inr_error_string(code).decode()
        raise RuntimeError(f"b200inr: {what} failed: {msg} ({code})")


def make_net(in_features, hidden_features, hidden_layers, out_features, first_omega_0=30.0, hidden_omega_0=30.0,
             activation=ACT_SINE, input_mode=IN_COORDS, mapping_size=0, scale_0=0.0, flags=None):
    if flags is None:  # B200INR_PIPED_BWD=1/0 selects the pipelined / staged backward of SIRENs built afterwards (A/B runs)
        flags = 0 if os.environ.get("B200INR_PIPED_BWD", DEFAULT_PIPED_BWD) == "1" else NET_STAGED_BWD
    return Net(int(in_features), int(hidden_features), int(hidden_layers), int(out_features), float(first_omega_0),
               float(hidden_omega_0), int(activation), int(input_mode), int(mapping_size), float(scale_0), int(flags))


def make_grid(shape, row_begin=0):
    g = Grid()
    g.ndim = len(shape)
    for j in range(4):
        g.shape[j] = int(shape[j]) if j < len(shape) else 1
    g.row_begin = int(row_begin)
    return g


def param_count(net):
    n = _i64(0)
    check(load().b200inr_param_count(ctypes.byref(net), ctypes.byref(n)), "param_count")
    return n.value


def param_offsets(net):
    cnt = _i32(0)
    check(load().b200inr_param_offset_count(ctypes.byref(net), ctypes.byref(cnt)), "param_offset_count")
    off = (_i64 * cnt.value)()
    check(load().b200inr_param_offsets(ctypes.byref(net), off), "param_offsets")
    return list(off)


def packed_bytes(net):
    n = _sz(0)
    check(load().b200inr_packed_bytes(ctypes.byref(net), ctypes.byref(n)), "packed_bytes")
    return n.value


def stash_bytes(net, rows):
    n = _sz(0)
    check(load().b200inr_stash_bytes(ctypes.byref(net), int(rows), ctypes.byref(n)), "stash_bytes")
    return n.value


def build_band_tables(n_hr, blur):
    """(fwd6 [n_hr/2, 6], adj3 [n_hr, 3]) float32 NumPy arrays: the banded form of the degradation along one axis."""
    import numpy as np
    fwd = np.zeros((n_hr // 2, 6), dtype=np.float32)
    adj = np.zeros((n_hr, 3), dtype=np.float32)
    check(load().b200inr_degrade_build_band_host(int(n_hr), int(bool(blur)), fwd.ctypes.data_as(_P(_f32)),
                                                 adj.ctypes.data_as(_P(_f32))), "degrade_build_band_host")
    return fwd, adj


def build_axis_taps(n_hr, blur):
    fwd = (AxisTaps * (n_hr // 2))()
    adj = (AxisTaps * n_hr)()
    check(load().b200inr_degrade_build_axis_host(int(n_hr), int(bool(blur)), fwd, adj), "degrade_build_axis_host")
    return fwd, adj
