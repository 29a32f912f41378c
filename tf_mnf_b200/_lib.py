"""ctypes binding of libmnf_b200.so (include/mnf_b200.h).  No fallback of any kind: if the
shared library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libmnf_b200.so")

c_f32p = C.POINTER(C.c_float)
vp = C.c_void_p


class Noise(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("row_offset", C.c_uint64), ("stream", C.c_uint32),
                ("reserved", C.c_uint32), ("injected", vp), ("call_delta", vp)]


class FlowStep(C.Structure):
    _fields_ = [("kind", C.c_int32), ("parity", C.c_int32), ("hidden", C.c_int32), ("reserved", C.c_int32),
                ("w1", vp), ("b1", vp), ("ws", vp), ("bs", vp), ("wt", vp), ("bt", vp),
                ("ld2", C.c_int64), ("m1", vp), ("m2", vp), ("bern", Noise),
                ("n_extra", C.c_int32), ("activation", C.c_int32), ("extra_h", C.c_int32 * 3), ("reserved2", C.c_int32),
                ("extra_w", vp * 3), ("extra_b", vp * 3), ("extra_m", vp * 3)]


class ConvShape(C.Structure):
    _fields_ = [("B", C.c_int64), ("H", C.c_int64), ("W", C.c_int64), ("C", C.c_int64),
                ("kh", C.c_int64), ("kw", C.c_int64), ("F", C.c_int64), ("stride", C.c_int64),
                ("same", C.c_int32), ("fuse_relu_pool", C.c_int32)]


class McSource(C.Structure):
    _fields_ = [("n_rep", C.c_int64), ("Ho", C.c_int64), ("Wo", C.c_int64), ("F", C.c_int64),
                ("mean", vp), ("sd", vp), ("z", vp), ("eps", Noise)]


class KlArgs(C.Structure):
    _fields_ = [("conv", C.c_int32), ("use_z", C.c_int32), ("r_all_states", C.c_int32), ("n_r_states", C.c_int32),
                ("rows", C.c_int64), ("cols", C.c_int64),
                ("mean_W", vp), ("log_std_W", vp), ("mean_b", vp), ("log_var_b", vp),
                ("prior_var_r_p", vp), ("prior_var_r_p_bias", vp),
                ("q0_log_var", vp), ("r0_mean", vp), ("r0_log_var", vp), ("r0_apvar", vp),
                ("z", vp), ("ld_q", vp), ("r_states", vp), ("ld_r", vp),
                ("eps_a", Noise), ("eps_b", Noise)]


class KlGrads(C.Structure):
    _fields_ = [("mean_W", vp), ("log_std_W", vp), ("mean_b", vp), ("log_var_b", vp),
                ("prior_var_r_p", vp), ("prior_var_r_p_bias", vp),
                ("q0_log_var", vp), ("r0_mean", vp), ("r0_log_var", vp), ("r0_apvar", vp),
                ("d_z", vp), ("d_r_states", vp)]


class TensorView(C.Structure):
    _fields_ = [("data", vp), ("ndim", C.c_int32), ("shape", C.c_int64 * 6), ("device_type", C.c_int32),
                ("device_id", C.c_int32), ("dtype_code", C.c_int32), ("dtype_bits", C.c_int32),
                ("contiguous", C.c_int32)]


RELEASE_FN = C.CFUNCTYPE(None, vp)

MNF_FLOW_IAF, MNF_FLOW_RNVP, MNF_FLOW_PLANAR = 0, 1, 2
MNF_FLOW_MAX_HIDDEN = 4
ACTIVATIONS = {None: 0, "default": 0, "relu": 1, "tanh": 2, "sigmoid": 3, "linear": 4}

# name -> (argtypes); every function returns int status except the two noted below
_i64, _i32, _f, _u64, _u32, _sz = C.c_int64, C.c_int32, C.c_float, C.c_uint64, C.c_uint32, C.c_size_t
SIGNATURES = {
    "mnf_device_count": [C.POINTER(C.c_int)],
    "mnf_ctx_create": [C.c_int, vp, C.POINTER(vp)],
    "mnf_ctx_fork": [vp, _i32, C.POINTER(vp)],
    "mnf_ctx_destroy": [vp],
    "mnf_ctx_set_stream": [vp, vp],
    "mnf_ctx_get_stream": [vp, C.POINTER(vp)],
    "mnf_ctx_sync": [vp],
    "mnf_ctx_wait": [vp, vp],
    "mnf_ctx_wait_stream": [vp, vp],
    "mnf_ctx_sm_count": [vp, C.POINTER(C.c_int)],
    "mnf_ctx_launch_count": [vp, C.POINTER(_i64)],
    "mnf_ctx_set_math": [vp, C.c_int],
    "mnf_ctx_freeze_weights": [vp, C.c_int32],
    "mnf_ctx_probe": [vp, C.c_int, C.c_int, vp, vp],
    "mnf_malloc": [vp, _sz, C.POINTER(vp)],
    "mnf_free": [vp, vp],
    "mnf_host_alloc": [_sz, C.POINTER(vp)],
    "mnf_host_free": [vp],
    "mnf_memcpy_h2d": [vp, vp, vp, _sz],
    "mnf_memcpy_d2h": [vp, vp, vp, _sz],
    "mnf_memcpy_d2d": [vp, vp, vp, _sz],
    "mnf_memset": [vp, vp, C.c_int, _sz],
    "mnf_event_create": [C.POINTER(vp)],
    "mnf_event_destroy": [vp],
    "mnf_event_record": [vp, vp],
    "mnf_event_elapsed_ms": [vp, vp, C.POINTER(_f)],
    "mnf_graph_begin": [vp],
    "mnf_graph_end": [vp, C.POINTER(vp)],
    "mnf_graph_launch": [vp, vp],
    "mnf_graph_destroy": [vp],
    "mnf_dlpack_view": [vp, C.POINTER(TensorView)],
    "mnf_dlpack_wrap": [vp, _i32, C.POINTER(_i64), _i32, RELEASE_FN, vp, C.POINTER(vp)],
    "mnf_dlpack_delete": [vp],
    "mnf_philox_normal": [vp, _u64, _u32, _u64, _i64, _i64, _i64, vp],
    "mnf_philox_bernoulli": [vp, _u64, _u32, _u64, _i64, _i64, _f, vp],
    "mnf_philox_raw": [vp, _u64, _u32, _u64, _i64, _i64, _i64, vp],
    "mnf_flow_forward": [vp, _i64, _i64, _i32, C.POINTER(FlowStep), vp, vp, vp, C.POINTER(Noise), vp, vp, vp],
    "mnf_flow_backward": [vp, _i64, _i64, _i32, C.POINTER(FlowStep), vp, vp, vp, vp, vp, vp, C.POINTER(FlowStep)],
    "mnf_sample_z_backward": [vp, _i64, _i64, vp, vp, C.POINTER(Noise), vp, vp],
    "mnf_maf_decode": [vp, _i64, _i64, C.POINTER(FlowStep), vp, vp, vp],
    "mnf_dense_forward": [vp, _i64, _i64, _i64, vp, vp, vp, vp, vp, vp, _f, C.POINTER(Noise), vp, vp],
    "mnf_dense_forward_act": [vp, _i64, _i64, _i64, vp, vp, vp, vp, vp, vp, _f, C.POINTER(Noise), C.c_int32, vp, vp],
    "mnf_dense_can_fuse_act": [vp, _i64, _i64, C.c_int32, C.POINTER(C.c_int32)],
    "mnf_dense_backward": [vp, _i64, _i64, _i64, vp, vp, vp, vp, vp, _f, C.POINTER(Noise), vp, vp,
                           vp, vp, vp, vp, vp, vp],
    "mnf_conv_out_hw": [C.POINTER(ConvShape), C.POINTER(_i64), C.POINTER(_i64)],
    "mnf_conv2d_can_fuse_pool": [vp, C.POINTER(ConvShape), C.POINTER(_i32)],
    "mnf_conv2d_forward": [vp, C.POINTER(ConvShape), vp, vp, vp, vp, vp, vp, _f, C.POINTER(Noise), vp, vp, vp],
    "mnf_conv2d_backward": [vp, C.POINTER(ConvShape), vp, vp, vp, vp, vp, _f, C.POINTER(Noise), vp, vp, vp,
                            vp, vp, vp, vp, vp, vp],
    "mnf_kl_forward": [vp, C.POINTER(KlArgs), vp],
    "mnf_kl_backward": [vp, C.POINTER(KlArgs), _f, C.POINTER(KlGrads)],
    "mnf_relu_maxpool2": [vp, _i64, _i64, _i64, _i64, vp, vp, vp],
    "mnf_relu_maxpool2_backward": [vp, _i64, _i64, _i64, _i64, vp, vp, vp, vp],
    "mnf_relu": [vp, _i64, vp, vp],
    "mnf_relu_backward": [vp, _i64, vp, vp, vp],
    "mnf_softmax": [vp, _i64, _i64, vp, vp],
    "mnf_softmax_xent": [vp, _i64, _i64, vp, vp, _f, vp, vp],
    "mnf_mse_loss": [vp, _i64, vp, vp, _f, vp, vp],
    "mnf_scale": [vp, _i64, _f, vp, vp],
    "mnf_batchnorm_forward": [vp, _i64, _i64, vp, vp, vp, vp, vp, _f, _f, _i32, vp, vp, vp],
    "mnf_batchnorm_backward": [vp, _i64, _i64, vp, vp, vp, vp, _i32, vp, vp, vp, vp],
    "mnf_axpy": [vp, _i64, _f, vp, vp],
    "mnf_tile_rows": [vp, _i64, _i64, _i64, vp, vp],
    "mnf_conv2d_mc_maps": [vp, C.POINTER(ConvShape), vp, vp, vp, vp, vp, _f, vp, vp],
    "mnf_conv2d_mc_can_expand": [vp, _i64, _i64, _i64, _i32, C.POINTER(_i32)],
    "mnf_conv2d_mc_expand": [vp, _i64, _i64, _i64, _i64, _i64, vp, vp, vp, C.POINTER(Noise), _i32, vp],
    "mnf_conv2d_forward_mc": [vp, C.POINTER(ConvShape), C.POINTER(McSource), vp, vp, vp, vp, vp, C.c_float, C.POINTER(Noise), vp,
                              C.POINTER(_i32)],
    "mnf_u32_add": [vp, vp, C.c_uint32],
    "mnf_adam_step": [vp, _i64, vp, vp, vp, vp, _f, _f, _f, _f, _i32],
    "mnf_adam_step_dev": [vp, _i64, vp, vp, vp, vp, _f, _f, _f, _f, vp],
    "mnf_ctx_set_capturing": [vp, _i32],
    "mnf_comm_nccl_version": [C.POINTER(C.c_int)],
    "mnf_comm_unique_id": [vp],
    "mnf_comm_init": [vp, vp, C.c_int, C.c_int, C.POINTER(vp)],
    "mnf_comm_info": [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "mnf_comm_destroy": [vp],
    "mnf_allreduce_grads": [vp, vp, vp, _i64],
}
MNF_COMM_ID_BYTES = 128
MNF_ACT_NONE, MNF_ACT_RELU, MNF_ACT_SOFTMAX = 0, 1, 2
NON_STATUS = {"mnf_version": ([], C.c_int), "mnf_last_error": ([], C.c_char_p), "mnf_last_path": ([], C.c_char_p)}
ALL_SYMBOLS = sorted([*SIGNATURES, *NON_STATUS])


class MnfError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libmnf_b200.so (building is __graft_entry__.build()'s / _build.py's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library has not been built "
            "(run `python -m tf_mnf_b200._build`).  tf_mnf_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    for name, (args, res) in NON_STATUS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def last_path() -> str:
    """Kernel family that served this thread's last dense / conv / flow forward call."""
    return load().mnf_last_path().decode()


def check(rc: int) -> None:
    if rc == 0:
        return
    msg = load().mnf_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise NotImplementedError(msg)
    if rc == -4:
        raise ValueError(msg)
    raise MnfError(msg)
