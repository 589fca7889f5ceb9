"""ctypes binding of libsivae_b200.so (include/sivae.h).  There is NO fallback: if the shared library is missing
or a call fails, a RuntimeError is raised."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SIVAE_LIB_PATH: load another build of the same library (A/B timing of kernel variants on one GPU box)
LIB_PATH = os.environ.get("SIVAE_LIB_PATH") or os.path.join(_HERE, "libsivae_b200.so")

NET_ENCODER, NET_DECODER, NET_TARGET = 0, 1, 2
CONV_AUTO, CONV_SIMT, CONV_TCGEN05, CONV_TC3X, CONV_TF32 = 0, 1, 2, 3, 4
T_CONV, T_BN_WEIGHT, T_BN_BIAS, T_LINEAR, T_BIAS = 0, 1, 2, 3, 4
LOSS_MSE, LOSS_L1, LOSS_BCE = 0, 1, 2
LOSS_TYPES = {"mse": LOSS_MSE, "l1": LOSS_L1, "bce": LOSS_BCE}      # recon_loss_type strings of the reference (:268-294)


class Config(C.Structure):
    _fields_ = [("cdim", C.c_int), ("zdim", C.c_int), ("image_size", C.c_int), ("n_channels", C.c_int),
                ("channels", C.c_int * 16), ("max_batch", C.c_int), ("variant", C.c_int), ("conv_backend", C.c_int),
                ("cond_dim", C.c_int)]


class Hyper(C.Structure):
    _fields_ = [("beta_kl", C.c_float), ("beta_rec", C.c_float), ("beta_neg", C.c_float), ("gamma_r", C.c_float),
                ("scale", C.c_float)]


class TensorInfo(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("kind", C.c_int), ("offset", C.c_longlong), ("numel", C.c_longlong),
                ("shape", C.c_int * 4), ("ndim", C.c_int)]


class BnInfo(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("channels", C.c_int), ("bn_offset", C.c_longlong), ("index", C.c_int)]


_P = C.c_void_p
_SIGS = {
    "sivae_last_error": (C.c_char_p, []),
    "sivae_version": (C.c_int, []),
    "sivae_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "sivae_destroy": (None, [_P]),
    "sivae_num_tensors": (C.c_int, [_P, C.c_int]),
    "sivae_tensor": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(TensorInfo)]),
    "sivae_param_count": (C.c_longlong, [_P, C.c_int]),
    "sivae_num_bn": (C.c_int, [_P, C.c_int]),
    "sivae_bn": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(BnInfo)]),
    "sivae_bn_floats": (C.c_longlong, [_P, C.c_int]),
    "sivae_workspace_bytes": (C.c_longlong, [_P]),
    "sivae_bind_net": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P]),
    "sivae_bind_workspace": (C.c_int, [_P, _P, C.c_longlong]),
    "sivae_params_changed": (C.c_int, [_P, C.c_int]),
    "sivae_e_step": (C.c_int, [_P, _P, _P, _P, C.c_int, C.POINTER(Hyper), _P, _P]),
    "sivae_d_step": (C.c_int, [_P, _P, C.POINTER(Hyper), _P, _P]),
    "sivae_vae_step": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(Hyper), _P, _P]),
    "sivae_set_reuse_decoder_passes": (C.c_int, [_P, C.c_int]),
    "sivae_get_reuse_decoder_passes": (C.c_int, [_P]),
    "sivae_set_recon_loss": (C.c_int, [_P, C.c_int]),
    "sivae_get_recon_loss": (C.c_int, [_P]),
    "sivae_adam_step": (C.c_int, [_P, C.c_int, C.c_float, C.c_float, _P]),
    "sivae_adam_set_step": (C.c_int, [_P, C.c_int, C.c_longlong]),
    "sivae_adam_get_step": (C.c_longlong, [_P, C.c_int]),
    "sivae_comm_unique_id": (C.c_int, [_P]),
    "sivae_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "sivae_comm_global_world": (C.c_int, []),
    "sivae_comm_finalize": (C.c_int, []),
    "sivae_allreduce_attach": (C.c_int, [_P, _P]),
    "sivae_comm_world": (C.c_int, [_P]),
    "sivae_allreduce_grads": (C.c_int, [_P, C.c_int, _P]),
    "sivae_iteration": (C.c_int, [_P, _P, _P, _P, C.c_int, C.POINTER(Hyper), C.c_float, C.c_float, _P, _P]),
    "sivae_encode": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P]),
    "sivae_decode": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_int, _P]),
    "sivae_encode_cond": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, C.c_int, _P]),
    "sivae_decode_cond": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P]),
    "sivae_launch_count": (C.c_ulonglong, []),
    "sivae_profile_enable": (C.c_int, [C.c_int]),
    "sivae_profile_read": (C.c_int, [C.POINTER(C.c_double)]),
    "sivae_profile_dump": (C.c_int, [C.c_char_p, C.c_int]),
    "sivae_last_image": (C.c_int, [_P, C.c_int, _P, _P]),
    "sivae_last_batch": (C.c_int, [_P]),
    "sivae_set_last_batch": (C.c_int, [_P, C.c_int]),
    "sivae_conv2d_fwd": (C.c_int, [_P, _P, _P, _P, _P] + [C.c_int] * 7 + [_P]),
    "sivae_conv2d_dgrad": (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 7 + [_P, C.c_longlong, _P]),
    "sivae_conv2d_wgrad": (C.c_int, [_P, _P, _P] + [C.c_int] * 8 + [_P, C.c_longlong, _P]),
    "sivae_bn_act_fwd": (C.c_int, [_P] * 9 + [C.c_int] * 6 + [_P, C.c_longlong, _P]),
    "sivae_bn_act_bwd": (C.c_int, [_P] * 10 + [C.c_int] * 6 + [_P, C.c_longlong, _P]),
    "sivae_bn_act_fwd_m": (C.c_int, [_P] * 9 + [C.c_int] * 6 + [_P, C.c_longlong, _P, _P]),
    "sivae_bn_act_bwd_m": (C.c_int, [_P] * 10 + [C.c_int] * 6 + [_P, C.c_longlong, _P, _P]),
    "sivae_mse3": (C.c_int, [_P] * 6 + [C.c_int, C.c_longlong, _P, C.c_longlong, _P]),
    "sivae_kl_reparam": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "sivae_adam_flat": (C.c_int, [_P, _P, _P, _P, C.c_longlong, C.c_float, C.c_float, C.c_longlong, _P]),
    "sivae_linear_fwd": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "sivae_linear_dgrad": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_longlong, _P]),
    "sivae_linear_dgrad_workspace_bytes": (C.c_longlong, [C.c_int, C.c_int, C.c_int]),
    "sivae_resample_coeffs": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), _P, _P, C.c_longlong]),
    "sivae_image_plan_bytes": (C.c_longlong, [C.c_int] * 4),
    "sivae_image_plan_init": (C.c_int, [C.c_int] * 4 + [_P, C.c_longlong, _P]),
    "sivae_image_batch_u8": (C.c_int, [_P, _P] + [C.c_int] * 6 + [_P, _P, _P]),
    "sivae_image_batch_u8_ex": (C.c_int, [_P, _P, _P] + [C.c_int] * 8 + [_P, _P, _P, _P]),
    "sivae_jpeg_info": (C.c_int, [_P, C.c_longlong, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sivae_jpeg_decode_batch": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built -- by design there is no
    Python/torch fallback for the hot path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libsivae_b200.so not found at %s -- build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "The B200 engine has no CPU / PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what=""):
    if code != 0:
        msg = load().sivae_last_error()
        raise RuntimeError("libsivae_b200: %s failed (code %d): %s" % (what, code, msg.decode() if msg else "?"))


def ptr(t):
    """raw device pointer of a torch tensor (or None)"""
    return None if t is None else C.c_void_p(t.data_ptr())
