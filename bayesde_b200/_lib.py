"""ctypes binding of libbsde.so (include/bsde.h).  There is NO fallback: if the library is missing
or a call fails, an exception is raised."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# BSDE_LIB_PATH: tools-only override (A/B timing of two builds on the same box); the default is the in-tree library
_LIB_PATH = os.environ.get("BSDE_LIB_PATH") or os.path.join(_HERE, "libbsde.so")

FW_MAX_MID = 4
FW_MAX_EDGE = 4
FW_MAX_HID = 256
MAX_LAYERS = 4

ACT = {"softplus": 0, "tanh": 1, "elu": 2, "none": 3, "rbf": 4, "swish": 5}
METHOD = {"euler_maruyama": 0, "milstein": 1, "euler_heun": 2}
MATH = {"fp32": 0, "tc": 1}
EPI_BIAS, EPI_BIAS_ACT, EPI_EULER, EPI_DACT, EPI_ACCUM = range(5)

_fp = C.POINTER(C.c_float)


class FwParams(C.Structure):
    _fields_ = ([("P", C.c_int32), ("nhid", C.c_int32), ("dims", C.c_int32 * (FW_MAX_MID + 2)), ("act", C.c_int32)]
                + [(n, C.c_void_p) for n in ("W1", "b1", "wt1", "wtb1", "bt1")]
                + [(n, C.c_void_p * FW_MAX_MID) for n in ("Wm", "bm", "wtm", "wtbm", "btm")]
                + [(n, C.c_void_p) for n in ("W4", "b4", "wt4", "wtb4", "bt4")]
                + [("beta1", C.c_void_p), ("betam", C.c_void_p * FW_MAX_MID), ("plain", C.c_int32)])


class WsdeDesc(C.Structure):
    _fields_ = [("S", C.c_int32), ("nit", C.c_int32), ("stl", C.c_int32), ("w0_per_sample", C.c_int32),
                ("sigma", C.c_float), ("kl_weight", C.c_float), ("ou_coef", C.c_float),
                ("seed", C.c_uint64), ("stream_id", C.c_uint32), ("sample_offset", C.c_uint32),
                ("d_dtkl", C.c_void_p), ("d_nscale", C.c_void_p), ("d_nidx", C.c_void_p), ("d_back", C.c_void_p),
                ("d_kltraj", C.c_void_p)]


class ConvLayer(C.Structure):
    _fields_ = [("cin", C.c_int32), ("cout", C.c_int32), ("ksize", C.c_int32),
                ("hin", C.c_int32), ("win", C.c_int32), ("hout", C.c_int32), ("wout", C.c_int32),
                ("sn", C.c_int32), ("kn", C.c_int32), ("off", C.c_int32), ("sd", C.c_int32),
                ("has_time", C.c_int32), ("time_first", C.c_int32),
                ("w_off", C.c_int64), ("b_off", C.c_int64),
                ("w_tap_stride", C.c_int64), ("w_k_stride", C.c_int64), ("w_n_stride", C.c_int64)]


class XpathDesc(C.Structure):
    _fields_ = [("S", C.c_int32), ("B", C.c_int32), ("nit", C.c_int32), ("nlayers", C.c_int32),
                ("act", C.c_int32), ("math", C.c_int32), ("P", C.c_int64), ("x_numel", C.c_int64),
                ("layers", ConvLayer * MAX_LAYERS), ("beta_off", C.c_int64 * MAX_LAYERS)]


class SDELayerDesc(C.Structure):
    _fields_ = [("S", C.c_int32), ("B", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("nt", C.c_int32), ("act", C.c_int32),
                ("driftw", C.c_int32), ("diffw", C.c_int32), ("prior_ou", C.c_int32), ("diff_drift", C.c_int32),
                ("stop_grad", C.c_int32), ("prior_solve", C.c_int32), ("w0_per_sample", C.c_int32), ("sigma", C.c_float),
                ("P", C.c_int64), ("seed", C.c_uint64), ("stream_id", C.c_uint32), ("sample_offset", C.c_uint32)]


class ToyDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("nt", C.c_int32), ("h1", C.c_int32), ("h2", C.c_int32),
                ("sigma", C.c_float), ("theta", C.c_float)]


class BsdeError(RuntimeError):
    pass


_lib = None

_EXPORTS = ["bsde_version", "bsde_last_error", "bsde_launch_count", "bsde_wsde_workspace_bytes", "bsde_wsde_fwd", "bsde_wsde_bwd",
            "bsde_tconv_fwd", "bsde_tconv_fwd_workspace_bytes", "bsde_tconv_wgrad_workspace_bytes", "bsde_tconv_wgrad",
            "bsde_xpath_workspace_bytes", "bsde_xpath_acts_bytes", "bsde_xpath_fwd", "bsde_xpath_bwd", "bsde_step_update",
            "bsde_philox_increments", "bsde_toy_em_fwd", "bsde_toy_em_bwd", "bsde_chain_workspace_bytes", "bsde_chain_acts_floats",
            "bsde_chain_out_offset", "bsde_chain_fwd", "bsde_chain_bwd", "bsde_nchw_reinterp_axpy", "bsde_brownian_tree",
            "bsde_step_update_prod", "bsde_brownian_tree_threefry", "bsde_threefry_split", "bsde_sdelayer_fwd", "bsde_sdelayer_bwd"]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise BsdeError(f"{_LIB_PATH} is missing: run `python -m bayesde_b200.build` (there is no CPU/eager fallback)")
    L = C.CDLL(_LIB_PATH)
    for name in _EXPORTS:
        if not hasattr(L, name):
            raise BsdeError(f"libbsde.so does not export {name}")
    L.bsde_version.restype = C.c_int
    L.bsde_last_error.restype = C.c_char_p
    L.bsde_launch_count.restype = C.c_uint64
    L.bsde_wsde_workspace_bytes.restype = C.c_size_t
    L.bsde_wsde_workspace_bytes.argtypes = [C.POINTER(WsdeDesc), C.POINTER(FwParams)]
    vp = C.c_void_p
    L.bsde_wsde_fwd.argtypes = [C.POINTER(WsdeDesc), C.POINTER(FwParams), vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                C.c_size_t, vp]
    L.bsde_wsde_bwd.argtypes = [C.POINTER(WsdeDesc), C.POINTER(FwParams), vp, vp, vp, vp, vp, vp, vp, vp,
                                C.POINTER(FwParams), vp, C.c_size_t, vp]
    L.bsde_tconv_fwd.argtypes = [C.POINTER(ConvLayer), C.c_int32, C.c_int32, vp, vp, C.c_int64, C.c_float,
                                 C.c_int32, C.c_int32, C.c_float, vp, vp, C.c_int32, vp, C.c_size_t, vp]
    L.bsde_tconv_fwd_workspace_bytes.restype = C.c_size_t
    L.bsde_tconv_fwd_workspace_bytes.argtypes = [C.POINTER(ConvLayer), C.c_int32, C.c_int32]
    L.bsde_tconv_wgrad_workspace_bytes.restype = C.c_size_t
    L.bsde_tconv_wgrad_workspace_bytes.argtypes = [C.POINTER(ConvLayer), C.c_int32, C.c_int32]
    L.bsde_tconv_wgrad.argtypes = [C.POINTER(ConvLayer), C.c_int32, C.c_int32, vp, vp, C.c_float, C.c_float, vp,
                                   C.c_int64, C.c_int32, vp, C.c_size_t, vp]
    L.bsde_xpath_workspace_bytes.restype = C.c_size_t
    L.bsde_xpath_workspace_bytes.argtypes = [C.POINTER(XpathDesc), C.c_int32]
    L.bsde_xpath_fwd.argtypes = [C.POINTER(XpathDesc), _fp, _fp, vp, vp, vp, vp, C.c_size_t, vp]
    L.bsde_xpath_bwd.argtypes = [C.POINTER(XpathDesc), _fp, _fp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]
    L.bsde_xpath_acts_bytes.restype = C.c_size_t
    L.bsde_xpath_acts_bytes.argtypes = [C.POINTER(XpathDesc)]
    L.bsde_step_update.argtypes = [C.c_int32, C.c_int64, C.c_float, vp, vp, vp, vp, vp, vp, vp, vp]
    L.bsde_philox_increments.argtypes = [C.c_uint64, C.c_uint32, C.c_int32, C.c_int32, C.c_int64, vp, vp, vp]
    L.bsde_brownian_tree.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, vp, vp]
    L.bsde_brownian_tree_threefry.argtypes = [C.c_uint32, C.c_uint32, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float,
                                              C.c_int64, C.c_int64, vp, vp]
    L.bsde_threefry_split.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(C.c_uint32)]
    L.bsde_step_update_prod.argtypes = [C.c_int64, C.c_float, vp, vp, vp, vp, vp, vp]
    L.bsde_sdelayer_fwd.argtypes = [C.POINTER(SDELayerDesc), C.POINTER(FwParams), vp, vp, vp, vp, vp, vp]
    L.bsde_sdelayer_bwd.argtypes = [C.POINTER(SDELayerDesc), C.POINTER(FwParams), vp, vp, vp, vp, vp, vp, C.POINTER(FwParams), vp]
    L.bsde_toy_em_fwd.argtypes = [C.POINTER(ToyDesc), vp, vp, vp, C.c_uint64] + [vp] * 8 + [vp, vp, vp]
    L.bsde_toy_em_bwd.argtypes = [C.POINTER(ToyDesc), vp, vp, vp, vp] + [vp] * 8 + [vp] * 8 + [vp]
    if L.bsde_version() != 200:
        raise BsdeError(f"libbsde.so version {L.bsde_version()} != 200: rebuild")
    _lib = L
    return L


def check(status, what):
    if status == 0:
        return
    msg = lib().bsde_last_error().decode()
    if status == -1:
        raise ValueError(f"{what}: {msg}")
    if status == -2:
        raise NotImplementedError(f"{what}: {msg}")
    raise BsdeError(f"{what}: status {status}: {msg}")


def ptr(t):
    """Device pointer of a contiguous fp32 CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise BsdeError("bayesde_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if t.device.index != torch.cuda.current_device():
        # libbsde.so launches on the CUDA current device and on torch's current stream of that device; a tensor living on
        # another GPU would be dereferenced by the wrong context
        raise BsdeError(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: wrap the call in "
                        f"`with torch.cuda.device({t.device.index}):` (one process per GPU is the supported layout)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError(f"expected a contiguous float32 tensor, got {t.dtype} contiguous={t.is_contiguous()}")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    """torch's current stream on the current device (ptr() guarantees that the tensors of the call live there)."""
    return C.c_void_p(torch.cuda.current_stream(torch.cuda.current_device()).cuda_stream)


_ws_cache = {}


def workspace(nbytes, device, tag="ws"):
    """Grow-only scratch buffer per (device, stream, tag); owned by PyTorch's allocator."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _ws_cache.pop(key, None)
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def host_floats(vals):
    arr = (C.c_float * len(vals))(*[float(v) for v in vals])
    return arr
