"""ctypes binding of libidee_b200.so (the C ABI declared in include/idee_b200.h).

There is no fallback: if the library is missing and cannot be built, or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

from . import build as _build

c_i64, c_int, c_f32, c_vp, c_sz = C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_size_t


class SwinDesc(C.Structure):
    _fields_ = [(n, c_int) for n in ("N", "V", "T", "H", "W", "C", "heads", "hidden", "wd", "wh", "ww", "st", "sh", "sw", "rpb_rows")] + \
               [("scale", c_f32), ("param_stride", c_i64), ("precision", c_int)] + \
               [(n, c_vp) for n in ("embed_x", "embed_w", "embed_b", "embed_gw", "embed_gb")] + [("act_dtype", c_int), ("x_dtype", c_int), ("out_dtype", c_int)]


class ConvDesc(C.Structure):
    _fields_ = [(n, c_int) for n in ("N", "V", "Vw", "Cin", "Cout", "Ti", "Hi", "Wi", "To", "Ho", "Wo", "proj", "relu", "precision")] + \
               [(n, c_i64) for n in ("x_sn", "x_sv", "x_st", "x_sh", "x_sw", "x_sg")] + [("in_cpg", c_int)] + \
               [(n, c_i64) for n in ("y_sn", "y_sv", "y_st", "y_sh", "y_sw", "y_sg")] + [("out_cpg", c_int)] + \
               [(n, c_int) for n in ("x_dtype", "y_dtype", "gx_dtype", "umma16", "umma96", "cin_real")]


_SIGS = {
    "idee_last_error": (C.c_char_p, []),
    "idee_version": (c_int, []),
    "idee_check_device": (c_int, []),
    "idee_embed_ln_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp] + [c_int] * 7 + [c_vp]),
    "idee_embed_ln_bwd_workspace_bytes": (c_sz, [c_int]),
    "idee_embed_ln_bwd": (c_int, [c_vp] * 7 + [c_int] * 7 + [c_vp, c_sz, c_vp]),
    "idee_swin_block_packed_floats": (c_int, [c_int]),
    "idee_swin_block_fwd": (c_int, [C.POINTER(SwinDesc)] + [c_vp] * 7),
    "idee_swin_block_bwd_workspace_bytes": (c_sz, [C.POINTER(SwinDesc)]),
    "idee_swin_block_bwd": (c_int, [C.POINTER(SwinDesc)] + [c_vp] * 8 + [c_sz, c_vp]),
    "idee_conv3d_fwd_workspace_bytes": (c_sz, [C.POINTER(ConvDesc)]),
    "idee_conv3d_fwd": (c_int, [C.POINTER(ConvDesc)] + [c_vp] * 5 + [c_sz, c_vp]),
    "idee_conv3d_dgrad_workspace_bytes": (c_sz, [C.POINTER(ConvDesc)]),
    "idee_conv3d_dgrad": (c_int, [C.POINTER(ConvDesc)] + [c_vp] * 5 + [c_sz, c_vp]),
    "idee_conv3d_wgrad_workspace_bytes": (c_sz, [C.POINTER(ConvDesc)]),
    "idee_conv3d_wgrad": (c_int, [C.POINTER(ConvDesc)] + [c_vp] * 5 + [c_sz, c_vp]),
    "idee_conv3d_bwd_workspace_bytes": (c_sz, [C.POINTER(ConvDesc)]),
    "idee_conv3d_bwd": (c_int, [C.POINTER(ConvDesc)] + [c_vp] * 8 + [c_sz, c_vp]),
    "idee_lfq_workspace_bytes": (c_sz, [c_i64]),
    "idee_lfq_fwd": (c_int, [c_vp] * 9 + [c_i64, c_int, c_int, c_int] + [c_f32] * 4 + [c_vp, c_sz, c_vp, c_vp]),
    "idee_lfq_bwd": (c_int, [c_vp] * 10 + [c_i64] + [c_f32] * 4 + [c_vp, c_sz, c_vp]),
    "idee_lfqk_workspace_bytes": (c_sz, [c_int]),
    "idee_lfqk_fwd": (c_int, [c_vp] * 8 + [c_i64, c_int, c_int, c_int] + [c_f32] * 4 + [c_vp, c_sz, c_vp]),
    "idee_lfqk_bwd": (c_int, [c_vp] * 9 + [c_i64, c_int, c_int] + [c_f32] * 4 + [c_vp, c_sz, c_vp]),
    "idee_bce_loss_workspace_bytes": (c_sz, [c_int]),
    "idee_bce_loss_fwd": (c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_i64] + [c_vp] * 5 + [c_sz, c_vp]),
    "idee_anomaly_l1_workspace_bytes": (c_sz, [c_i64]),
    "idee_anomaly_l1_fwd": (c_int, [c_vp] * 3 + [c_int] * 3 + [c_i64, c_int, c_vp, c_vp, c_sz, c_vp]),
    "idee_anomaly_l1_bwd": (c_int, [c_vp] * 3 + [c_int] * 3 + [c_i64, c_int] + [c_vp] * 4),
    "idee_anomaly_rank1_workspace_bytes": (c_sz, [c_i64]),
    "idee_anomaly_rank1_fwd": (c_int, [c_vp] * 5 + [c_int] * 3 + [c_i64, c_int, c_vp, c_vp, c_sz, c_vp]),
    "idee_anomaly_rank1_bwd": (c_int, [c_vp] * 5 + [c_int] * 3 + [c_i64, c_int] + [c_vp] * 6),
    "idee_rank1_planes_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_vp]),
    "idee_rank1_planes_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_vp]),
    "idee_ln_act_res_fwd": (c_int, [c_vp] * 6 + [c_int, c_int, c_i64, c_int, c_vp]),
    "idee_ln_act_res_bwd_workspace_bytes": (c_sz, [c_int]),
    "idee_ln_act_res_bwd": (c_int, [c_vp] * 7 + [c_int, c_int, c_i64, c_int, c_vp, c_sz, c_vp]),
    "idee_adam_step": (c_int, [c_vp] * 4 + [c_i64] + [c_f32] * 5 + [c_int, c_vp]),
    "idee_adam_step_state": (c_int, [c_vp] * 4 + [c_i64, c_vp] + [c_f32] * 4 + [c_vp]),
}
EXPORTS = tuple(_SIGS)

_lock = threading.Lock()
_lib = None


def library_path() -> str:
    return _build.LIB


def header_version() -> int:
    """IDEE_B200_VERSION of include/idee_b200.h: bumped on every ABI change (argument lists, descriptor structs)."""
    import re
    with open(os.path.join(_build.INCLUDE, "idee_b200.h")) as fh:
        m = re.search(r"#define\s+IDEE_B200_VERSION\s+(\d+)", fh.read())
    if m is None:
        raise RuntimeError("include/idee_b200.h does not define IDEE_B200_VERSION")
    return int(m.group(1))


def load(build_if_missing: bool = True):
    """Load (building in-tree with nvcc if needed) libidee_b200.so and declare every prototype.

    A stale library is never bound to the current prototypes: if the sources changed and the rebuild fails, the error is
    raised (the old binary is only used when its build stamp matches the sources), and the library's idee_version() must
    equal the header's IDEE_B200_VERSION."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB
        if build_if_missing:
            try:
                path = _build.build()          # no-op when the in-tree library matches the sources (fingerprint stamp)
            except Exception as e:
                if not os.path.exists(path):
                    raise RuntimeError(f"{path} is missing and could not be built ({e}); there is no fallback path") from e
                if not _build.up_to_date():
                    raise RuntimeError(f"{path} is older than its sources and could not be rebuilt ({e}); refusing to bind the "
                                       f"current prototypes to a stale binary") from e
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -m idee_b200.build` (needs nvcc); there is no fallback path")
        lib = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)   # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        want = header_version()
        if lib.idee_version() != want:
            raise RuntimeError(f"{path} reports ABI version {lib.idee_version()}, include/idee_b200.h declares {want}: rebuild with "
                               f"`python -m idee_b200.build --force`")
        _lib = lib
        return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().idee_last_error().decode(errors="replace")
        raise RuntimeError(f"idee_b200 {what} failed (rc={rc}): {msg}")


# kernels launched by each C-ABI entry point (used for the gpu_launches figure of bench.py)
# numeric policy of the GEMM-shaped kernels: "fp32" = exact CUDA-core path, "bf16" = bf16 tensor-core operands with fp32
# accumulation (activations stay fp32 in HBM, the LFQ quantiser and all losses are always fp32)
PRECISION = os.environ.get("IDEE_B200_PRECISION", "fp32")
if PRECISION not in ("fp32", "bf16"):
    raise ValueError("IDEE_B200_PRECISION must be fp32 or bf16")


# tcgen05 + TMEM kernel for the bf16-storage 16 -> 16 proj conv (forward + data gradient), bf16 mode only; on by default
# (measured 9 % faster than the mma.sync kernel on the forward conv), IDEE_B200_UMMA16=0 selects the mma.sync kernel
UMMA16 = os.environ.get("IDEE_B200_UMMA16", "1") == "1"


# warp-specialised tcgen05 + TMEM kernel for the dense 96 -> 96 classifier conv (forward + data gradient), bf16 mode only;
# on by default (measured 1.9x / 1.5x faster than the mma.sync kernel), IDEE_B200_UMMA96=0 selects the mma.sync kernel
UMMA96 = os.environ.get("IDEE_B200_UMMA96", "1") == "1"


# tcgen05 / TMEM Swin block kernels on bf16 token storage (swin_umma.cuh), bf16 mode only; IDEE_B200_SWIN_UMMA=0 selects the
# mma.sync kernels on fp32 token storage (swin_tc.cuh)
SWIN_UMMA = os.environ.get("IDEE_B200_SWIN_UMMA", "1") == "1"


def set_swin_umma(on: bool) -> None:
    global SWIN_UMMA
    SWIN_UMMA = bool(on)


def swin_umma() -> bool:
    return SWIN_UMMA and PRECISION == "bf16"


def set_umma96(on: bool) -> None:
    global UMMA96
    UMMA96 = bool(on)


def set_umma16(on: bool) -> None:
    global UMMA16
    UMMA16 = bool(on)


def set_precision(mode: str) -> None:
    global PRECISION
    if mode not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    PRECISION = mode


LAUNCHES = {"embed_ln_fwd": 1, "embed_ln_bwd": 2, "swin_block_fwd": 1, "swin_block_bwd": 4, "conv3d_fwd": 1, "conv3d_dgrad": 1,
            "conv3d_wgrad": 2, "conv3d_bwd": 3, "conv3d_bwd_bf16": 5, "conv3d_fwd_bf16": 2, "conv3d_dgrad_bf16": 3, "conv3d_wgrad_bf16": 2, "lfq_fwd": 2, "lfq_fwd_eval": 1, "lfq_bwd": 2, "bce_loss_fwd": 4, "anomaly_l1_fwd": 2,
            "anomaly_l1_bwd": 1, "anomaly_rank1_fwd": 2, "anomaly_rank1_bwd": 1, "rank1_planes_fwd": 1,
            "rank1_planes_bwd": 1, "adam_step": 1, "adam_step_state": 2, "ln_act_res_fwd": 1, "ln_act_res_bwd": 2, "lfqk_fwd": 2, "lfqk_fwd_eval": 1, "lfqk_bwd": 2}


class Profile:
    """Opt-in per-entry-point accounting: launch counts always, CUDA-event timing when ``events`` is on."""
    launches = 0
    events = False
    records = []      # (name, tag, start_event, end_event)

    @classmethod
    def reset(cls, events: bool = False):
        cls.launches, cls.events, cls.records = 0, events, []

    @classmethod
    def summary(cls):
        """{name: (calls, total_ms)} -- call after torch.cuda.synchronize()."""
        out = {}
        for name, tag, e0, e1 in cls.records:
            key = name if tag is None else f"{name}[{tag}]"
            c, t = out.get(key, (0, 0.0))
            out[key] = (c + 1, t + e0.elapsed_time(e1))
        return out


def run(what: str, fn, *args, tag=None):
    """Call one C-ABI entry point on the current stream of the CURRENT device; raise on a non-zero return code.
    Callers pass ``stream()`` and tensors validated by ``require_cuda`` (all on the current device)."""
    Profile.launches += LAUNCHES[what]
    if Profile.events:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        Profile.records.append((what, tag, e0, e1))
    else:
        rc = fn(*args)
    check(rc, what)


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def require_cuda(*tensors):
    """Every tensor of a call must live on the CURRENT CUDA device: the kernels launch on the current device's stream with raw
    pointers, so a tensor on another GPU (model.to('cuda:1') without torch.cuda.set_device(1), nn.DataParallel replicas that
    share the packed parameters of device 0) would be dereferenced on the wrong device."""
    cur = torch.cuda.current_device() if torch.cuda.is_available() else -1
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("idee_b200 kernels run on CUDA tensors only (there is no CPU path); got a %s tensor" % t.device)
        if t.device.index != cur:
            raise RuntimeError(f"idee_b200: tensor on {t.device} but the current device is cuda:{cur}; call torch.cuda.set_device() "
                               f"(one process per GPU, see idee_b200.trainer). nn.DataParallel is not supported: use Trainer/torchrun")


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)
