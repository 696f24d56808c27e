"""ctypes binding of libddmi_b200.so (the C ABI declared in include/ddmi_b200.h).

There is deliberately no fallback: if the shared library is missing, or a call
returns a non-zero status, a RuntimeError is raised (the reference's native ops
behave the same way: TORCH_CHECK -> RuntimeError, op/fused_bias_act.cpp:7-13).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# DDMI_B200_LIB: dev tools point this at libddmi_b200_prof.so (the build with in-kernel counters, `make prof`)
LIB_PATH = os.environ.get("DDMI_B200_LIB") or os.path.join(_HERE, "libddmi_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

ABI_VERSION = 10
PREC_FP32 = 0
PREC_BF16X3 = 1
PREC_F16F8 = 2
STORE_F32, STORE_F32_CLAMP, STORE_U8_CHANNELS_LAST = 0, 1, 2
STORE_MODES = {None: 0, 'f32': 0, 'clamp': 1, 'u8': 2}
NOISE_NONE, NOISE_TENSORS, NOISE_PHILOX = 0, 1, 2

EXPORTS = (
    "ddmi_abi_version", "ddmi_last_error", "ddmi_status_string", "ddmi_device_info",
    "ddmi_decode_image", "ddmi_decode_image_store", "ddmi_decode_image_noise", "ddmi_planes_to_channels_last", "ddmi_decode_occupancy",
    "ddmi_decode_video", "ddmi_decode_video_store", "ddmi_decode_video_ws", "ddmi_video_workspace_bytes",
    "ddmi_decode_occupancy_lattice", "ddmi_occupancy_lattice_workspace_bytes",
    "ddmi_nerf_mlp", "ddmi_nerf_render", "ddmi_nerf_render_z", "ddmi_sample_pdf", "ddmi_plane_head", "ddmi_plane_tail", "ddmi_mcubes_workspace_bytes", "ddmi_mcubes_count", "ddmi_mcubes_emit", "ddmi_selftest_tma", "ddmi_selftest_umma", "ddmi_selftest_umma2", "ddmi_selftest_f16f8", "ddmi_debug_profile",
    "ddmi_debug_trace", "ddmi_debug_set", "ddmi_debug_gatherbench", "ddmi_debug_ringbench", "ddmi_debug_microbench",
)


class Plane(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("height", ctypes.c_int32), ("width", ctypes.c_int32)]


class Weights(ctypes.Structure):
    _fields_ = [("precision", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("gemm", ctypes.c_void_p), ("gemm_bytes", ctypes.c_uint64),
                ("vec", ctypes.c_void_p), ("vec_floats", ctypes.c_uint64),
                ("program", ctypes.c_void_p), ("program_host", ctypes.c_void_p), ("program_words", ctypes.c_uint64),
                ("vec_host", ctypes.c_void_p)]


def build_library(verbose=False):
    """Compile the CUDA sources for sm_100a into LIB_PATH (nvcc cross-compiles
    without a GPU).  Used by __graft_entry__.build()."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libddmi_b200.so failed (see output above)")
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC_DIR}` "
                "(or __graft_entry__.build()). ddmi_b200 has no CPU / eager fallback.")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
        L.ddmi_abi_version.restype = ctypes.c_int
        L.ddmi_last_error.restype = ctypes.c_char_p
        L.ddmi_status_string.restype = ctypes.c_char_p
        L.ddmi_status_string.argtypes = [ctypes.c_int]
        L.ddmi_device_info.argtypes = [ctypes.POINTER(i32)] * 3
        L.ddmi_decode_image.argtypes = [ctypes.POINTER(Plane), i32, i32, vp, vp, i64,
                                        ctypes.POINTER(Weights), vp, vp]
        L.ddmi_decode_image_store.argtypes = [ctypes.POINTER(Plane), i32, i32, vp, vp, i64,
                                              ctypes.POINTER(Weights), i32, vp, vp]
        L.ddmi_decode_image_noise.argtypes = [ctypes.POINTER(Plane), i32, i32, vp, vp, i64, ctypes.POINTER(Weights), i32,
                                              i32, ctypes.POINTER(vp), ctypes.c_uint64, vp, vp]
        L.ddmi_planes_to_channels_last.argtypes = [vp, vp, i32, i32, i32, i32, vp]
        L.ddmi_decode_occupancy.argtypes = [ctypes.POINTER(Plane), i32, i32, i32, vp, i64, i64, f32,
                                            ctypes.POINTER(Weights), vp, vp]
        L.ddmi_decode_video.argtypes = [ctypes.POINTER(Plane), i32, i32, vp, vp, vp, i32, i32, i32,
                                        ctypes.POINTER(Weights), vp, vp]
        L.ddmi_decode_video_store.argtypes = [ctypes.POINTER(Plane), i32, i32, vp, vp, vp, i32, i32, i32,
                                              ctypes.POINTER(Weights), i32, vp, vp]
        L.ddmi_decode_video_ws.argtypes = [ctypes.POINTER(Plane), i32, i32, vp, vp, vp, i32, i32, i32,
                                           ctypes.POINTER(Weights), i32, vp, vp, ctypes.c_uint64, vp]
        L.ddmi_video_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
        L.ddmi_occupancy_lattice_workspace_bytes.argtypes = [i32, i32, i32, i32]
        L.ddmi_occupancy_lattice_workspace_bytes.restype = ctypes.c_int64
        L.ddmi_decode_occupancy_lattice.argtypes = [ctypes.POINTER(Plane), i32, i32, i32, vp, i32, i32, i32, f32,
                                                    ctypes.POINTER(Weights), vp, vp, ctypes.c_uint64, vp]
        L.ddmi_video_workspace_bytes.restype = ctypes.c_int64
        L.ddmi_nerf_mlp.argtypes = [vp, i64, i32, i32, f32, ctypes.POINTER(Weights), vp, vp]
        L.ddmi_nerf_render.argtypes = [ctypes.POINTER(Plane), i32, i32, i32, vp, i64, i32, vp, i32, f32, f32,
                                       i32, ctypes.POINTER(Weights), vp, vp, vp]
        L.ddmi_nerf_render_z.argtypes = L.ddmi_nerf_render.argtypes
        L.ddmi_sample_pdf.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp]
        L.ddmi_plane_head.argtypes = [vp, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp]
        L.ddmi_plane_tail.argtypes = [vp, i32, i32, i32, i32, vp, vp, i32, ctypes.c_float, vp, vp, i32, i32, i32, vp, vp, vp]
        L.ddmi_mcubes_workspace_bytes.argtypes = [i32, i32, i32, i32, vp]
        L.ddmi_mcubes_count.argtypes = [vp, i32, i32, i32, i32, ctypes.c_double, ctypes.c_double, vp, ctypes.c_uint64, vp, vp]
        L.ddmi_mcubes_emit.argtypes = [vp, i32, i32, i32, i32, ctypes.c_double, ctypes.c_double, vp, vp, vp, vp, vp]
        L.ddmi_selftest_tma.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
        L.ddmi_selftest_umma.argtypes = [vp, vp, vp, i32, i32, vp]
        L.ddmi_selftest_umma2.argtypes = [vp, vp, vp, i32, i32, vp]
        L.ddmi_selftest_f16f8.argtypes = [vp, vp, vp, i32, i32, vp]
        L.ddmi_debug_profile.argtypes = [ctypes.POINTER(ctypes.c_uint64), i32]
        L.ddmi_debug_trace.argtypes = [ctypes.POINTER(ctypes.c_uint64), i32, ctypes.POINTER(i32), i32]
        L.ddmi_debug_set.argtypes = [i32]
        L.ddmi_debug_gatherbench.argtypes = [i32, i32, vp, ctypes.c_uint32, i32, i32, i32, vp, vp, vp]
        L.ddmi_debug_ringbench.argtypes = [vp, ctypes.c_uint64, i32, i32, i32, i32, vp, vp]
        L.ddmi_debug_microbench.argtypes = [i32, i32, vp, vp, vp, vp]
        for name in EXPORTS:
            getattr(L, name)  # AttributeError here = header / library out of sync
        if L.ddmi_abi_version() != ABI_VERSION:
            raise RuntimeError("libddmi_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(status):
    if status != 0:
        L = lib()
        kind = L.ddmi_status_string(status).decode()
        msg = L.ddmi_last_error().decode()
        raise RuntimeError(f"ddmi_b200: {kind}: {msg}")


def planes_array(tensors):
    """ctypes array of ddmi_plane_t from contiguous fp32 CUDA tensors (B,C,H,W)."""
    arr = (Plane * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i].data = t.data_ptr()
        arr[i].height = t.shape[-2]
        arr[i].width = t.shape[-1]
    return arr


def is_channels_last(t):
    """(B,C,H,W) fp32 tensor whose memory is (B,H,W,C) dense -- torch.channels_last, e.g. the planes ddmi_b200.plane_tail emits."""
    import torch
    return (t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] > 1 and t.stride(1) == 1
            and t.is_contiguous(memory_format=torch.channels_last))


def planes_channels_last(tensors, stream):
    """Channels-last copies (B,H,W,C) of contiguous fp32 CUDA planes (B,C,H,W), made by the library's own
    transpose kernel.  Returns (list of NHWC tensors, ctypes plane array)."""
    import torch
    outs = []
    for t in tensors:
        b, c, h, w = t.shape
        if is_channels_last(t):          # emitted channels-last by the producer (plane_tail): consumed as is, no copy
            outs.append(t.permute(0, 2, 3, 1))
            continue
        t = t.contiguous()
        o = torch.empty((b, h, w, c), device=t.device, dtype=torch.float32)
        check(lib().ddmi_planes_to_channels_last(t.data_ptr(), o.data_ptr(), b, c, h, w, stream))
        outs.append(o)
    arr = (Plane * len(outs))()
    for i, o in enumerate(outs):
        arr[i].data = o.data_ptr()
        arr[i].height = o.shape[1]
        arr[i].width = o.shape[2]
    return outs, arr


def weights_struct(packed):
    """packed: packing.Packed (precision, gemm tensor, vec tensor)."""
    w = Weights()
    w.precision = packed.precision
    w.reserved = ((1 if getattr(packed, "pair", False) else 0) | (2 if os.environ.get("DDMI_B200_NO_TMA_PATCH") == "1" else 0)
                  | (4 if getattr(packed, "ts", False) else 0))
    w.gemm = packed.gemm.data_ptr()
    w.gemm_bytes = packed.gemm.numel() * packed.gemm.element_size()
    w.vec = packed.vec.data_ptr()
    w.vec_floats = packed.vec.numel()
    if packed.program is not None:
        w.program = packed.program.data_ptr()
        w.program_host = packed.program_host.data_ptr()
        w.program_words = packed.program_host.numel()
    if getattr(packed, "vec_host", None) is not None:
        w.vec_host = packed.vec_host.data_ptr()
    return w
