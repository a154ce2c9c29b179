"""clumpy_b200 -- ctypes binding of libclumpy_cuda.so (include/clumpy_cuda.h).

The product is the C-ABI library; this module is the thin Python door that tests and bench.py use.
Function names and argument meaning follow clumpy's commands (generate_simplex, curl_2d,
splat_points, advect_points).  There is no CPU fallback: a missing library or a machine without a
CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclumpy_cuda.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE = 0, 1, 2, 3, 4
MAX_KERNEL_SIZE = 31

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i16p = C.POINTER(C.c_int16)
FRAME_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, _u8p)

# name -> (restype, argtypes): every symbol include/clumpy_cuda.h declares
SIGNATURES = {
    "clumpy_cuda_last_error": (C.c_char_p, []),
    "clumpy_cuda_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "clumpy_cuda_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_destroy": (None, [C.c_void_p]),
    "clumpy_cuda_sync": (C.c_int, [C.c_void_p]),
    "clumpy_cuda_host_alloc": (C.c_void_p, [C.c_size_t]),
    "clumpy_cuda_host_free": (None, [C.c_void_p]),
    "clumpy_cuda_dev_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_dev_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clumpy_cuda_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "clumpy_cuda_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "clumpy_cuda_timer_start": (C.c_int, [C.c_void_p]),
    "clumpy_cuda_timer_stop": (C.c_int, [C.c_void_p, _f32p]),
    "clumpy_cuda_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "clumpy_cuda_atomic_peak": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int, _f32p]),
    "clumpy_cuda_simplex_perm": (C.c_int, [C.c_int64, _i16p]),
    "clumpy_cuda_generate_simplex": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_double,
                                               C.c_double, C.c_int64, C.c_void_p, _f32p, _f32p]),
    "clumpy_cuda_generate_simplex_dev": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                   C.c_uint32, C.c_double, C.c_double, C.c_int64,
                                                   C.c_void_p, _f32p, _f32p]),
    "clumpy_cuda_generate_simplex_octaves_dev": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                           C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                                           C.c_void_p, _f32p, _f32p]),
    "clumpy_cuda_curl_2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                      _f32p, _f32p]),
    "clumpy_cuda_curl_2d_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                          C.c_void_p, C.c_void_p, _f32p, _f32p]),
    "clumpy_cuda_splat_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                           C.c_uint32, C.c_void_p]),
    "clumpy_cuda_splat_disks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                          C.c_uint32, C.c_void_p, C.c_float, C.c_int, C.c_int]),
    "clumpy_cuda_splat_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                          C.c_float, C.c_int, C.c_int]),
    "clumpy_cuda_age_offsets": (C.c_int, [C.c_uint32, C.c_uint32, _u32p]),
    "clumpy_cuda_age_offsets_dev": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "clumpy_cuda_advect_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                           C.c_uint32, C.c_uint32, C.c_float, C.c_int, C.c_float,
                                           C.c_int, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_advect_begin_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_row_histogram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, _u32p]),
    "clumpy_cuda_advect_step": (C.c_int, [C.c_void_p, C.c_uint32]),
    "clumpy_cuda_advect_record": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "clumpy_cuda_advect_profile_stages": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, _f32p, _f32p, _u32p]),
    "clumpy_cuda_advect_count": (C.c_int, [C.c_void_p]),
    "clumpy_cuda_advect_cells": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                           _u32p, _u32p, _u32p, _u32p]),
    "clumpy_cuda_advect_composite_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]),
    "clumpy_cuda_advect_finish_frame": (C.c_int, [C.c_void_p]),
    "clumpy_cuda_advect_image_dev": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "clumpy_cuda_ipc_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clumpy_cuda_advect_route_bytes": (C.c_int, [C.c_uint32, C.c_int, C.POINTER(C.c_size_t)]),
    "clumpy_cuda_advect_route_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint32,
                                                 C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "clumpy_cuda_advect_route_peers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_advect_owned_rows": (C.c_int, [C.c_void_p, _u32p, _u32p]),
    "clumpy_cuda_advect_route_steps": (C.c_int, [C.c_void_p, C.c_uint32]),
    "clumpy_cuda_advect_route_record": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "clumpy_cuda_advect_route_errors": (C.c_int, [C.c_void_p, _u32p]),
    "clumpy_cuda_pendulum_phase": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "clumpy_cuda_pendulum_phase_dev": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float,
                                                 C.c_float, C.c_float, C.c_void_p]),
    "clumpy_cuda_cull_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                          C.c_void_p, _u32p]),
    "clumpy_cuda_advect_route_profile": (C.c_int, [C.c_void_p, C.c_uint32, _f32p, _f32p, _f32p, _u32p]),
    "clumpy_cuda_advect_simframe": (C.c_int, [C.c_void_p, _u32p]),
    "clumpy_cuda_advect_npts": (C.c_int, [C.c_void_p, _u32p]),
    "clumpy_cuda_advect_read_points_indexed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "clumpy_cuda_advect_frame_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clumpy_cuda_advect_wait_frames": (C.c_int, [C.c_void_p]),
    "clumpy_cuda_advect_read_points": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clumpy_cuda_advect_read_image": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clumpy_cuda_advect_end": (C.c_int, [C.c_void_p]),
    "clumpy_cuda_advect_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                            C.c_uint32, C.c_uint32, C.c_float, C.c_int, C.c_float,
                                            C.c_int, C.c_uint32, C.c_void_p, FRAME_FN, C.c_void_p]),
    "clumpy_cuda_advect_points_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_uint32, C.c_void_p,
                                                  C.c_uint32, C.c_uint32, C.c_float, C.c_int, C.c_float,
                                                  C.c_int, C.c_uint32, C.c_void_p, FRAME_FN, C.c_void_p]),
}

CELLS_AUTO, CELLS_COUNT, CELLS_F16, CELLS_U32, CELLS_EXACT = 0, 1, 2, 3, 4
_CELLS = {"auto": CELLS_AUTO, "count": CELLS_COUNT, "f16": CELLS_F16, "u32": CELLS_U32, "exact": CELLS_EXACT}


class AdvectDesc(C.Structure):
    """struct clumpy_advect_desc (include/clumpy_cuda.h)."""
    _fields_ = [("size", C.c_uint32), ("pts", C.c_void_p), ("npts", C.c_uint32), ("pts_on_device", C.c_int),
                ("vel", C.c_void_p), ("vel_on_device", C.c_int), ("width", C.c_uint32), ("height", C.c_uint32),
                ("step_size", C.c_float), ("kernel_size", C.c_int), ("decay", C.c_float), ("wrapx", C.c_int),
                ("nframes", C.c_uint32), ("age_offset", C.c_void_p), ("age_on_device", C.c_int),
                ("shard_row_lo", C.c_uint32), ("shard_row_hi", C.c_uint32), ("cells", C.c_int), ("overlap", C.c_int),
                ("fuse", C.c_int), ("regroup", C.c_int), ("cell_rows_multiple", C.c_int)]


_LIB = None


class ClumpyCudaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libclumpy_cuda error {code}: {message}")
        self.code = code


def lib():
    """Loads libclumpy_cuda.so; raises if it has not been built (no fallback of any kind)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `make lib` (or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _check(rc):
    if rc != OK:
        raise ClumpyCudaError(rc, lib().clumpy_cuda_last_error().decode(errors="replace"))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _vp(a):
    """void* of a numpy array or a raw integer address."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)


def splat_multi(devices, pts, width, height, alpha, kernel_size, wrapx=False, img=None):
    """`clumpy splat_points` on several GPUs from this process (clumpy_cuda_splat_multi): fast path when alpha == 1
    and kernel_size == 1, splat_disks otherwise."""
    pts = _f32(pts)
    img = np.zeros((height, width), np.uint8) if img is None else np.ascontiguousarray(img).copy()
    devs = (C.c_int * len(devices))(*devices)
    _check(lib().clumpy_cuda_splat_multi(devs, len(devices), _vp(pts), len(pts), width, height, _vp(img), alpha, kernel_size, int(wrapx)))
    return img


def advect_points_multi(devices, pts, vel, step_size, kernel_size, decay, nframes, on_frame=None):
    """The whole `clumpy advect_points` command on several GPUs from this process (clumpy_cuda_advect_points_multi);
    returns (nframes, H, W) u8 unless on_frame(animframe, pixels) consumes the frames."""
    pts, vel = _f32(pts), _f32(vel)
    h, w = vel.shape[:2]
    wrapx, decay = (1, -decay) if decay < 0 else (0, decay)
    frames = None if on_frame else np.empty((nframes, h, w), np.uint8)

    def cb(_user, animframe, pixels):
        px = np.ctypeslib.as_array(pixels, shape=(h, w))
        if on_frame:
            return int(on_frame(animframe, px) or 0)
        frames[animframe] = px
        return 0

    devs = (C.c_int * len(devices))(*devices)
    _check(lib().clumpy_cuda_advect_points_multi(devs, len(devices), _vp(pts), len(pts), _vp(vel), w, h, step_size,
                                                 kernel_size, decay, wrapx, nframes, None, FRAME_FN(cb), None))
    return frames


def route_inbox_bytes(capacity, world):
    n = C.c_size_t()
    _check(lib().clumpy_cuda_advect_route_bytes(capacity, world, C.byref(n)))
    return n.value


def age_offsets(nframes, npts):
    """Reset offsets of advect_points.cc:127-135 (libstdc++ mt19937 + uniform_int_distribution)."""
    out = np.empty(npts, np.uint32)
    _check(lib().clumpy_cuda_age_offsets(nframes, npts, out.ctypes.data_as(_u32p)))
    return out


def simplex_perm(seed):
    out = np.empty(256, np.int16)
    _check(lib().clumpy_cuda_simplex_perm(int(seed), out.ctypes.data_as(_i16p)))
    return out


class Context:
    """One GPU.  `stream` is an integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        _check(lib().clumpy_cuda_create(device, C.c_void_p(stream) if stream else None,
                                        C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            arena = getattr(self, "_route_arena", None)
            if arena:  # the inbox arena kept between routed multi-GPU runs (clumpy_b200/multigpu.py)
                for p in arena["peers"]:
                    if p != arena["inbox"]:
                        lib().clumpy_cuda_ipc_close(self._h, C.c_void_p(p))
                lib().clumpy_cuda_dev_free(self._h, C.c_void_p(arena["inbox"]))
                self._route_arena = None
            lib().clumpy_cuda_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(lib().clumpy_cuda_sync(self._h))

    def timer_start(self):
        _check(lib().clumpy_cuda_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        _check(lib().clumpy_cuda_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def ipc_export(self, dptr):
        """64-byte CUDA IPC handle of a device allocation (bytes), to hand to another process."""
        buf = C.create_string_buffer(64)
        _check(lib().clumpy_cuda_ipc_export(self._h, C.c_void_p(int(dptr)), buf))
        return buf.raw

    def ipc_open(self, handle):
        p = C.c_void_p()
        _check(lib().clumpy_cuda_ipc_open(self._h, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def ipc_close(self, dptr):
        _check(lib().clumpy_cuda_ipc_close(self._h, C.c_void_p(int(dptr))))

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        _check(lib().clumpy_cuda_dev_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, dptr):
        _check(lib().clumpy_cuda_dev_free(self._h, C.c_void_p(int(dptr))))

    def launch_count(self):
        n = C.c_uint64()
        _check(lib().clumpy_cuda_launch_count(self._h, C.byref(n)))
        return n.value

    def atomic_peak(self, op, pattern, region_bytes, nupdates, hot_cells=1, iters=5):
        """Mean ms of one launch of `nupdates` reductions (op 0: f16x2, 1: u32; pattern 0: uniform, 1: hot
        spot, 2: local) -- the atomic leg of the splat roofline (tools/atomic_peak.py)."""
        ms = C.c_float()
        _check(lib().clumpy_cuda_atomic_peak(self._h, op, pattern, region_bytes, nupdates, hot_cells, iters, C.byref(ms)))
        return ms.value

    # ---- commands, host buffers in and out ------------------------------------------------------

    def generate_simplex(self, width, height, amplitude, frequency, seed, out=None):
        """-> (H, W) float32, min, max   [clumpy generate_simplex <WxH> <amp> <freq> <seed>]"""
        out = np.empty((height, width), np.float32) if out is None else out
        lo, hi = C.c_float(), C.c_float()
        _check(lib().clumpy_cuda_generate_simplex(self._h, width, height, amplitude, frequency,
                                                  int(seed), _vp(out), C.byref(lo), C.byref(hi)))
        return out, lo.value, hi.value

    def curl_2d(self, src, out=None):
        """(H, W) float32 -> (H, W, 2) float32, min, max   [clumpy curl_2d]"""
        src = _f32(src)
        if src.ndim != 2:
            raise ValueError("Input data has wrong shape.")
        h, w = src.shape
        out = np.empty((h, w, 2), np.float32) if out is None else out
        lo, hi = C.c_float(), C.c_float()
        _check(lib().clumpy_cuda_curl_2d(self._h, _vp(src), w, h, _vp(out), C.byref(lo),
                                         C.byref(hi)))
        return out, lo.value, hi.value

    def pendulum_phase(self, width, height, friction, scale_x, scale_y):
        """-> (H, W, 2) float32   [clumpy pendulum_phase <WxH> <friction> <scale_x> <scale_y>]"""
        out = np.empty((height, width, 2), np.float32)
        _check(lib().clumpy_cuda_pendulum_phase(self._h, width, height, friction, scale_x, scale_y, _vp(out)))
        return out

    def pendulum_phase_dev(self, width, height, row0, nrows, friction, scale_x, scale_y, out_dev):
        _check(lib().clumpy_cuda_pendulum_phase_dev(self._h, width, height, row0, nrows, friction, scale_x,
                                                    scale_y, _vp(out_dev)))

    def cull_points(self, pts, mask):
        """(N, 2) points, (H, W) float32 mask -> the points over positive texels, in order   [clumpy cull_points]"""
        pts, mask = _f32(pts), _f32(mask)
        if pts.ndim != 2 or pts.shape[1] != 2:
            raise ValueError("Input points must be 2D.")
        if mask.ndim != 2:
            raise ValueError("Input data has wrong shape.")
        out = np.empty((max(len(pts), 1), 2), np.float32)
        kept = C.c_uint32()
        _check(lib().clumpy_cuda_cull_points(self._h, _vp(pts), len(pts), _vp(mask), mask.shape[1], mask.shape[0],
                                             _vp(out), C.byref(kept)))
        return out[:kept.value].copy()

    def splat_points(self, pts, width, height, img=None):
        """k=1, alpha=1 fast path of `clumpy splat_points` (splat_points.cc:108-116)."""
        pts = _f32(pts)
        img = np.zeros((height, width), np.uint8) if img is None else np.ascontiguousarray(img).copy()
        _check(lib().clumpy_cuda_splat_points(self._h, _vp(pts), len(pts), width, height, _vp(img)))
        return img

    def splat_disks(self, pts, width, height, alpha, kernel_size, wrapx=False, img=None):
        """splat_disks() of splat_points.cc:118-161 onto `img` (zeros when omitted)."""
        pts = _f32(pts)
        img = np.zeros((height, width), np.uint8) if img is None else np.ascontiguousarray(img).copy()
        _check(lib().clumpy_cuda_splat_disks(self._h, _vp(pts), len(pts), width, height, _vp(img),
                                             alpha, kernel_size, int(wrapx)))
        return img

    # ---- device-pointer variants (integers are raw device addresses) ----------------------------

    def generate_simplex_dev(self, width, height, row0, nrows, amplitude, frequency, seed, out_dev,
                             want_range=False):
        lo, hi = C.c_float(), C.c_float()
        _check(lib().clumpy_cuda_generate_simplex_dev(
            self._h, width, height, row0, nrows, amplitude, frequency, int(seed), _vp(out_dev),
            C.byref(lo) if want_range else None, C.byref(hi) if want_range else None))
        return (lo.value, hi.value) if want_range else None

    def age_offsets_dev(self, nframes, npts, out_dev):
        """`npts` reset offsets (advect_points.cc:127-135) drawn on the device into uint32 device memory."""
        _check(lib().clumpy_cuda_age_offsets_dev(self._h, nframes, npts, _vp(out_dev)))

    def generate_simplex_octaves_dev(self, width, height, row0, nrows, amplitudes, frequencies, seeds, out_dev, want_range=False):
        """Rows [row0, row0 + nrows) of the FP32 sum of the octaves (amplitude, frequency, seed), in that order."""
        n = len(amplitudes)
        amps, freqs, sds = (C.c_double * n)(*amplitudes), (C.c_double * n)(*frequencies), (C.c_int64 * n)(*[int(s) for s in seeds])
        lo, hi = C.c_float(), C.c_float()
        _check(lib().clumpy_cuda_generate_simplex_octaves_dev(
            self._h, width, height, row0, nrows, n, amps, freqs, sds, _vp(out_dev),
            C.byref(lo) if want_range else None, C.byref(hi) if want_range else None))
        return (lo.value, hi.value) if want_range else None

    def curl_2d_dev(self, src_dev, width, nrows, next_row_dev, dst_dev, want_range=False):
        lo, hi = C.c_float(), C.c_float()
        _check(lib().clumpy_cuda_curl_2d_dev(
            self._h, _vp(src_dev), width, nrows, _vp(next_row_dev), _vp(dst_dev),
            C.byref(lo) if want_range else None, C.byref(hi) if want_range else None))
        return (lo.value, hi.value) if want_range else None

    def advect(self, pts, vel, step_size, kernel_size, decay, nframes, age_offset=None, **options):
        """options: the fields of clumpy_advect_desc beyond the command's arguments -- cells ("auto" | "count" |
        "f16" | "u32" | "exact"), overlap, fuse, regroup, cell_rows_multiple, shard_rows=(lo, hi), and
        width / height when pts / vel / age_offset are raw device addresses (integers)."""
        return Advect(self, pts, vel, step_size, kernel_size, decay, nframes, age_offset, **options)

    def row_histogram(self, pts, height, npts=None):
        """Points per home row (numpy uint32, `height` entries); pts: (N, 2) array or a device address + npts."""
        on_device = isinstance(pts, (int, np.integer))
        if not on_device:
            pts = _f32(pts)
            npts = len(pts)
        rows = np.empty(height, np.uint32)
        _check(lib().clumpy_cuda_row_histogram(self._h, _vp(pts), npts, int(on_device), height, rows.ctypes.data_as(_u32p)))
        return rows

    def advect_points(self, pts, vel, step_size, kernel_size, decay, nframes, on_frame=None):
        """The whole `clumpy advect_points` command; returns (nframes, H, W) u8 unless on_frame
        (animframe, pixels) consumes the frames."""
        pts, vel = _f32(pts), _f32(vel)
        h, w = vel.shape[:2]
        wrapx, decay = (1, -decay) if decay < 0 else (0, decay)
        frames = None if on_frame else np.empty((nframes, h, w), np.uint8)

        def cb(_user, animframe, pixels):
            px = np.ctypeslib.as_array(pixels, shape=(h, w))
            if on_frame:
                return int(on_frame(animframe, px) or 0)
            frames[animframe] = px
            return 0

        fn = FRAME_FN(cb)
        _check(lib().clumpy_cuda_advect_points(self._h, _vp(pts), len(pts), _vp(vel), w, h,
                                               step_size, kernel_size, decay, wrapx, nframes,
                                               None, fn, None))
        return frames


class Advect:
    """Device-resident advect_points run (advect_points.cc:64-183).  A negative `decay` selects the
    reference's x-wrap mode, as the command does (:78-81).  pts / vel / age_offset may be numpy arrays (host)
    or integers (device addresses; then npts= / width= / height= say how big they are)."""

    def __init__(self, ctx, pts, vel, step_size, kernel_size, decay, nframes, age_offset=None, *, npts=None,
                 width=None, height=None, cells="auto", overlap=0, fuse=0, regroup=0, cell_rows_multiple=0,
                 shard_rows=None):
        self.ctx = ctx
        d = AdvectDesc()
        d.size = C.sizeof(AdvectDesc)
        self._keep = []
        if isinstance(pts, (int, np.integer)):
            d.pts, d.npts, d.pts_on_device = int(pts), int(npts), 1
        else:
            pts = _f32(pts)
            if pts.ndim != 2 or pts.shape[1] != 2:
                raise ValueError("Input points have wrong shape.")
            d.pts, d.npts, d.pts_on_device = pts.ctypes.data, len(pts), 0
            self._keep.append(pts)
        if isinstance(vel, (int, np.integer)):
            d.vel, d.vel_on_device, self.w, self.h = int(vel), 1, int(width), int(height)
        else:
            vel = _f32(vel)
            if vel.ndim != 3 or vel.shape[2] != 2:
                raise ValueError("Velocities have wrong shape.")
            d.vel, d.vel_on_device = vel.ctypes.data, 0
            self.h, self.w = vel.shape[:2]
            self._keep.append(vel)
        if age_offset is None:
            d.age_offset = None
        elif isinstance(age_offset, (int, np.integer)):
            d.age_offset, d.age_on_device = int(age_offset), 1
        else:
            age_offset = np.ascontiguousarray(age_offset, np.uint32)
            d.age_offset, d.age_on_device = age_offset.ctypes.data, 0
            self._keep.append(age_offset)
        d.width, d.height = self.w, self.h
        wrapx, decay = (1, -decay) if decay < 0 else (0, decay)
        d.step_size, d.kernel_size, d.decay, d.wrapx, d.nframes = step_size, kernel_size, decay, wrapx, nframes
        if shard_rows is not None:
            d.shard_row_lo, d.shard_row_hi = int(shard_rows[0]), int(shard_rows[1])
        d.cells, d.overlap, d.fuse, d.regroup, d.cell_rows_multiple = _CELLS[cells], overlap, fuse, regroup, cell_rows_multiple
        self.nframes = nframes
        self._h = C.c_void_p()
        _check(lib().clumpy_cuda_advect_begin_ex(ctx._h, C.byref(d), C.byref(self._h)))
        self._keep = []  # the library has synchronised: the inputs are on the device
        n = C.c_uint32()
        _check(lib().clumpy_cuda_advect_npts(self._h, C.byref(n)))
        self.n = n.value

    def step(self, count=1):
        _check(lib().clumpy_cuda_advect_step(self._h, count))

    def record(self, count, dst, frame_stride=None):
        """step(count), every frame copied to dst (pinned host memory: array or address) + i * frame_stride."""
        _check(lib().clumpy_cuda_advect_record(self._h, count, _vp(dst), frame_stride or self.w * self.h))

    def profile_stages(self, groups, serial=False):
        """-> (mean ms of one advect launch, mean ms of one group's composite work, frames per launch);
        serial: each kernel alone on the GPU instead of as the pipeline overlaps them."""
        a, c, nf = C.c_float(), C.c_float(), C.c_uint32()
        _check(lib().clumpy_cuda_advect_profile_stages(self._h, groups, int(serial), C.byref(a), C.byref(c), C.byref(nf)))
        return a.value, c.value, nf.value

    @property
    def simframe(self):
        s = C.c_uint32()
        _check(lib().clumpy_cuda_advect_simframe(self._h, C.byref(s)))
        return s.value

    # ---- one frame split at the exchange point (multi-GPU, see clumpy_b200/multigpu.py) -----------

    def count(self):
        _check(lib().clumpy_cuda_advect_count(self._h))

    def cells(self):
        """-> (device address, bytes, pitch in cells, rows, halo, bytes per cell) of the private
        counter image the last count() wrote."""
        p, nbytes = C.c_void_p(), C.c_size_t()
        pitch, rows, halo, bpc = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().clumpy_cuda_advect_cells(self._h, C.byref(p), C.byref(nbytes), C.byref(pitch),
                                              C.byref(rows), C.byref(halo), C.byref(bpc)))
        return p.value, nbytes.value, pitch.value, rows.value, halo.value, bpc.value

    def composite_rows(self, cells_dev, row0, nrows):
        _check(lib().clumpy_cuda_advect_composite_rows(self._h, _vp(cells_dev), row0, nrows))

    def finish_frame(self):
        _check(lib().clumpy_cuda_advect_finish_frame(self._h))

    # ---- exchange fused into the advect kernel (peer-memory inboxes) ----------------------------

    def route_begin(self, rank, world, capacity, arena=None):
        """-> (device address, bytes) of this rank's inbox: allocated by the run, or the caller's `arena` =
        (device address, bytes) kept from an earlier run."""
        p, n = C.c_void_p(arena[0] if arena else None), C.c_size_t(arena[1] if arena else 0)
        _check(lib().clumpy_cuda_advect_route_begin(self._h, rank, world, capacity, C.byref(p), C.byref(n)))
        return p.value, n.value

    def route_peers(self, inbox_ptrs):
        arr = (C.c_void_p * len(inbox_ptrs))(*[C.c_void_p(int(p) if p else 0) for p in inbox_ptrs])
        _check(lib().clumpy_cuda_advect_route_peers(self._h, arr))

    def route_steps(self, count):
        _check(lib().clumpy_cuda_advect_route_steps(self._h, count))

    def route_record(self, count, dst, frame_stride):
        """route_steps(count), this rank's rows of every frame copied to dst + i * frame_stride (pinned)."""
        _check(lib().clumpy_cuda_advect_route_record(self._h, count, _vp(dst), frame_stride))

    def route_profile(self, groups):
        """-> (route, drain incl. wait, band composite) mean ms per GROUP, frames per group."""
        a, b, c, nf = C.c_float(), C.c_float(), C.c_float(), C.c_uint32()
        _check(lib().clumpy_cuda_advect_route_profile(self._h, groups, C.byref(a), C.byref(b), C.byref(c), C.byref(nf)))
        return a.value, b.value, c.value, nf.value

    def route_errors(self):
        n = C.c_uint32()
        _check(lib().clumpy_cuda_advect_route_errors(self._h, C.byref(n)))
        return n.value

    def owned_rows(self):
        r0, n = C.c_uint32(), C.c_uint32()
        _check(lib().clumpy_cuda_advect_owned_rows(self._h, C.byref(r0), C.byref(n)))
        return r0.value, r0.value + n.value

    def image_dev(self):
        p = C.c_void_p()
        _check(lib().clumpy_cuda_advect_image_dev(self._h, C.byref(p)))
        return p.value

    def frame_async(self, dst):
        _check(lib().clumpy_cuda_advect_frame_async(self._h, _vp(dst)))

    def wait_frames(self):
        _check(lib().clumpy_cuda_advect_wait_frames(self._h))

    def points(self):
        out = np.empty((self.n, 2), np.float32)
        _check(lib().clumpy_cuda_advect_read_points(self._h, _vp(out)))
        return out

    def points_indexed(self):
        """Shards: (positions in storage order, index of each in the caller's list)."""
        pos, idx = np.empty((self.n, 2), np.float32), np.empty(self.n, np.uint32)
        _check(lib().clumpy_cuda_advect_read_points_indexed(self._h, _vp(pos), _vp(idx)))
        return pos, idx

    def image(self):
        out = np.empty((self.h, self.w), np.uint8)
        _check(lib().clumpy_cuda_advect_read_image(self._h, _vp(out)))
        return out

    def close(self):
        if self._h:
            lib().clumpy_cuda_advect_end(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
