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
    "clumpy_cuda_simplex_perm": (C.c_int, [C.c_int64, _i16p]),
    "clumpy_cuda_generate_simplex": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_double,
                                               C.c_double, C.c_int64, C.c_void_p, _f32p, _f32p]),
    "clumpy_cuda_generate_simplex_dev": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                   C.c_uint32, C.c_double, C.c_double, C.c_int64,
                                                   C.c_void_p, _f32p, _f32p]),
    "clumpy_cuda_curl_2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                      _f32p, _f32p]),
    "clumpy_cuda_curl_2d_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                          C.c_void_p, C.c_void_p, _f32p, _f32p]),
    "clumpy_cuda_splat_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                           C.c_uint32, C.c_void_p]),
    "clumpy_cuda_splat_disks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                          C.c_uint32, C.c_void_p, C.c_float, C.c_int, C.c_int]),
    "clumpy_cuda_age_offsets": (C.c_int, [C.c_uint32, C.c_uint32, _u32p]),
    "clumpy_cuda_advect_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                           C.c_uint32, C.c_uint32, C.c_float, C.c_int, C.c_float,
                                           C.c_int, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_advect_step": (C.c_int, [C.c_void_p, C.c_uint32]),
    "clumpy_cuda_advect_profile": (C.c_int, [C.c_void_p, C.c_uint32, _f32p, _f32p, _f32p]),
    "clumpy_cuda_advect_profile_fused": (C.c_int, [C.c_void_p, C.c_uint32, _f32p, _u32p]),
    "clumpy_cuda_advect_count": (C.c_int, [C.c_void_p]),
    "clumpy_cuda_advect_cells": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                           _u32p, _u32p, _u32p, _u32p]),
    "clumpy_cuda_advect_composite_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]),
    "clumpy_cuda_advect_finish_frame": (C.c_int, [C.c_void_p]),
    "clumpy_cuda_advect_image_dev": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "clumpy_cuda_ipc_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clumpy_cuda_advect_route_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint32,
                                                 C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "clumpy_cuda_advect_route_peers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_advect_route": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "clumpy_cuda_advect_drain": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clumpy_cuda_advect_owned_rows": (C.c_int, [C.c_void_p, _u32p, _u32p]),
    "clumpy_cuda_advect_route_steps": (C.c_int, [C.c_void_p, C.c_uint32]),
    "clumpy_cuda_advect_route_errors": (C.c_int, [C.c_void_p, _u32p]),
    "clumpy_cuda_pendulum_phase": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "clumpy_cuda_pendulum_phase_dev": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float,
                                                 C.c_float, C.c_float, C.c_void_p]),
    "clumpy_cuda_cull_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                          C.c_void_p, _u32p]),
    "clumpy_cuda_advect_route_profile": (C.c_int, [C.c_void_p, C.c_uint32, _f32p, _f32p, _f32p]),
    "clumpy_cuda_advect_simframe": (C.c_int, [C.c_void_p, _u32p]),
    "clumpy_cuda_advect_frame_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clumpy_cuda_advect_wait_frames": (C.c_int, [C.c_void_p]),
    "clumpy_cuda_advect_read_points": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clumpy_cuda_advect_read_image": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clumpy_cuda_advect_end": (C.c_int, [C.c_void_p]),
    "clumpy_cuda_advect_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                            C.c_uint32, C.c_uint32, C.c_float, C.c_int, C.c_float,
                                            C.c_int, C.c_uint32, C.c_void_p, FRAME_FN, C.c_void_p]),
}

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

    def launch_count(self):
        n = C.c_uint64()
        _check(lib().clumpy_cuda_launch_count(self._h, C.byref(n)))
        return n.value

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

    def curl_2d_dev(self, src_dev, width, nrows, next_row_dev, dst_dev, want_range=False):
        lo, hi = C.c_float(), C.c_float()
        _check(lib().clumpy_cuda_curl_2d_dev(
            self._h, _vp(src_dev), width, nrows, _vp(next_row_dev), _vp(dst_dev),
            C.byref(lo) if want_range else None, C.byref(hi) if want_range else None))
        return (lo.value, hi.value) if want_range else None

    def advect(self, pts, vel, step_size, kernel_size, decay, nframes, age_offset=None):
        return Advect(self, pts, vel, step_size, kernel_size, decay, nframes, age_offset)

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
    reference's x-wrap mode, as the command does (:78-81)."""

    def __init__(self, ctx, pts, vel, step_size, kernel_size, decay, nframes, age_offset=None):
        self.ctx = ctx
        pts, vel = _f32(pts), _f32(vel)
        if pts.ndim != 2 or pts.shape[1] != 2:
            raise ValueError("Input points have wrong shape.")
        if vel.ndim != 3 or vel.shape[2] != 2:
            raise ValueError("Velocities have wrong shape.")
        self.h, self.w = vel.shape[:2]
        self.n = len(pts)
        self.nframes = nframes
        wrapx, decay = (1, -decay) if decay < 0 else (0, decay)
        if age_offset is not None:
            age_offset = np.ascontiguousarray(age_offset, np.uint32)
        self._h = C.c_void_p()
        _check(lib().clumpy_cuda_advect_begin(ctx._h, _vp(pts), self.n, _vp(vel), self.w, self.h,
                                              step_size, kernel_size, decay, wrapx, nframes,
                                              _vp(age_offset), C.byref(self._h)))

    def step(self, count=1):
        _check(lib().clumpy_cuda_advect_step(self._h, count))

    def profile(self, count):
        """-> mean ms per frame of (advect+count kernel, composite kernel, counter clear)."""
        a, c, z = C.c_float(), C.c_float(), C.c_float()
        _check(lib().clumpy_cuda_advect_profile(self._h, count, C.byref(a), C.byref(c), C.byref(z)))
        return a.value, c.value, z.value

    def profile_fused(self, groups):
        """-> (mean ms of one advect launch inside the running pipeline, frames per launch)."""
        ms, nf = C.c_float(), C.c_uint32()
        _check(lib().clumpy_cuda_advect_profile_fused(self._h, groups, C.byref(ms), C.byref(nf)))
        return ms.value, nf.value

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

    def route_begin(self, rank, world, capacity):
        """-> (device address, bytes) of this rank's inbox allocation."""
        p, n = C.c_void_p(), C.c_size_t()
        _check(lib().clumpy_cuda_advect_route_begin(self._h, rank, world, capacity, C.byref(p), C.byref(n)))
        return p.value, n.value

    def route_peers(self, inbox_ptrs):
        arr = (C.c_void_p * len(inbox_ptrs))(*[C.c_void_p(int(p) if p else 0) for p in inbox_ptrs])
        _check(lib().clumpy_cuda_advect_route_peers(self._h, arr))

    def route(self):
        """Advect + route this frame; -> device address of the world per-destination counts."""
        p = C.c_void_p()
        _check(lib().clumpy_cuda_advect_route(self._h, C.byref(p)))
        return p.value

    def drain(self, counts_all_dev):
        _check(lib().clumpy_cuda_advect_drain(self._h, _vp(counts_all_dev)))

    def route_steps(self, count):
        _check(lib().clumpy_cuda_advect_route_steps(self._h, count))

    def route_profile(self, count):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _check(lib().clumpy_cuda_advect_route_profile(self._h, count, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

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
