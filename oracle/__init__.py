"""ctypes doors to the CHECKERS: oracle/liboracle.so (our C restatement, clumpy_oracle.c) and, when
it was built, oracle/_ref/libclumpy_ref.so + oracle/_ref/clumpy_ref (the unmodified reference).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from clumpy_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


def build(with_ref=True):
    """Compile liboracle.so (and _ref/ when /root/reference is present) with oracle/Makefile."""
    target = ["all"] if with_ref else ["liboracle.so"]
    subprocess.run(["make", "-s", "-C", _HERE, "-j8"] + target, check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(with_ref=False)
        L = C.CDLL(path)
        L.oracle_simplex_perm.argtypes = [C.c_int64, C.POINTER(C.c_int16)]
        L.oracle_simplex2.argtypes = [C.POINTER(C.c_int16), C.c_double, C.c_double]
        L.oracle_simplex2.restype = C.c_double
        L.oracle_generate_simplex.argtypes = [C.c_uint32, C.c_uint32, C.c_double, C.c_double,
                                              C.c_int64, _f32p, _f32p, _f32p]
        L.oracle_curl_2d.argtypes = [_f32p, C.c_uint32, C.c_uint32, _f32p, _f32p, _f32p]
        L.oracle_sprite.argtypes = [C.c_int, C.c_float, _f32p]
        L.oracle_splat_points.argtypes = [_f32p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p]
        L.oracle_splat_disks.argtypes = [_f32p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p,
                                         C.c_float, C.c_int, C.c_int]
        L.oracle_splat_disks.restype = C.c_int
        L.oracle_age_offsets.argtypes = [C.c_uint32, C.c_uint32, _u32p]
        L.oracle_advect_new.argtypes = [_f32p, C.c_uint32, _f32p, C.c_uint32, C.c_uint32,
                                        C.c_float, C.c_int, C.c_float, C.c_uint32]
        L.oracle_advect_new.restype = C.c_void_p
        L.oracle_advect_free.argtypes = [C.c_void_p]
        L.oracle_advect_step.argtypes = [C.c_void_p]
        L.oracle_advect_step.restype = C.c_int
        L.oracle_advect_points.argtypes = [C.c_void_p]
        L.oracle_advect_points.restype = _f32p
        L.oracle_advect_image.argtypes = [C.c_void_p]
        L.oracle_advect_image.restype = _u8p
        L.oracle_advect_simframe.argtypes = [C.c_void_p]
        L.oracle_advect_simframe.restype = C.c_uint32
        L.oracle_bridson_capacity.argtypes = [C.c_uint32, C.c_uint32, C.c_float]
        L.oracle_bridson_capacity.restype = C.c_uint32
        L.oracle_bridson_points.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_int, _f32p]
        L.oracle_bridson_points.restype = C.c_uint32
        L.oracle_pendulum_phase.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, _f32p]
        L.oracle_cull_points.argtypes = [_f32p, C.c_uint32, _f32p, C.c_uint32, C.c_uint32, _f32p]
        L.oracle_cull_points.restype = C.c_uint32
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t):
    return a.ctypes.data_as(t)


# ---- restatement ---------------------------------------------------------------------------

def simplex_perm(seed):
    p = np.zeros(256, np.int16)
    lib().oracle_simplex_perm(int(seed), _ptr(p, C.POINTER(C.c_int16)))
    return p


def simplex2(perm, x, y):
    perm = np.ascontiguousarray(perm, np.int16)
    return lib().oracle_simplex2(_ptr(perm, C.POINTER(C.c_int16)), float(x), float(y))


def generate_simplex(width, height, amplitude, frequency, seed):
    """-> (H, W) float32, min, max   [generate_simplex.cc:43-87]"""
    out = np.empty((height, width), np.float32)
    lo, hi = C.c_float(), C.c_float()
    lib().oracle_generate_simplex(width, height, amplitude, frequency, int(seed),
                                  _ptr(out, _f32p), C.byref(lo), C.byref(hi))
    return out, lo.value, hi.value


def generate_simplex_rows(width, height, row0, nrows, amplitude, frequency, seed):
    """Rows [row0, row0 + nrows) of generate_simplex's (H, W) image -> (nrows, W) float32, min, max."""
    out = np.empty((nrows, width), np.float32)
    lib().oracle_generate_simplex_rows(width, height, row0, nrows, C.c_double(amplitude), C.c_double(frequency),
                                       C.c_int64(int(seed)), _ptr(out, _f32p))
    return out, float(out.min()) if out.size else 0.0, float(out.max()) if out.size else 0.0


def curl_2d(src):
    """(H, W) float32 -> (H, W, 2) float32, min, max   [curl_2d.cc:29-77]"""
    src = _f32(src)
    h, w = src.shape
    out = np.empty((h, w, 2), np.float32)
    lo, hi = C.c_float(), C.c_float()
    lib().oracle_curl_2d(_ptr(src, _f32p), w, h, _ptr(out, _f32p), C.byref(lo), C.byref(hi))
    return out, lo.value, hi.value


def sprite(kernel_size, alpha=1.0):
    s = np.empty(kernel_size * kernel_size, np.float32)
    lib().oracle_sprite(kernel_size, alpha, _ptr(s, _f32p))
    return s


def splat_points(pts, width, height, img=None):
    pts = _f32(pts)
    img = np.zeros((height, width), np.uint8) if img is None else np.ascontiguousarray(img).copy()
    lib().oracle_splat_points(_ptr(pts, _f32p), len(pts), width, height, _ptr(img, _u8p))
    return img


def splat_disks(pts, width, height, alpha, kernel_size, wrapx=False, img=None):
    pts = _f32(pts)
    img = np.zeros((height, width), np.uint8) if img is None else np.ascontiguousarray(img).copy()
    rc = lib().oracle_splat_disks(_ptr(pts, _f32p), len(pts), width, height, _ptr(img, _u8p),
                                  alpha, kernel_size, int(wrapx))
    if rc:
        raise ValueError("Kernel size must be an odd integer.")
    return img


def age_offsets(nframes, npts):
    out = np.empty(npts, np.uint32)
    lib().oracle_age_offsets(nframes, npts, _ptr(out, _u32p))
    return out


def bridson_points(width, height, radius, seed):
    """-> (N, 2) float32   [bridson_points.cc:38-177]"""
    cap = lib().oracle_bridson_capacity(width, height, radius)
    out = np.empty((max(cap, 1), 2), np.float32)
    n = lib().oracle_bridson_points(width, height, radius, int(seed), _ptr(out, _f32p))
    return out[:n].copy()


def pendulum_phase(width, height, friction, scalex, scaley):
    """-> (H, W, 2) float32   [pendulum_phase.cc:38-71]"""
    out = np.empty((height, width, 2), np.float32)
    lib().oracle_pendulum_phase(width, height, friction, scalex, scaley, _ptr(out, _f32p))
    return out


def cull_points(pts, mask):
    """(N, 2) points, (H, W) mask -> kept points in order   [cull_points.cc:97-111]"""
    pts, mask = _f32(pts), _f32(mask)
    h, w = mask.shape
    out = np.empty((max(len(pts), 1), 2), np.float32)
    n = lib().oracle_cull_points(_ptr(pts, _f32p), len(pts), _ptr(mask, _f32p), w, h, _ptr(out, _f32p))
    return out[:n].copy()


class Advect:
    """Stateful restatement of advect_points.cc:64-183; step() is one simulation frame."""

    def __init__(self, pts, vel, step_size, kernel_size, decay, nframes):
        self.pts0 = _f32(pts)
        self.vel = _f32(vel)
        self.h, self.w = self.vel.shape[:2]
        self.n = len(self.pts0)
        self.nframes = nframes
        self._st = lib().oracle_advect_new(_ptr(self.pts0, _f32p), self.n, _ptr(self.vel, _f32p),
                                           self.w, self.h, step_size, kernel_size, decay, nframes)

    def step(self):
        return lib().oracle_advect_step(self._st)

    @property
    def simframe(self):
        return lib().oracle_advect_simframe(self._st)

    def points(self):
        p = lib().oracle_advect_points(self._st)
        return np.ctypeslib.as_array(p, shape=(self.n, 2)).copy()

    def image(self):
        p = lib().oracle_advect_image(self._st)
        return np.ctypeslib.as_array(p, shape=(self.h, self.w)).copy()

    def close(self):
        if self._st:
            lib().oracle_advect_free(self._st)
            self._st = None

    def __del__(self):
        self.close()


def advect_points(pts, vel, step_size, kernel_size, decay, nframes):
    """Full command: returns (frames (nframes, H, W) u8, final points)."""
    a = Advect(pts, vel, step_size, kernel_size, decay, nframes)
    frames = np.empty((nframes, a.h, a.w), np.uint8)
    k = 0
    while True:
        r = a.step()
        if r < 0:
            break
        if r == 1:
            frames[k] = a.image()
            k += 1
    pts_out = a.points()
    a.close()
    return frames, pts_out


# ---- the compiled reference (present only where oracle/Makefile could build it) --------------

REF_CLI = os.path.join(_HERE, "_ref", "clumpy_ref")
REF_SO = os.path.join(_HERE, "_ref", "libclumpy_ref.so")


def have_ref():
    return os.path.exists(REF_CLI) and os.path.exists(REF_SO)


def ref():
    global _REF
    if _REF is None:
        R = C.CDLL(REF_SO)
        R.ref_splat_disks.argtypes = [_f32p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, C.c_float,
                                      C.c_int, C.c_int]
        R.ref_splat_points.argtypes = [_f32p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p]
        R.ref_simplex_perm.argtypes = [C.c_int64, C.POINTER(C.c_int16)]
        R.ref_simplex2_many.argtypes = [C.c_int64, C.POINTER(C.c_double), C.c_uint32,
                                        C.POINTER(C.c_double)]
        R.ref_exec.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p)]
        R.ref_exec.restype = C.c_int
        R.ref_std_age_offsets.argtypes = [C.c_uint32, C.c_uint32, _u32p]
        _REF = R
    return _REF


def ref_cli(*args, cwd=None, quiet=True):
    """Run `clumpy_ref <args...>` (the reference CLI); raises on a non-zero exit code."""
    out = subprocess.run([REF_CLI] + [str(a) for a in args], cwd=cwd, check=True,
                         stdout=subprocess.PIPE if quiet else None)
    return out.stdout.decode(errors="replace") if quiet else ""


def ref_splat_disks(pts, width, height, alpha, kernel_size, wrapx=False, img=None):
    pts = _f32(pts)
    img = np.zeros((height, width), np.uint8) if img is None else np.ascontiguousarray(img).copy()
    ref().ref_splat_disks(_ptr(pts, _f32p), len(pts), width, height, _ptr(img, _u8p), alpha,
                          kernel_size, int(wrapx))
    return img


def ref_std_age_offsets(nframes, npts):
    """advect_points.cc:127-135 through <random> itself (std::mt19937 + std::uniform_int_distribution)."""
    out = np.empty(npts, np.uint32)
    ref().ref_std_age_offsets(nframes, npts, _ptr(out, _u32p))
    return out


def ref_simplex2(seed, xy):
    xy = np.ascontiguousarray(xy, np.float64)
    out = np.empty(len(xy), np.float64)
    ref().ref_simplex2_many(int(seed), _ptr(xy, C.POINTER(C.c_double)), len(xy),
                            _ptr(out, C.POINTER(C.c_double)))
    return out
