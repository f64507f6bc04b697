"""ctypes binding of libaltair_b200.so (include/altair_b200.h).

This is plumbing: argument marshalling and error translation.  All compute happens in the CUDA
library; if it is missing or no GPU is present the calls raise -- there is no CPU fallback.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("ALTB_LIB") or os.path.join(_PKG, "libaltair_b200.so")   # ALTB_LIB: experiment builds

EXITED, ABSORBED, SUSPENDED, TAPE_END = 1, 2, 3, 4
MAP_LINE, MAP_TRACEONCE_COMPAT, MAP_DIRECTION, MAP_PER_POSITION, MAP_TWOFOLD = 0, 1, 2, 3, 4
CONTRACT_EXACT, CONTRACT_FAST, CONTRACT_FAST7 = 0, 1, 2

RECORD_DTYPE = np.dtype([("pos", "<f4", 3), ("dir", "<f4", 3), ("n_hits", "<u4"), ("status", "<u4")])


class AltbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"altair_b200 error {code}: {msg}")
        self.code = code


class Scene(C.Structure):
    """altb_scene; defaults of scene() follow flux_at_observer/fluxAtObserverFast.C:33-41,192-230."""
    _fields_ = [("r_inner", C.c_double), ("r_outer", C.c_double), ("theta_max_deg", C.c_double),
                ("world_half", C.c_double), ("reflectance", C.c_double), ("roughness_rad", C.c_double),
                ("lambertian", C.c_int32), ("max_bounces", C.c_int32), ("brdf_kind", C.c_int32),
                ("count_all_status", C.c_int32), ("brdf_param", C.c_double * 4), ("exit_z", C.c_double)]


class Source(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("dir", C.c_double * 3)]


class MapSpec(C.Structure):
    _fields_ = [("n_theta", C.c_int32), ("n_phi", C.c_int32), ("det_radius", C.c_double),
                ("det_width", C.c_double), ("map_mode", C.c_int32), ("rays_per_position", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_rays", C.c_uint64), ("n_exited", C.c_uint64), ("n_exit_port", C.c_uint64),
                ("n_absorbed", C.c_uint64), ("n_suspended", C.c_uint64), ("n_bounces", C.c_uint64),
                ("t_trace_s", C.c_double), ("t_map_s", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def scene(theta_max=170.0, world_half=300.0, reflectance=0.99, roughness=0.01, max_bounces=50000,
          r_inner=100.1, r_outer=101.0, lambertian=1, brdf_kind=0, brdf_param=(0.3, 0.4, 0.6, 0.0),
          count_all_status=0, exit_z=-100.0):
    s = Scene()
    s.r_inner, s.r_outer, s.theta_max_deg, s.world_half = r_inner, r_outer, theta_max, world_half
    s.reflectance, s.roughness_rad = reflectance, roughness
    s.lambertian, s.max_bounces, s.brdf_kind, s.count_all_status = lambertian, max_bounces, brdf_kind, count_all_status
    for i in range(4):
        s.brdf_param[i] = brdf_param[i]
    s.exit_z = exit_z
    return s


def source(pos=(-60.0, 0.0, -75.0), direction=(5.0, 0.0, 0.0)):
    s = Source()
    for i in range(3):
        s.pos[i] = pos[i]
        s.dir[i] = direction[i]
    return s


def map_spec(n_theta=180, n_phi=90, det_radius=100.0, det_width=40.0, mode=MAP_LINE, rays_per_position=0):
    m = MapSpec()
    m.n_theta, m.n_phi, m.det_radius, m.det_width, m.map_mode, m.rays_per_position = n_theta, n_phi, det_radius, det_width, mode, rays_per_position
    return m


def library_path():
    return _LIB_PATH


def build_library(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (csrc/Makefile); cross-compiles without a GPU."""
    csrc = os.path.join(_PKG, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", "Makefile"))]
    srcs.append(os.path.join(os.path.dirname(_PKG), "include", "altair_b200.h"))
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", csrc, "-B"], stdout=None if verbose else subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def load_library():
    """dlopen the CUDA library.  Raises if it has not been built -- never falls back to anything."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise AltbError(-4, f"{_LIB_PATH} not built; run __graft_entry__.build() (there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    P = C.POINTER
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    L.altb_last_error.restype = C.c_char_p
    L.altb_create.argtypes = [P(vp), P(C.c_int), C.c_int]
    L.altb_destroy.argtypes = [vp]
    L.altb_destroy.restype = None
    L.altb_set_batch.argtypes = [vp, u64]
    L.altb_collective.argtypes = [vp]
    L.altb_set_contract.argtypes = [vp, C.c_int]
    L.altb_get_contract.argtypes = [vp]
    L.altb_launch_count.argtypes = [vp]
    L.altb_launch_count.restype = u64
    L.altb_trace_launch_count.argtypes = [vp]
    L.altb_trace_launch_count.restype = u64
    L.altb_trace_fluxmap.argtypes = [vp, P(Scene), C.c_int, P(Source), u64, u64, u64, P(MapSpec), vp, P(Stats)]
    L.altb_trace_fluxmap_dev.argtypes = [vp, P(Scene), C.c_int, P(Source), u64, u64, u64, P(MapSpec), vp, vp, vp]
    L.altb_trace_exit_rays.argtypes = [vp, P(Scene), P(Source), u64, u64, u64, vp, vp, vp, vp, P(Stats)]
    L.altb_trace_records.argtypes = [vp, P(Scene), P(Source), u64, u64, u64, vp, P(Stats)]
    L.altb_detector_sweep.argtypes = [vp, P(Scene), P(Source), u64, u64, u64, vp, vp, u32, C.c_double, C.c_double,
                                      vp, P(Stats)]
    L.altb_replay.argtypes = [vp, P(Scene), vp, vp, vp, u64, P(MapSpec), vp, vp, vp]
    L.altb_replay_ex.argtypes = [vp, P(Scene), vp, vp, vp, u64, P(MapSpec), u32, vp, vp, vp]
    L.altb_map_records.argtypes = [vp, P(Scene), P(MapSpec), vp, u64, vp]
    L.altb_map_records_at.argtypes = [vp, P(Scene), P(MapSpec), vp, u64, u64, vp]
    L.altb_probe_f32.argtypes = [vp, C.c_int, vp, u64, vp]
    L.altb_draws.argtypes = [vp, u64, u64, u64, u32, vp]
    L.altb_draws_lobe.argtypes = [vp, u64, u64, u64, u32, C.c_int, C.c_double, vp]
    L.altb_trace_paths.argtypes = [vp, P(Scene), P(Source), u64, u64, u64, u32, vp, vp, vp]
    L.altb_measure_fp32_peak.argtypes = [vp, P(C.c_double)]
    L.altb_count_horizon.argtypes = [vp, P(Scene), P(Source), u64, u64, u64, P(u64), P(u64), P(u64)]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """altb_ctx: owns streams and device buffers of one or more GPUs of this process."""

    def __init__(self, devices=None):
        self._L = load_library()
        self._h = C.c_void_p()
        if devices is None:
            rc = self._L.altb_create(C.byref(self._h), None, 1)
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self._L.altb_create(C.byref(self._h), arr, len(devices))
        self._check(rc)

    def _check(self, rc):
        if rc != 0:
            raise AltbError(rc, self._L.altb_last_error().decode())

    def close(self):
        if self._h:
            self._L.altb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def launches(self):
        return int(self._L.altb_launch_count(self._h))

    @property
    def collective(self):
        """How a multi-device context merges its maps: "none" (one device), "nccl" or "host"."""
        return ("none", "nccl", "host")[int(self._L.altb_collective(self._h))]

    @property
    def trace_launches(self):
        return int(self._L.altb_trace_launch_count(self._h))

    def set_contract(self, contract):
        """CONTRACT_EXACT (bit-exact vs the CPU oracle, default), CONTRACT_FAST (special-function unit) or CONTRACT_FAST7 (+ Philox4x32-7); include/altair_b200.h."""
        self._check(self._L.altb_set_contract(self._h, int(contract)))

    @property
    def contract(self):
        return int(self._L.altb_get_contract(self._h))

    def set_batch(self, batch_rays):
        self._check(self._L.altb_set_batch(self._h, int(batch_rays)))

    # -- the hot path -------------------------------------------------------------------------
    def trace_fluxmap(self, scenes, src, n_rays, mp, seed=4357, ray_id0=0, counts=None):
        """-> (counts[n_scenes, n_theta*n_phi] uint64, [stats dict per scene]); host buffers."""
        if isinstance(scenes, Scene):
            scenes = [scenes]
        ns = len(scenes)
        arr = (Scene * ns)(*scenes)
        nb = mp.n_theta * mp.n_phi
        if counts is None:
            counts = np.zeros((ns, nb), dtype=np.uint64)
        st = (Stats * ns)()
        self._check(self._L.altb_trace_fluxmap(self._h, arr, ns, C.byref(src), ray_id0, n_rays, seed, C.byref(mp),
                                               _ptr(counts), st))
        return counts, [s.as_dict() for s in st]

    def trace_fluxmap_dev(self, scenes, src, n_rays, mp, d_counts_ptr, d_stats_ptr, seed=4357, ray_id0=0, stream=0):
        """Asynchronous, device-resident variant: pointers are raw device addresses (e.g. tensor.data_ptr())."""
        if isinstance(scenes, Scene):
            scenes = [scenes]
        ns = len(scenes)
        arr = (Scene * ns)(*scenes)
        self._check(self._L.altb_trace_fluxmap_dev(self._h, arr, ns, C.byref(src), ray_id0, n_rays, seed,
                                                   C.byref(mp), C.c_void_p(d_counts_ptr),
                                                   C.c_void_p(d_stats_ptr) if d_stats_ptr else None,
                                                   C.c_void_p(stream) if stream else None))

    def trace_records(self, sc, src, n_rays, seed=4357, ray_id0=0):
        rec = np.zeros(n_rays, dtype=RECORD_DTYPE)
        st = Stats()
        self._check(self._L.altb_trace_records(self._h, C.byref(sc), C.byref(src), ray_id0, n_rays, seed, _ptr(rec),
                                               C.byref(st)))
        return rec, st.as_dict()

    def trace_exit_rays(self, sc, src, n_rays, seed=4357, ray_id0=0):
        pos = np.zeros((n_rays, 3)); d = np.zeros((n_rays, 3))
        npts = np.zeros(n_rays, dtype=np.uint32); status = np.zeros(n_rays, dtype=np.uint8)
        st = Stats()
        self._check(self._L.altb_trace_exit_rays(self._h, C.byref(sc), C.byref(src), ray_id0, n_rays, seed, _ptr(pos),
                                                 _ptr(d), _ptr(npts), _ptr(status), C.byref(st)))
        return pos, d, npts, status, st.as_dict()

    def detector_sweep(self, sc, src, n_rays, centers, rots, det_r=5.0, det_halfthick=0.1, seed=4357, ray_id0=0):
        centers = np.ascontiguousarray(centers, dtype=np.float64)
        rots = np.ascontiguousarray(rots, dtype=np.float64)
        m = len(centers)
        hits = np.zeros(m, dtype=np.uint64)
        st = Stats()
        self._check(self._L.altb_detector_sweep(self._h, C.byref(sc), C.byref(src), ray_id0, n_rays, seed,
                                                _ptr(centers), _ptr(rots), m, det_r, det_halfthick, _ptr(hits),
                                                C.byref(st)))
        return hits, st.as_dict()

    def replay(self, sc, ray0, tape, tape_off, mp=None, full_azimuth=False):
        ray0 = np.ascontiguousarray(ray0, dtype=np.float64)
        tape = np.ascontiguousarray(tape, dtype=np.float32)
        tape_off = np.ascontiguousarray(tape_off, dtype=np.uint64)
        n = len(tape_off) - 1
        rec = np.zeros(n, dtype=RECORD_DTYPE)
        bins = np.full(n, -1, dtype=np.int32)
        port = np.zeros(n, dtype=np.uint8)
        self._check(self._L.altb_replay_ex(self._h, C.byref(sc), _ptr(ray0), _ptr(tape), _ptr(tape_off), n,
                                           C.byref(mp) if mp is not None else None, 1 if full_azimuth else 0, _ptr(rec),
                                           _ptr(bins) if mp is not None else None, _ptr(port)))
        return rec, bins, port

    def map_records(self, sc, mp, rec, ray_id0=0):
        """Map stage alone; rec[i] is the ray with global id ray_id0 + i (matters for the per-position modes)."""
        rec = np.ascontiguousarray(rec)
        counts = np.zeros(mp.n_theta * mp.n_phi, dtype=np.uint64)
        self._check(self._L.altb_map_records_at(self._h, C.byref(sc), C.byref(mp), _ptr(rec), ray_id0, len(rec), _ptr(counts)))
        return counts

    def draws(self, seed, ray_id0, n, k, lobe_n=0, lobe_deg=0.0):
        out = np.zeros((n, 8), dtype=np.float32)
        self._check(self._L.altb_draws_lobe(self._h, seed, ray_id0, n, k, lobe_n, lobe_deg, _ptr(out)))
        return out

    def probe_f32(self, op, x):
        """y = op(x) with the kernels' f32 primitives (0 sqrt, 1 reciprocal, 2 log, 3/4 table sin/cos of 2 pi x / 2^20)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.zeros_like(x)
        self._check(self._L.altb_probe_f32(self._h, op, _ptr(x), x.size, _ptr(y)))
        return y

    def measure_fp32_peak(self):
        v = C.c_double()
        self._check(self._L.altb_measure_fp32_peak(self._h, C.byref(v)))
        return v.value

    def count_horizon(self, sc, src, n_rays, seed=4357, ray_id0=0):
        """(hits whose roughness-tilted normal no longer faces the incoming ray, rays with at least one, all hits)."""
        e, r, h = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._L.altb_count_horizon(self._h, C.byref(sc), C.byref(src), ray_id0, n_rays, seed,
                                               C.byref(e), C.byref(r), C.byref(h)))
        return e.value, r.value, h.value

    def trace_paths(self, sc, src, n_rays, max_points, seed=4357, ray_id0=0):
        """Polylines for small N: (points[n, max_points, 3] f32, n_points[n], status[n])."""
        pts = np.zeros((n_rays, max_points, 3), dtype=np.float32)
        npts = np.zeros(n_rays, dtype=np.uint32); status = np.zeros(n_rays, dtype=np.uint8)
        self._check(self._L.altb_trace_paths(self._h, C.byref(sc), C.byref(src), ray_id0, n_rays, seed, max_points,
                                             _ptr(pts), _ptr(npts), _ptr(status)))
        return pts, npts, status
