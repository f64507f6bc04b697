"""ctypes loader for the CPU oracle (oracle/libaltair_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libaltair_oracle.so")

EXITED, ABSORBED, SUSPENDED, TAPE_END = 1, 2, 3, 4
MAP_LINE, MAP_TRACEONCE_COMPAT, MAP_DIRECTION, MAP_PER_POSITION, MAP_TWOFOLD = 0, 1, 2, 3, 4
F64, F32 = 0, 1

RECORD_DTYPE = np.dtype([("pos", "<f4", 3), ("dir", "<f4", 3), ("n_hits", "<u4"), ("status", "<u4")])


class Scene(C.Structure):
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
    """Defaults = flux_at_observer/fluxAtObserverFast.C:33-41,192-230."""
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


def build(force=False):
    """Compile the oracle (building the checker is not using it)."""
    srcs = [os.path.join(_HERE, f) for f in ("altair_oracle.c", "oracle_core.inc", "altair_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB)
            and os.path.getmtime(_LIB) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB
    subprocess.check_call(["make", "-C", _HERE, "-B", "libaltair_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        build()
    L = C.CDLL(_LIB)
    P = C.POINTER
    L.orc_philox4x32_10.argtypes = [P(C.c_uint32), P(C.c_uint32), P(C.c_uint32)]
    L.orc_philox4x32_7.argtypes = [P(C.c_uint32), P(C.c_uint32), P(C.c_uint32)]
    L.orc_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, P(C.c_float)]
    L.orc_draws_lobe.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_float, P(C.c_float)]
    L.orc_sincos2pi_f32.argtypes = [C.c_float, P(C.c_float), P(C.c_float)]
    L.orc_sincos_f32.argtypes = [C.c_float, P(C.c_float), P(C.c_float)]
    L.orc_sincos2pi_q13.argtypes = [C.c_uint32, P(C.c_float), P(C.c_float)]
    L.orc_sincos2pi_q20.argtypes = [C.c_uint32, P(C.c_float), P(C.c_float)]
    L.orc_log_u20.argtypes = [C.c_uint32]
    L.orc_log_u20.restype = C.c_float
    L.orc_trace.argtypes = [P(Scene), P(Source), C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p,
                            P(Stats), C.c_int]
    L.orc_trace_f64.argtypes = [P(Scene), P(Source), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int]
    L.orc_replay.argtypes = [P(Scene), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    L.orc_replay_ex.argtypes = [P(Scene), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p]
    L.orc_make_tape.argtypes = [P(Scene), P(Source), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64,
                                C.c_void_p]
    L.orc_make_tape.restype = C.c_int64
    L.orc_map_records.argtypes = [P(Scene), P(MapSpec), C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int]
    L.orc_fluxmap.argtypes = [P(Scene), P(Source), C.c_uint64, C.c_uint64, C.c_uint64, P(MapSpec), C.c_int,
                              C.c_void_p, P(Stats), C.c_int]
    L.orc_direction_bin.argtypes = [P(MapSpec), P(C.c_float)]
    L.orc_direction_bin.restype = C.c_int32
    L.orc_detector_pose.argtypes = [C.c_double, C.c_double, C.c_double, P(C.c_double), P(C.c_double)]
    L.orc_detector_hit.argtypes = [P(C.c_double), P(C.c_double), C.c_double, P(C.c_double), P(C.c_double)]
    L.orc_disk_hits.argtypes = [P(Scene), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double,
                                C.c_double, C.c_void_p]
    L.orc_sweep_pose.argtypes = [C.c_double, C.c_double, C.c_double, P(C.c_double), P(C.c_double)]
    L.orc_trace_paths.argtypes = [P(Scene), P(Source), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_port_flag.argtypes = [P(Scene), C.c_void_p]
    L.orc_count_horizon.argtypes = [P(Scene), P(Source), C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                                    P(C.c_uint64), P(C.c_uint64), P(C.c_uint64)]
    L.orc_num_threads.restype = C.c_int
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def philox(ctr, key, rounds=10):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    (lib().orc_philox4x32_10 if rounds == 10 else lib().orc_philox4x32_7)(c, k, o)
    return list(o)


def draws(seed, ray_id, k):
    o = (C.c_float * 8)()
    lib().orc_draws(seed, ray_id, k, o)
    return np.array(list(o), dtype=np.float32)


def trace(sc, src, n, seed=4357, ray_id0=0, prec=F32, n_threads=0, want_records=True):
    rec = np.zeros(n if want_records else 0, dtype=RECORD_DTYPE)
    st = Stats()
    rc = lib().orc_trace(C.byref(sc), C.byref(src), ray_id0, n, seed, prec,
                         _ptr(rec) if want_records else None, C.byref(st), n_threads)
    if rc:
        raise RuntimeError(f"orc_trace rc={rc}")
    return rec, st.as_dict()


def trace_f64(sc, src, n, seed=4357, ray_id0=0, n_threads=0):
    pos = np.zeros((n, 3)); d = np.zeros((n, 3))
    nh = np.zeros(n, dtype=np.uint32); st = np.zeros(n, dtype=np.uint8)
    rc = lib().orc_trace_f64(C.byref(sc), C.byref(src), ray_id0, n, seed, _ptr(pos), _ptr(d), _ptr(nh), _ptr(st),
                             n_threads)
    if rc:
        raise RuntimeError(f"orc_trace_f64 rc={rc}")
    return pos, d, nh, st


def make_tape(sc, src, n, seed=4357, ray_id0=0):
    off = np.zeros(n + 1, dtype=np.uint64)
    total = lib().orc_make_tape(C.byref(sc), C.byref(src), ray_id0, n, seed, None, 0, _ptr(off))
    if total < 0:
        raise RuntimeError(f"orc_make_tape rc={total}")
    tape = np.zeros((max(total, 1), 8), dtype=np.float32)
    got = lib().orc_make_tape(C.byref(sc), C.byref(src), ray_id0, n, seed, _ptr(tape), total, _ptr(off))
    if got != total:
        raise RuntimeError(f"orc_make_tape rc={got}")
    return tape[:total], off


def replay(sc, ray0, tape, tape_off, prec=F32, full_azimuth=False):
    n = len(tape_off) - 1
    ray0 = np.ascontiguousarray(ray0, dtype=np.float64)
    tape = np.ascontiguousarray(tape, dtype=np.float32)
    tape_off = np.ascontiguousarray(tape_off, dtype=np.uint64)
    rec = np.zeros(n, dtype=RECORD_DTYPE)
    rc = lib().orc_replay_ex(C.byref(sc), _ptr(ray0), _ptr(tape), _ptr(tape_off), n, prec, 1 if full_azimuth else 0, _ptr(rec))
    if rc:
        raise RuntimeError(f"orc_replay rc={rc}")
    return rec


def map_records(sc, mp, rec, prec=F32, n_threads=0):
    counts = np.zeros(mp.n_theta * mp.n_phi, dtype=np.uint64)
    rec = np.ascontiguousarray(rec)
    rc = lib().orc_map_records(C.byref(sc), C.byref(mp), _ptr(rec), len(rec), prec, _ptr(counts), n_threads)
    if rc:
        raise RuntimeError(f"orc_map_records rc={rc}")
    return counts


def fluxmap(sc, src, n, mp, seed=4357, ray_id0=0, prec=F32, n_threads=0):
    counts = np.zeros(mp.n_theta * mp.n_phi, dtype=np.uint64)
    st = Stats()
    rc = lib().orc_fluxmap(C.byref(sc), C.byref(src), ray_id0, n, seed, C.byref(mp), prec, _ptr(counts),
                           C.byref(st), n_threads)
    if rc:
        raise RuntimeError(f"orc_fluxmap rc={rc}")
    return counts, st.as_dict()


def port_flags(sc, rec):
    ok = (rec["status"] == EXITED) | bool(sc.count_all_status)
    return ok & (rec["pos"][:, 2] < np.float32(sc.exit_z))


def sweep_pose(theta, phi, r=200.0):
    c = (C.c_double * 3)(); m = (C.c_double * 9)()
    lib().orc_sweep_pose(theta, phi, r, c, m)
    return np.array(list(c)), np.array(list(m))


def disk_hits(sc, rec, centers, rots, det_r=5.0, det_halfthick=0.1):
    centers = np.ascontiguousarray(centers, dtype=np.float64)
    rots = np.ascontiguousarray(rots, dtype=np.float64)
    m = len(centers)
    hits = np.zeros(m, dtype=np.uint64)
    rec = np.ascontiguousarray(rec)
    rc = lib().orc_disk_hits(C.byref(sc), _ptr(rec), len(rec), _ptr(centers), _ptr(rots), m, det_r, det_halfthick,
                             _ptr(hits))
    if rc:
        raise RuntimeError(f"orc_disk_hits rc={rc}")
    return hits


def set_philox_rounds(rounds):
    """10 (default) or 7 (mirror of the library's CONTRACT_FAST7); process-wide."""
    assert lib().orc_set_philox_rounds(int(rounds)) == 0


def count_horizon(sc, src, n, seed=4357, ray_id0=0, prec=F32):
    e, r, h = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = lib().orc_count_horizon(C.byref(sc), C.byref(src), ray_id0, n, seed, prec, C.byref(e), C.byref(r), C.byref(h))
    assert rc == 0, rc
    return e.value, r.value, h.value


def trace_paths(sc, src, n, max_points, seed=4357, ray_id0=0):
    pts = np.zeros((n, max_points, 3), dtype=np.float32)
    npts = np.zeros(n, dtype=np.uint32); status = np.zeros(n, dtype=np.uint8)
    rc = lib().orc_trace_paths(C.byref(sc), C.byref(src), ray_id0, n, seed, max_points, _ptr(pts), _ptr(npts), _ptr(status))
    if rc:
        raise RuntimeError(f"orc_trace_paths rc={rc}")
    return pts, npts, status
