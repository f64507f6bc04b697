"""Python access to the C++ macro mirror (libaltair_macros.so): the reference's entry points by their own names.

    from altair_raytracing_b200 import macros
    macros.set("traceonce_rays", 100000)
    macros.sweepDetectorTraceOnce(False, "results", 1, -60, 0, -75, 5, 0, 0, 164.0)     # fluxAtObserverFast.C:1068
    print(macros.last_csv())

Everything here is marshalling; the work happens in C++ (altb_macros.cpp) over the C ABI.
"""
import ctypes as _C
import os as _os

_PKG = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
_lib = None


def _L():
    global _lib
    if _lib is None:
        from ..binding import library_path, load_library
        load_library()
        _C.CDLL(library_path(), mode=_C.RTLD_GLOBAL)
        path = _os.path.join(_PKG, "libaltair_macros.so")
        if not _os.path.exists(path):
            raise RuntimeError(f"{path} not built; run __graft_entry__.build()")
        L = _C.CDLL(path)
        d = _C.c_double
        for name in ("altbm_sweepDetector", "altbm_sweepDetectorTwofold", "altbm_sweepDetectorTraceOnce"):
            getattr(L, name).argtypes = [_C.c_int, _C.c_char_p, _C.c_int, d, d, d, d, d, d, d]
        L.altbm_set.argtypes = [_C.c_char_p, d]
        L.altbm_set_output_dir.argtypes = [_C.c_char_p]
        L.altbm_last_csv.restype = _C.c_char_p
        L.altbm_last_count.argtypes = [_C.c_char_p]
        L.altbm_last_count.restype = _C.c_longlong
        L.altbm_last_fluxmap.argtypes = [_C.c_int, _C.c_int]
        L.altbm_last_fluxmap.restype = d
        _lib = L
    return _lib


def set(key, value):                                   # noqa: A001  (mirrors altb_macros::Settings)
    if _L().altbm_set(key.encode(), float(value)) != 0:
        raise KeyError(key)


def set_output_dir(path):
    _L().altbm_set_output_dir(str(path).encode())


def last_csv():
    return _L().altbm_last_csv().decode()


def last_count(what):
    return int(_L().altbm_last_count(what.encode()))


def last_fluxmap(bin_x, bin_y):
    """TH2D::GetBinContent(bin_x, bin_y) of the last map (1-based, as fluxMap->SetBinContent(i+1, j+1, f))."""
    return float(_L().altbm_last_fluxmap(bin_x, bin_y))


def _sweep(fn, notify, saveFolder, threads, srcX, srcY, srcZ, dirX, dirY, dirZ, thetaMax):
    getattr(_L(), fn)(int(bool(notify)), saveFolder.encode(), int(threads), srcX, srcY, srcZ, dirX, dirY, dirZ, thetaMax)


def sweepDetector(notify=True, saveFolder="results", threads=-1, srcX=-60.0, srcY=0.0, srcZ=-80.0, dirX=5.0, dirY=2.0, dirZ=0.0,
                  thetaMax=170.0):
    """fluxAtObserverOptimize.C:433 / fluxAtObserverFast.C sweepDetector."""
    _sweep("altbm_sweepDetector", notify, saveFolder, threads, srcX, srcY, srcZ, dirX, dirY, dirZ, thetaMax)


def sweepDetectorTwofold(notify=True, saveFolder="results", threads=-1, srcX=-60.0, srcY=0.0, srcZ=-80.0, dirX=5.0, dirY=2.0,
                         dirZ=0.0, thetaMax=170.0):
    """fluxAtObserverFast.C:518."""
    _sweep("altbm_sweepDetectorTwofold", notify, saveFolder, threads, srcX, srcY, srcZ, dirX, dirY, dirZ, thetaMax)


def sweepDetectorTraceOnce(notify=True, saveFolder="results", threads=-1, srcX=-60.0, srcY=0.0, srcZ=-80.0, dirX=5.0, dirY=2.0,
                           dirZ=0.0, thetaMax=170.0):
    """fluxAtObserverFast.C:1068."""
    _sweep("altbm_sweepDetectorTraceOnce", notify, saveFolder, threads, srcX, srcY, srcZ, dirX, dirY, dirZ, thetaMax)


def sweepSeries(which="fast"):
    """fluxAtObserverFast.C:1641 ("fast") or fluxAtObserverOptimize.C:892 ("optimize")."""
    (_L().altbm_sweepSeriesFast if which == "fast" else _L().altbm_sweepSeriesOptimize)()


def makeIntegratingSphereNRays():
    _L().altbm_makeIntegratingSphereNRays()


def integratingSphereDetectorSweep():
    _L().altbm_integratingSphereDetectorSweep()


def distributionSphereDetectorSweep():
    _L().altbm_distributionSphereDetectorSweep()


def nonLambertianFlux_sweepDetector():
    _L().altbm_nonLambertianFlux_sweepDetector()


def fluxAtObserver_sweepDetector():
    _L().altbm_fluxAtObserver_sweepDetector()
