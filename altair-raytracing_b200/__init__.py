"""altair-raytracing_b200 -- B200-native integrating-sphere photon tracer (hot path only).

Host-side mirror of the reference's macro interface over the C ABI in include/altair_b200.h.
The directory name carries a hyphen (it is the project name); import it as
``altair_raytracing_b200`` (the sibling alias package re-exports everything).
"""
from .binding import (  # noqa: F401
    ABSORBED, CONTRACT_EXACT, CONTRACT_FAST, CONTRACT_FAST7, EXITED, MAP_DIRECTION, MAP_LINE, MAP_PER_POSITION, MAP_TRACEONCE_COMPAT, MAP_TWOFOLD, RECORD_DTYPE, SUSPENDED, TAPE_END,
    AltbError, Context, MapSpec, Scene, Source, Stats, build_library, library_path, load_library, map_spec,
    scene, source,
)
from . import macros  # noqa: F401
