"""ctypes binding of the C-ABI in include/sugarscape_b200.h (libsugarscape_b200.so).

The shared library is the drop-in boundary; this file is the thin host layer the
north star asks for.  If the CUDA library is missing the import of the engine fails
loudly -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

RNG_MT19937 = 0
RNG_KEYED = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
# (SS_B200_LIB: an A/B build of the same library, tools/build_alt.sh)
LIB_PATH = os.environ.get("SS_B200_LIB") or os.path.join(_HERE, "csrc", "libsugarscape_b200.so")


class SsConfig(C.Structure):
    """`ss_config` of include/sugarscape_b200.h."""
    _fields_ = [
        ("grid_size", C.c_int32), ("rule_grow", C.c_int32), ("rule_seasons", C.c_int32),
        ("rule_move_eat", C.c_int32), ("rule_utilicalc", C.c_int32), ("rule_can_starve", C.c_int32),
        ("rule_pollution", C.c_int32), ("season_period", C.c_int32),
        ("season_north", C.c_int32 * 4), ("season_south", C.c_int32 * 4),
        ("diffusion_rate", C.c_int32), ("pollution_start_time", C.c_int32),
        ("diffusion_start_time", C.c_int32), ("rng_mode", C.c_int32),
        ("grow_factor", C.c_double), ("season_on_factor", C.c_double), ("season_off_factor", C.c_double),
        ("pollution_a", C.c_double), ("pollution_b", C.c_double), ("max_capacity", C.c_double),
        ("keyed_seed", C.c_uint64), ("iteration", C.c_int64), ("env_time", C.c_int64),
        ("rule_limited_lifespan", C.c_int32), ("rule_procreate", C.c_int32), ("tags_length", C.c_int32),
        ("foresight_lo", C.c_int32), ("foresight_hi", C.c_int32), ("rule_spice", C.c_int32),
    ]


class SsStepStats(C.Structure):
    """`ss_step_stats` of include/sugarscape_b200.h."""
    _fields_ = [("population", C.c_double), ("metabolism_mean", C.c_double), ("vision_mean", C.c_double),
                ("gini", C.c_double), ("acted", C.c_double), ("deaths", C.c_double)]


class SsInitParams(C.Structure):
    """`ss_init_params` of include/sugarscape_b200.h."""
    _fields_ = [
        ("n_sites", C.c_int32), ("site_x", C.c_int32 * 4), ("site_y", C.c_int32 * 4), ("site_radius", C.c_int32 * 4),
        ("grow_to_capacity", C.c_int32), ("n_groups", C.c_int32), ("group_count", C.c_int32 * 4),
        ("max_metabolism", C.c_int32), ("max_vision", C.c_int32), ("endowment_min", C.c_int32), ("endowment_max", C.c_int32),
        ("age_min", C.c_int32), ("age_max", C.c_int32), ("childbearing", (C.c_int32 * 4) * 2),
        ("foresight_lo", C.c_int32), ("foresight_hi", C.c_int32), ("group_tags", C.c_int32 * 4),
        ("n_spice_sites", C.c_int32), ("spice_site_x", C.c_int32 * 4), ("spice_site_y", C.c_int32 * 4), ("spice_site_radius", C.c_int32 * 4),
    ]


STATS_DTYPE = np.dtype([("population", "f8"), ("metabolism_mean", "f8"), ("vision_mean", "f8"),
                        ("gini", "f8"), ("acted", "f8"), ("deaths", "f8")])

#: every symbol include/sugarscape_b200.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "ss_create", "ss_destroy", "ss_last_error", "ss_upload_cells", "ss_upload_agents",
    "ss_set_rng_mt19937", "ss_get_rng_mt19937", "ss_set_rng_keyed", "ss_step", "ss_agent_count",
    "ss_download_cells", "ss_download_agents", "ss_get_counters", "ss_get_kernel_times",
    "ss_set_option", "ss_version", "ss_shard_enable", "ss_shard_begin_step", "ss_shard_pass", "ss_shard_apply",
    "ss_shard_apply_gathered", "ss_shard_pending", "ss_shard_end_step", "ss_host_alloc", "ss_host_free", "ss_seed_mt19937", "ss_init_world",
    "ss_create_strip", "ss_strip_init_directory", "ss_strip_export", "ss_strip_connect", "ss_strip_ipc_export",
    "ss_strip_ipc_connect", "ss_strip_step", "ss_strip_phase",
    "ss_upload_agents_life", "ss_download_agents_life", "ss_event_count", "ss_get_events",
    "ss_upload_cells_spice", "ss_download_cells_spice", "ss_upload_agents_spice", "ss_download_agents_spice",
)

_lib = None


class EngineLibraryMissing(ImportError):
    pass


def lib():
    """Load libsugarscape_b200.so (built by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineLibraryMissing(
            "CUDA engine library not built: %s (run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.ss_version.restype = C.c_char_p
    L.ss_last_error.restype = C.c_char_p
    L.ss_last_error.argtypes = [vp]
    L.ss_create.argtypes = [C.POINTER(SsConfig), C.c_int32, C.POINTER(vp)]
    L.ss_destroy.argtypes = [vp]
    L.ss_destroy.restype = None
    L.ss_upload_cells.argtypes = [vp, vp, vp, vp]
    L.ss_upload_agents.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.ss_set_rng_mt19937.argtypes = [vp, vp, C.c_int32]
    L.ss_get_rng_mt19937.argtypes = [vp, vp, C.POINTER(C.c_int32)]
    L.ss_set_rng_keyed.argtypes = [vp, C.c_uint64]
    L.ss_seed_mt19937.argtypes = [vp, C.c_uint64]
    L.ss_init_world.argtypes = [vp, C.POINTER(SsInitParams)]
    L.ss_step.argtypes = [vp, C.c_int32, vp]
    L.ss_agent_count.argtypes = [vp]
    L.ss_download_cells.argtypes = [vp, vp, vp, vp, vp]
    L.ss_download_agents.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.ss_get_counters.argtypes = [vp, vp, C.c_int32]
    L.ss_get_kernel_times.argtypes = [vp, vp, C.c_int32]
    L.ss_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.ss_shard_enable.argtypes = [vp, C.c_int32, C.c_int32]
    L.ss_shard_begin_step.argtypes = [vp]
    L.ss_shard_pass.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(vp),
                                C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(vp)]
    L.ss_shard_apply.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64]
    L.ss_shard_apply_gathered.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_int32, C.c_uint64, C.c_uint64]
    L.ss_shard_pending.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.ss_shard_end_step.argtypes = [vp, vp]
    L.ss_create_strip.argtypes = [C.POINTER(SsConfig), C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.ss_strip_init_directory.argtypes = [vp, C.c_int32, vp, vp, vp, vp]
    L.ss_strip_export.argtypes = [vp, vp]
    L.ss_strip_connect.argtypes = [vp, C.c_int32, vp]
    L.ss_strip_ipc_export.argtypes = [vp, vp]
    L.ss_strip_ipc_connect.argtypes = [vp, C.c_int32, vp]
    L.ss_strip_step.argtypes = [vp, C.c_int32, vp]
    L.ss_strip_phase.argtypes = [vp, C.c_int32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), vp]
    L.ss_upload_agents_life.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ss_download_agents_life.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int32)]
    L.ss_event_count.argtypes = [vp, C.POINTER(C.c_int32)]
    L.ss_get_events.argtypes = [vp, C.c_int32, vp]
    L.ss_upload_cells_spice.argtypes = [vp, vp, vp]
    L.ss_download_cells_spice.argtypes = [vp, vp, vp]
    L.ss_upload_agents_spice.argtypes = [vp, C.c_int32, vp, vp, vp]
    L.ss_download_agents_spice.argtypes = [vp, vp, vp, vp]
    L.ss_host_alloc.argtypes = [C.c_size_t]
    L.ss_host_alloc.restype = vp
    L.ss_host_free.argtypes = [vp]
    L.ss_host_free.restype = None
    _lib = L
    return L
