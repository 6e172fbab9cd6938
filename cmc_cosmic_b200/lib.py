"""ctypes binding of libcmcb200.so (include/cmcb200.h).  Fails loudly when the CUDA library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CMCB200_LIB") or os.path.join(_HERE, "libcmcb200.so")   # CMCB200_LIB: another build of the same library (A/B measurements)
N_STAGES = 19
STAGE_NAMES = ["keys_hist", "radix_passes", "record_gather", "binaries", "potential", "orbit_prepare", "transfer", "energy", "h2d", "d2h",
               "orbit_scan", "orbit_tries", "orbit_tables", "shard_keys", "shard_rank", "shard_tables", "lagrange", "dynamics", "sigma"]


class OrbitParams(C.Structure):
    _fields_ = [("n_local", C.c_long), ("n_old_local", C.c_long), ("first_global", C.c_long), ("N_MAX", C.c_long), ("N_STAR", C.c_long),
                ("n_table", C.c_long), ("cenma_m", C.c_double), ("Rtidal", C.c_double), ("draws", C.c_void_p), ("draw_off", C.c_void_p),
                ("seed", C.c_ulonglong), ("tcount", C.c_ulonglong), ("fuse_compute_energy", C.c_int)]


class OrbitResult(C.Structure):
    _fields_ = [("max_r", C.c_double), ("n_draws", C.c_ulonglong), ("n_exact", C.c_ulonglong), ("n_brent", C.c_ulonglong),
                ("n_status", C.c_ulonglong * 6), ("n_deferred", C.c_ulonglong)]


class MassPcParams(C.Structure):
    _fields_ = [("count", C.c_long), ("mass_pc", C.c_void_p), ("bh_include", C.c_int), ("minimum_r", C.c_double), ("initial_total_mass", C.c_double),
                ("Mtotal", C.c_double), ("cenma_m", C.c_double), ("N_MAX", C.c_long), ("N_STAR", C.c_long), ("n_local", C.c_long), ("first_global", C.c_long)]


class MassPcResult(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in ("mass_r", "ave_mass_r", "no_star_r", "densities_r", "ke_rad_r", "ke_tan_r", "v2_rad_r", "v2_tan_r", "owner")] + \
               [("totals", C.c_double * 4), ("filled", C.c_long)]


class StepResult(C.Structure):
    _fields_ = [("n_local", C.c_long), ("n_bin", C.c_long), ("N_MAX", C.c_long), ("cenma_m", C.c_double), ("cenma_m_new", C.c_double),
                ("Mtotal", C.c_double), ("Rtidal", C.c_double)]


class DynamicsParams(C.Structure):
    _fields_ = [(k, C.c_long) for k in ("n_local", "first_global", "N_MAX", "N_STAR")] + [("dt", C.c_double)] + \
               [(k, C.c_long) for k in ("relaxation", "ss_collision", "binsingle", "binbin", "tidal_capture", "tc_polytrope", "bh_capture",
                                        "bh_loss_cone", "avekernel", "honor_threebb")] + \
               [(k, C.c_double) for k in ("cenma_m", "gamma", "coll_factor", "tc_factor", "wd_tc", "units_mstar", "units_l", "units_t")] + \
               [("draws", C.c_void_p), ("seed", C.c_ulonglong), ("tcount", C.c_ulonglong)]


class SigmaParams(C.Structure):
    _fields_ = [(k, C.c_long) for k in ("p", "n_local", "first_global", "N_MAX", "N_STAR")] + [("r_0_mave_1", C.c_int), ("force_sequential", C.c_int)] + \
               [(k, C.c_void_p) for k in ("ghost_prev_vr", "ghost_prev_vt", "ghost_next_vr", "ghost_next_vt")]


class SnapshotParams(C.Structure):
    _fields_ = [(k, C.c_long) for k in ("n_local", "n_old_local", "first_global", "N_MAX", "N_STAR")] + [(k, C.c_double) for k in ("units_m", "units_l", "madhoc")] + \
               [("avekernel", C.c_long), ("nan_value", C.c_double), ("bh_only", C.c_int)]


class DynamicsResult(C.Structure):
    _fields_ = [("n_pairs", C.c_long), ("n_encounters", C.c_long), ("Nrel", C.c_long), ("Nrelbeta", C.c_long * 4),
                ("qaverelbeta", C.c_double * 4), ("maverelbeta", C.c_double * 4), ("raverelbeta", C.c_double * 4)]


class CmcB200Error(RuntimeError):
    pass


_lib = None

EXPORTS = [
    "cmcb200_create", "cmcb200_create_sharded", "cmcb200_shard_handle_bytes", "cmcb200_shard_export", "cmcb200_shard_attach", "cmcb200_shard_detach",
    "cmcb200_shard_fuse_transfer", "cmcb200_shard_sort_local", "cmcb200_shard_exchange", "cmcb200_shard_publish", "cmcb200_shard_finish", "cmcb200_destroy", "cmcb200_last_error", "cmcb200_set_stream", "cmcb200_layout", "cmcb200_find_limits",
    "cmcb200_calc_potential_new", "cmcb200_qsorts_new", "cmcb200_new_orbits_calculate", "cmcb200_upload", "cmcb200_download",
    "cmcb200_dev_new_orbits", "cmcb200_dev_transfer_new_positions", "cmcb200_dev_qsorts_new", "cmcb200_dev_transfer_and_sort", "cmcb200_dev_calc_potential_new",
    "cmcb200_dev_compute_energy", "cmcb200_dev_read_status", "cmcb200_sync", "cmcb200_profile", "cmcb200_profile_read",
    "cmcb200_potential_stats", "cmcb200_launch_count", "cmcb200_host_register", "cmcb200_host_unregister", "cmcb200_transfer_bytes",
    "cmcb200_dev_comp_mass_percent", "cmcb200_dev_clusdyn_calculate", "cmcb200_dev_energy_conservation2", "cmcb200_dev_compute_intermediate_energy",
    "cmcb200_dev_energy_conservation1", "cmcb200_dev_set_velocities3", "cmcb200_dev_set_velocities3_pass", "cmcb200_dev_energy_new_radii",
    "cmcb200_dev_potential", "cmcb200_dev_tidally_strip_select", "cmcb200_dev_compute_energy_totals",
    "cmcb200_shard_step_async", "cmcb200_shard_step_result", "cmcb200_dev_new_orbits_async", "cmcb200_dev_orbit_result", "cmcb200_resident_begin", "cmcb200_resident_begin_sharded", "cmcb200_resident_end", "cmcb200_dev_get", "cmcb200_dev_put", "cmcb200_dev_comp_multi_mass_percent",
    "cmcb200_dev_dynamics_apply", "cmcb200_dev_dynamics_encounters", "cmcb200_dev_calc_sigma_r", "cmcb200_dev_sigma_arrays", "cmcb200_restart_write", "cmcb200_dev_simul_relax", "cmcb200_dev_step", "cmcb200_dev_snapshot_rows",
]


def load():
    """dlopen the library and declare the prototypes.  No GPU is needed for this (cudart initialises lazily)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CmcB200Error(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, lg, dbl, pl = C.c_void_p, C.c_long, C.c_double, C.POINTER(C.c_long)
    L.cmcb200_create.argtypes = [C.POINTER(vp), C.c_int, lg, lg]
    L.cmcb200_create_sharded.argtypes = [C.POINTER(vp), C.c_int, lg, lg, lg]
    L.cmcb200_shard_export.argtypes = [vp, C.c_int, C.c_int, vp]
    L.cmcb200_shard_attach.argtypes = [vp, vp]
    L.cmcb200_shard_detach.argtypes = [vp]
    L.cmcb200_shard_fuse_transfer.argtypes = [vp, C.c_int]
    L.cmcb200_shard_sort_local.argtypes = [vp, lg, pl]
    L.cmcb200_shard_exchange.argtypes = [vp, vp, vp, vp, lg, lg, lg]
    L.cmcb200_shard_publish.argtypes = [vp, lg, lg, lg, pl]
    L.cmcb200_shard_finish.argtypes = [vp]
    L.cmcb200_destroy.argtypes = [vp]
    L.cmcb200_destroy.restype = None
    L.cmcb200_last_error.restype = C.c_char_p
    L.cmcb200_set_stream.argtypes = [vp, vp]
    L.cmcb200_layout.argtypes = [vp]
    L.cmcb200_layout.restype = None
    L.cmcb200_find_limits.argtypes = [lg, C.c_int, C.c_int, vp, vp]
    L.cmcb200_find_limits.restype = None
    L.cmcb200_calc_potential_new.argtypes = [vp, vp, vp, vp, pl, lg, C.POINTER(dbl), C.POINTER(dbl), dbl, C.POINTER(dbl), C.POINTER(dbl)]
    L.cmcb200_qsorts_new.argtypes = [vp, vp, pl, vp, lg, pl, vp, vp]
    L.cmcb200_new_orbits_calculate.argtypes = [vp, vp, vp, vp, vp, C.POINTER(OrbitParams), vp, vp, vp, C.POINTER(OrbitResult)]
    L.cmcb200_upload.argtypes = [vp, vp, lg, vp, lg, vp, vp, vp, lg]
    L.cmcb200_download.argtypes = [vp, vp, lg, vp, lg, vp, vp, vp, lg]
    L.cmcb200_dev_new_orbits.argtypes = [vp, C.POINTER(OrbitParams), C.c_int, C.POINTER(OrbitResult)]
    L.cmcb200_dev_transfer_new_positions.argtypes = [vp, lg, lg, lg, lg]
    L.cmcb200_dev_qsorts_new.argtypes = [vp, pl, lg, pl]
    L.cmcb200_dev_transfer_and_sort.argtypes = [vp, pl, lg, lg, lg, lg, pl]
    L.cmcb200_dev_calc_potential_new.argtypes = [vp, pl, lg, C.POINTER(dbl), C.POINTER(dbl), dbl, C.POINTER(dbl), C.POINTER(dbl)]
    L.cmcb200_dev_compute_energy.argtypes = [vp, lg, lg]
    L.cmcb200_dev_read_status.argtypes = [vp, vp, vp, vp, lg]
    L.cmcb200_sync.argtypes = [vp]
    L.cmcb200_profile.argtypes = [vp, C.c_int]
    L.cmcb200_profile_read.argtypes = [vp, vp, vp, C.c_int]
    L.cmcb200_potential_stats.argtypes = [vp, pl, pl]
    L.cmcb200_launch_count.argtypes = [vp]
    L.cmcb200_launch_count.restype = lg
    L.cmcb200_host_register.argtypes = [vp, C.c_size_t]
    L.cmcb200_host_unregister.argtypes = [vp]
    L.cmcb200_transfer_bytes.argtypes = [vp, pl, pl, C.c_int]
    L.cmcb200_dev_energy_conservation2.argtypes = [vp, C.POINTER(OrbitParams)]
    L.cmcb200_dev_compute_intermediate_energy.argtypes = [vp, C.POINTER(OrbitParams)]
    L.cmcb200_dev_energy_conservation1.argtypes = [vp, C.POINTER(OrbitParams)]
    L.cmcb200_dev_energy_new_radii.argtypes = [vp, C.POINTER(OrbitParams)]
    L.cmcb200_dev_compute_energy_totals.argtypes = [vp, lg, lg, lg, lg, dbl, vp]
    L.cmcb200_dev_potential.argtypes = [vp, C.POINTER(OrbitParams), vp, lg, vp]
    L.cmcb200_dev_tidally_strip_select.argtypes = [vp, C.POINTER(OrbitParams), C.c_int, dbl, vp, lg, pl, vp, vp, vp]
    L.cmcb200_dev_set_velocities3.argtypes = [vp, C.POINTER(OrbitParams), C.POINTER(dbl)]
    L.cmcb200_dev_set_velocities3_pass.argtypes = [vp, C.POINTER(OrbitParams), dbl, C.POINTER(dbl)]
    L.cmcb200_dev_clusdyn_calculate.argtypes = [vp, lg, lg, dbl, C.POINTER(dbl)]
    L.cmcb200_dev_comp_mass_percent.argtypes = [vp, C.POINTER(MassPcParams), C.POINTER(MassPcResult)]
    L.cmcb200_resident_begin.argtypes = [vp, vp, lg, vp, lg, vp, vp, vp, lg, C.c_ulonglong]
    L.cmcb200_shard_step_async.argtypes = [vp, lg, lg, lg, lg, C.c_int, lg, lg, dbl, dbl]
    L.cmcb200_shard_step_result.argtypes = [vp, vp, vp, vp, pl, vp]
    L.cmcb200_dev_new_orbits_async.argtypes = [vp, C.POINTER(OrbitParams), C.c_int]
    L.cmcb200_dev_orbit_result.argtypes = [vp, C.POINTER(OrbitParams), C.POINTER(OrbitResult)]
    L.cmcb200_dev_comp_multi_mass_percent.argtypes = [vp, lg, vp, lg, dbl, vp, lg, vp]
    L.cmcb200_dev_dynamics_apply.argtypes = [vp, C.POINTER(DynamicsParams), C.POINTER(DynamicsResult)]
    L.cmcb200_dev_dynamics_encounters.argtypes = [vp, lg, vp, vp, vp, vp, vp, vp, pl]
    L.cmcb200_dev_calc_sigma_r.argtypes = [vp, C.POINTER(SigmaParams), vp, vp, vp]
    L.cmcb200_dev_sigma_arrays.argtypes = [vp, vp, vp]
    L.cmcb200_restart_write.argtypes = [vp, C.c_char_p, vp, lg, vp, lg, vp, lg, vp, lg, lg, vp, lg, lg, vp, lg]
    L.cmcb200_dev_simul_relax.argtypes = [vp, lg, lg, lg, lg, lg, dbl, C.POINTER(dbl)]
    L.cmcb200_dev_step.argtypes = [vp, C.POINTER(OrbitParams), lg, dbl, dbl, C.POINTER(OrbitResult), C.POINTER(StepResult)]
    L.cmcb200_dev_snapshot_rows.argtypes = [vp, C.POINTER(SnapshotParams), vp, vp, lg, C.POINTER(lg)]
    L.cmcb200_dev_get.argtypes = [vp, vp, lg, vp, lg]
    L.cmcb200_dev_put.argtypes = [vp, vp, lg, vp, lg, vp, lg]
    L.cmcb200_resident_begin_sharded.argtypes = [vp, vp, lg, vp, lg, vp, vp, vp, lg, C.c_ulonglong, lg, lg]
    L.cmcb200_resident_end.argtypes = [vp, vp, lg, vp, lg, lg, vp, vp, vp, lg]
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        raise CmcB200Error(f"{what}: {load().cmcb200_last_error().decode()}")
