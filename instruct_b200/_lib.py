"""ctypes binding of the C-ABI in include/instruct_b200.h.

The shared library is built in-tree (instruct_b200/libinstruct_b200.so) by
``make -C instruct_b200/csrc`` / ``__graft_entry__.build()``.  There is no fallback: if
the library is missing, or no CUDA device is usable, every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IG_LIB") or os.path.join(HERE, "libinstruct_b200.so")     # IG_LIB: kernel-lab builds (tools/kernel_lab.sh)

IG_OK = 0
IG_EMPTY_CLUSTER = 1

# ig_state_id
STATE_X, STATE_Z, STATE_Q, STATE_P, STATE_ALPHA, STATE_S, STATE_G, STATE_INDVLKH, STATE_TOTALLKH, \
    STATE_TALLY, STATE_CNT, STATE_GPROP, STATE_STATE, STATE_MASK, STATE_GENO, STATE_ITER = range(16)
STATE_TABLES, STATE_TABLES_PROP, STATE_EXFREQ, STATE_SPROP, STATE_DSTAT, STATE_GMAX = 16, 17, 18, 19, 20, 21
STATE_P2, STATE_TALLY2 = 22, 23      # allotetraploid: second subgenome
STATE_LLPARTS, STATE_GEOMETRY, STATE_FPROP, STATE_FK = 100, 101, 102, 103
STATE_DPWEIGHTS, STATE_DPCLUSTERS = 104, 105     # mode 3, DP prior: gen_post_prob weights per individual; number of clusters
# ig_phase
PHASE_UPDATE_P, PHASE_UPDATE_S, PHASE_ZQ, PHASE_ALPHA, PHASE_GENO = 1, 2, 4, 8, 16

# every symbol include/instruct_b200.h declares (tests check the built library exports them all)
EXPORTS = [
    "ig_version", "ig_last_error", "ig_device_count", "ig_create", "ig_destroy", "ig_load_genotypes",
    "ig_load_genotypes_device", "ig_comm_unique_id", "ig_comm_init", "ig_run_chain", "ig_mcmc_updating",
    "ig_chain_init", "ig_sweep", "ig_sync", "ig_time_sweeps", "ig_run_phase", "ig_get_state", "ig_set_state", "ig_loglik",
    "ig_proposal_loglik", "ig_alpha_logratio", "ig_profile", "ig_profile_read", "ig_algorithmic_bytes", "ig_get_rate_trace", "ig_release_cache",
]


class IgConfig(C.Structure):
    _fields_ = [
        ("ploid", C.c_int32), ("popnum", C.c_int32), ("locinum", C.c_int32), ("totalsize", C.c_int32),
        ("mode", C.c_int32), ("prior_flag", C.c_int32), ("back_refl", C.c_int32), ("type_freq", C.c_int32),
        ("alpha_dpm", C.c_double),
        ("nstep_check_empty_cluster", C.c_int32), ("print_iter", C.c_int32), ("print_freq", C.c_int32),
        ("autopoly", C.c_int32),
        ("update", C.c_int64), ("burnin", C.c_int64), ("thinning", C.c_int32), ("ckrep", C.c_int32),
        ("seed", C.c_uint64),
        ("device", C.c_int32), ("shard_begin", C.c_int32), ("shard_size", C.c_int32), ("shard_rank", C.c_int32),
        ("shard_count", C.c_int32), ("rng_rounds", C.c_int32), ("use_graph", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


class IgChainResult(C.Structure):
    _fields_ = [
        ("steps", C.c_int64), ("step", C.c_int64), ("flag_empty_cluster", C.c_int32), ("pad", C.c_int32),
        ("totallkh", C.c_double), ("totallkh2", C.c_double),
        ("indvlkh", C.c_void_p), ("qq", C.c_void_p), ("qq2", C.c_void_p), ("self_rates", C.c_void_p),
        ("self_rates2", C.c_void_p), ("gen", C.c_void_p), ("gen2", C.c_void_p), ("freq", C.c_void_p),
        ("freq2", C.c_void_p),
    ]


class InstructError(RuntimeError):
    pass


_lib = None


def load():
    """Load libinstruct_b200.so (raises if it has not been built -- no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise InstructError(
            f"{LIB_PATH} is missing: build it with `make -C instruct_b200/csrc` "
            "(or __graft_entry__.build()); instruct_b200 has no CPU or PyTorch fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32 = C.c_void_p, C.c_int32
    L.ig_version.restype = C.c_char_p
    L.ig_last_error.restype = C.c_char_p
    L.ig_device_count.restype = C.c_int
    L.ig_create.argtypes = [C.POINTER(IgConfig), C.POINTER(vp)]
    L.ig_destroy.argtypes = [vp]
    L.ig_destroy.restype = None
    L.ig_load_genotypes.argtypes = [vp, vp, vp]
    L.ig_load_genotypes_device.argtypes = [vp, vp, vp]
    L.ig_comm_unique_id.argtypes = [vp]
    L.ig_comm_init.argtypes = [vp, vp]
    L.ig_run_chain.argtypes = [vp, i32, vp, C.POINTER(IgChainResult), vp]
    L.ig_mcmc_updating.argtypes = [C.POINTER(IgConfig), vp, vp, i32, vp, C.POINTER(IgChainResult), vp]
    L.ig_chain_init.argtypes = [vp, i32, vp]
    L.ig_sweep.argtypes = [vp, i32]
    L.ig_sync.argtypes = [vp]
    L.ig_time_sweeps.argtypes = [vp, i32, vp]
    L.ig_run_phase.argtypes = [vp, i32]
    L.ig_get_state.argtypes = [vp, i32, vp, C.c_size_t]
    L.ig_set_state.argtypes = [vp, i32, vp, C.c_size_t]
    L.ig_loglik.argtypes = [vp, vp, vp]
    L.ig_proposal_loglik.argtypes = [vp, vp, vp]
    L.ig_alpha_logratio.argtypes = [vp, C.c_double, vp]
    L.ig_profile.argtypes = [vp, i32]
    L.ig_profile_read.argtypes = [vp, vp, vp, vp]
    L.ig_algorithmic_bytes.argtypes = [vp, vp, vp]
    L.ig_get_rate_trace.argtypes = [vp, vp, C.c_size_t, vp]
    _lib = L
    return L


def check(status, allow=(IG_OK,)):
    if status not in allow:
        raise InstructError(f"instruct_b200 status {status}: {load().ig_last_error().decode()}")
    return status
