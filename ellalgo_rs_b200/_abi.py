"""ctypes mirror of include/ellcuda.h (structs and constants only, no library loading).

Kept separate so that test infrastructure can describe its inputs with the very same records the
product takes without loading the library.
"""
import ctypes as C

# error codes
OK, E_ARG, E_CUDA, E_COMM, E_OOM = 0, -1, -2, -3, -4

# CutStatus (src/cutting_plane.rs:32-37)
SUCCESS, NOSOLN, NOEFFECT, UNKNOWN = 0, 1, 2, 3
# exit status of a cutting_plane_* run
EXIT_NOSOLN, EXIT_NOEFFECT, EXIT_TOL, EXIT_MAXIT, EXIT_FEASIBLE = 1, 2, 4, 5, 6
EXIT_NAMES = {1: "NoSoln", 2: "NoEffect", 4: "Tol", 5: "MaxIt", 6: "Feasible"}

KIND_ELL, KIND_ELLSTABLE = 0, 1
STORAGE_FULL, STORAGE_SYMMETRIC = 0, 1
UPDATE_BIAS, UPDATE_CENTRAL, UPDATE_Q = 0, 1, 2
MODE_OPTIM, MODE_FEAS, MODE_OPTIM_Q, MODE_BSEARCH = 0, 1, 2, 3

(ORACLE_EXAMPLE1, ORACLE_EXAMPLE1_RR, ORACLE_EXAMPLE4, ORACLE_QUASICVX, ORACLE_PROFIT,
 ORACLE_PROFIT_RB, ORACLE_PROFIT_Q, ORACLE_LOWPASS, ORACLE_BOWL, ORACLE_LMI, ORACLE_EXAMPLE2,
 ORACLE_EXAMPLE3, ORACLE_HALFPLANE, ORACLE_ALWAYS_CUT, ORACLE_OPTIM_BOX, ORACLE_LMI_LINOPT, ORACLE_LMI_OLD,
 ORACLE_LMI0) = range(18)

_pd = C.POINTER(C.c_double)


class BatchDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("oracle_id", C.c_int32), ("mode", C.c_int32), ("n", C.c_int32),
        ("count", C.c_int64), ("max_iters", C.c_int64), ("tolerance", C.c_double),
        ("no_defer_trick", C.c_int32), ("use_parallel_cut", C.c_int32),
        ("kappa0", C.c_void_p), ("kappa0_stride", C.c_int64),
        ("xc0", C.c_void_p), ("xc0_stride", C.c_int64),
        ("diag0", C.c_void_p), ("diag0_stride", C.c_int64),
        ("mq0", C.c_void_p), ("mq0_stride", C.c_int64),
        ("gamma0", C.c_void_p), ("gamma0_stride", C.c_int64),
        ("params", C.c_void_p), ("params_stride", C.c_int64),
        ("params_width", C.c_int32), ("_pad0", C.c_int32),
        ("table", C.c_void_p), ("table_len", C.c_int64),
        ("aux", C.c_int32 * 8),
        ("device", C.c_int32), ("on_device", C.c_int32),
        ("stream", C.c_void_p),
    ]


class BatchResult(C.Structure):
    _fields_ = [
        ("xbest", C.c_void_p), ("has_xbest", C.c_void_p), ("num_iters", C.c_void_p),
        ("exit_status", C.c_void_p), ("gamma", C.c_void_p), ("tsq", C.c_void_p),
        ("xc", C.c_void_p), ("kappa", C.c_void_p),
    ]


class CreateDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("device", C.c_int32), ("n", C.c_int64), ("kappa", C.c_double),
        ("mq", C.c_void_p), ("diag", C.c_void_p), ("xc", C.c_void_p),
        ("no_defer_trick", C.c_int32), ("use_parallel_cut", C.c_int32),
        ("row_begin", C.c_int64), ("row_end", C.c_int64),
        ("reduce_mode", C.c_int32), ("storage", C.c_int32),
        ("stream", C.c_void_p),
        ("shard_rank", C.c_int32), ("shard_count", C.c_int32),
    ]
