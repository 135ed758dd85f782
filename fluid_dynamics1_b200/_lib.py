"""ctypes binding of libcnavier_b200.so (include/cnavier_b200.h) and libcnavier_dropin.so.

The libraries are built in-tree by ``fluid_dynamics1_b200/csrc/Makefile`` (see ``__graft_entry__.build``).
There is no CPU fallback: if the CUDA library is missing, loading raises; if no CUDA device is
visible, every compute entry point raises ``RuntimeError`` before touching the library.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libcnavier_b200.so")
DROPIN_PATH = os.path.join(PKG, "libcnavier_dropin.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_vp = C.c_void_p


class Config(C.Structure):
    """Mirror of the reference's Config (include/config.h:4-36), same field order and types."""
    _fields_ = [("Re", C.c_double), ("Lx", C.c_int), ("Ly", C.c_int), ("nx", C.c_int), ("ny", C.c_int),
                ("dt", C.c_double), ("tf", C.c_double), ("max_co", C.c_double), ("order", C.c_int),
                ("poisson_max_it", C.c_int), ("poisson_tol", C.c_double), ("output_interval", C.c_int),
                ("poisson_type", C.c_int), ("openmp_enabled", C.c_int), ("ui", C.c_double), ("vi", C.c_double),
                ("u1", C.c_double), ("u2", C.c_double), ("u3", C.c_double), ("u4", C.c_double),
                ("v1", C.c_double), ("v2", C.c_double), ("v3", C.c_double), ("v4", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Mtrx(C.Structure):
    """The reference's mtrx (include/linearalg.h:6-11), passed and returned by value."""
    _fields_ = [("M", C.POINTER(C.POINTER(C.c_double))), ("m", C.c_int), ("n", C.c_int)]


# name -> (restype, argtypes); every symbol include/cnavier_b200.h declares
CNV_API = {
    "cnv_version": (C.c_char_p, []),
    "cnv_device_count": (C.c_int, []),
    "cnv_launch_count": (C.c_ulonglong, []),
    "cnv_set_device": (C.c_int, [C.c_int]),
    "cnv_get_device": (C.c_int, []),
    "cnv_device_synchronize": (None, []),
    "cnv_host_alloc": (_vp, [C.c_size_t]),
    "cnv_host_free": (None, [_vp]),
    "cnv_host_is_pinned": (C.c_int, [_vp]),
    "cnv_host_numa_node": (C.c_int, []),
    "cnv_sor_beta": (C.c_double, [C.c_int, C.c_int]),
    "cnv_num_steps": (C.c_int, [C.c_double, C.c_double]),
    "cnv_diff_dense": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, _dp]),
    "cnv_apply_host": (C.c_int, [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp]),
    "cnv_euler_host": (C.c_int, [_dp] * 7 + [C.c_int, C.c_int, C.c_double, C.c_double]),
    "cnv_continuity_host": (C.c_int, [_dp, _dp, C.c_int, C.c_int, _dp]),
    "cnv_vorticity_host": (C.c_int, [_dp, _dp, C.c_int, C.c_int, _dp]),
    "cnv_error_host": (C.c_double, [_dp, _dp, C.c_int, C.c_int]),
    "cnv_pressure_rhs_host": (C.c_int, [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _dp]),
    "cnv_poisson_host": (C.c_int, [_dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int,
                                   _dp, C.POINTER(C.c_int), C.POINTER(C.c_double), _vp]),
    "cnv_poisson_host_cache_clear": (None, []),
    "cnv_poisson_create": (_vp, [C.c_int, C.c_int, C.c_int]),
    "cnv_poisson_create_slab": (_vp, [C.c_int] * 7),
    "cnv_poisson_destroy": (None, [_vp]),
    "cnv_poisson_set_consts": (None, [_vp, C.c_double, C.c_double, C.c_double]),
    "cnv_poisson_ld": (C.c_int, [_vp]),
    "cnv_poisson_rhs_ptr": (_vp, [_vp]),
    "cnv_poisson_buf_ptr": (_vp, [_vp, C.c_int]),
    "cnv_poisson_num_buffers": (C.c_int, [_vp]),
    "cnv_poisson_norms_ptr": (_vp, [_vp]),
    "cnv_poisson_plan_info": (None, [_vp, C.POINTER(C.c_longlong)]),
    "cnv_poisson_onchip_profile": (C.c_int, [_vp, _vp, C.c_int]),
    "cnv_poisson_upload": (C.c_int, [_vp, _dp, C.c_double, _vp]),
    "cnv_poisson_upload_owned": (C.c_int, [_vp, _dp, C.c_double, _vp]),
    "cnv_poisson_download_owned_async": (C.c_int, [_vp, C.c_int, _dp, _vp]),
    "cnv_poisson_prepare": (C.c_int, [_vp, _vp, C.c_int, C.c_double, _vp]),
    "cnv_poisson_solve": (C.c_int, [_vp, C.c_int, C.c_double, _vp, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cnv_poisson_enable_history": (None, [_vp, C.c_int]),
    "cnv_poisson_read_history": (C.c_int, [_vp, _dp, C.c_int]),
    "cnv_poisson_reset": (None, [_vp, C.c_int, C.c_double, _vp]),
    "cnv_poisson_enqueue": (None, [_vp, C.c_int, _vp]),
    "cnv_poisson_enqueue_decide": (None, [_vp, _vp]),
    "cnv_poisson_set_distributed": (None, [_vp, C.c_int]),
    "cnv_poisson_state": (None, [_vp, _vp, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "cnv_poisson_download": (C.c_int, [_vp, C.c_int, _dp, _vp]),
    "cnv_poisson_download_async": (C.c_int, [_vp, C.c_int, _dp, _vp]),
    "cnv_comm_unique_id": (C.c_int, [C.c_char_p]),
    "cnv_comm_create": (_vp, [C.c_int, C.c_int, C.c_char_p]),
    "cnv_comm_destroy": (None, [_vp]),
    "cnv_poisson_attach_comm": (None, [_vp, _vp]),
    "cnv_poisson_enqueue_dist": (None, [_vp, C.c_int, _vp]),
    "cnv_poisson_exchange_halos": (None, [_vp, _vp, C.c_int, _vp]),
    "cnv_poisson_peer_export": (None, [_vp, C.c_char_p]),
    "cnv_poisson_peer_push_counts": (None, [_vp, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "cnv_poisson_peer_import": (C.c_int, [_vp, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_int)]),
    "cnv_poisson_peer_disable": (None, [_vp]),
    "cnv_poisson_peer_quiesce": (None, [_vp, _vp]),
    "cnv_poisson_peer_close": (None, [_vp]),
    "cnv_poisson_peer_enabled": (C.c_int, [_vp]),
    "cnv_poisson_peer_trace": (C.c_int, [_vp, C.c_int]),
    "cnv_poisson_peer_trace_read": (None, [_vp, _vp, C.c_longlong]),
    "cnv_sim_create": (_vp, [C.POINTER(Config), C.c_int]),
    "cnv_sim_create_slab": (_vp, [C.POINTER(Config), C.c_int, C.c_int, C.c_int]),
    "cnv_sim_layout": (None, [_vp, C.POINTER(C.c_int)]),
    "cnv_sim_poisson": (_vp, [_vp]),
    "cnv_sim_field_ptr": (_vp, [_vp, C.c_int]),
    "cnv_sim_set_psi_buf": (None, [_vp, C.c_int]),
    "cnv_sim_phase": (None, [_vp, C.c_int, _vp]),
    "cnv_sim_destroy": (None, [_vp]),
    "cnv_sim_step": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp]),
    "cnv_sim_step_slab": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp]),
    "cnv_sim_gather_fields_slab": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "cnv_sim_get_fields": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "cnv_sim_set_fields": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "cnv_sim_set_diagnostics": (None, [_vp, C.c_int]),
    "cnv_sim_stencil_phase": (None, [_vp, C.c_int, _vp]),
    "cnv_sim_counters": (None, [_vp, C.POINTER(C.c_longlong)]),
    "cnv_sim_pressure": (C.c_int, [_vp, C.c_int, C.c_double, _vp, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "cnv_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "cnv_boot_selftest": (C.c_int, [C.c_int]),
    "cnv_vtk_write": (C.c_int, [_dp, C.c_int, C.c_int, C.c_char_p, C.c_char_p]),
    "cnv_config_default": (None, [C.POINTER(Config)]),
    "cnv_config_from_file": (None, [C.c_char_p, C.POINTER(Config)]),
    "cnv_config_print": (None, [C.POINTER(Config)]),
}

# the reference-named symbols of include/cnavier_dropin.h
DROPIN_API = {
    "zerosm": (None, [Mtrx]), "allocm": (C.POINTER(C.POINTER(C.c_double)), [C.c_int, C.c_int]),
    "freem": (C.POINTER(C.POINTER(C.c_double)), [Mtrx]), "initm": (Mtrx, [C.c_int, C.c_int]), "eye": (Mtrx, [C.c_int]),
    "reshape": (Mtrx, [Mtrx, C.c_int, C.c_int]), "kronecker": (Mtrx, [Mtrx, Mtrx]), "mtrxmul": (Mtrx, [Mtrx, Mtrx]),
    "invsig": (None, [Mtrx]), "maxel": (C.c_double, [Mtrx]), "minel": (C.c_double, [Mtrx]), "mtrxcpy": (None, [Mtrx, Mtrx]),
    "set_openmp_config": (None, [C.c_int]),
    "Diff1": (Mtrx, [C.c_int, C.c_int, C.c_double]), "Diff2": (Mtrx, [C.c_int, C.c_int, C.c_double]),
    "euler": (None, [Mtrx] * 7 + [C.c_double, C.c_double]), "continuity": (Mtrx, [Mtrx, Mtrx]),
    "vorticity": (Mtrx, [Mtrx, Mtrx]), "set_fluiddyn_openmp_config": (None, [C.c_int]),
    "error": (C.c_double, [Mtrx, Mtrx]),
    "poisson": (Mtrx, [Mtrx, C.c_double, C.c_double, C.c_int, C.c_double]),
    "poisson_SOR": (Mtrx, [Mtrx, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double]),
    "poisson_log": (Mtrx, [Mtrx, C.c_double, C.c_double, C.c_int, C.c_double, _vp]),
    "poisson_SOR_log": (Mtrx, [Mtrx, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, _vp]),
    "set_poisson_openmp_config": (None, [C.c_int]),
    "load_default_config": (Config, []), "load_config_from_file": (Config, [C.c_char_p]),
    "print_config": (None, [C.POINTER(Config)]), "print_usage": (None, [C.c_char_p]),
    "print_openmp_status": (None, [C.POINTER(Config)]),
}

_libs: dict = {}


def _bind(path, api, mode):
    if not os.path.exists(path):
        raise ImportError(f"{os.path.basename(path)} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(there is no CPU fallback)")
    L = C.CDLL(path, mode=mode)
    for name, (res, args) in api.items():
        fn = getattr(L, name)  # AttributeError here = the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return L


def lib():
    """libcnavier_b200.so, bound."""
    if "core" not in _libs:
        _libs["core"] = _bind(LIB_PATH, CNV_API, C.RTLD_GLOBAL)
    return _libs["core"]


def dropin():
    """libcnavier_dropin.so, bound (loaded RTLD_LOCAL: it exports generic names such as error())."""
    if "dropin" not in _libs:
        lib()
        _libs["dropin"] = _bind(DROPIN_PATH, DROPIN_API, C.RTLD_LOCAL)
    return _libs["dropin"]


def require_gpu():
    if lib().cnv_device_count() < 1:
        raise RuntimeError("no CUDA device visible: the cnavier B200 path has no CPU fallback")
