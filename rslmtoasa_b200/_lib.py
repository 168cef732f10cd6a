"""ctypes binding of include/rsrec.h (the same symbols a Fortran ISO_C_BINDING module binds)."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_lib = None

SYMBOLS = [
    "rsrec_last_error", "rsrec_version", "rsrec_compiled_arch", "rsrec_device_count", "rsrec_create", "rsrec_destroy", "rsrec_set_lattice",
    "rsrec_set_hamiltonian", "rsrec_set_operator", "rsrec_lanczos_block", "rsrec_lanczos_scalar", "rsrec_zsqr",
    "rsrec_cheb_moments", "rsrec_cheb_moments_random", "rsrec_kubo_moments", "rsrec_ham_vec_matmul",
    "rsrec_velo_vec_matmul", "rsrec_cheb_begin_random", "rsrec_cheb_begin_sites", "rsrec_cheb_run_steps",
    "rsrec_cheb_end", "rsrec_synchronize", "rsrec_stream", "rsrec_launch_count", "rsrec_set_kernel_family",
    "rsrec_h2d_bytes", "rsrec_d2h_bytes", "rsrec_profile", "rsrec_profile_read", "rsrec_set_fusion",
    "rsrec_bpopt", "rsrec_get_terminf", "rsrec_bgreen", "rsrec_block_green", "rsrec_chebyshev_green", "rsrec_density",
    "rsrec_sgreen", "rsrec_conductivity_integrand", "rsrec_recur_b_green", "rsrec_cheb_recur_green",
    "rsrec_kubo_conductivity", "rsrec_create_ll_map", "rsrec_orbital_moments",
    "rsrec_build_nn", "rsrec_build_hamiltonian", "rsrec_rotate_to_local_axis", "rsrec_rotate_from_local_axis",
    "rsrec_lanczos_block_local_axis", "rsrec_set_positions",
    "rsrec_bands_set_g0", "rsrec_bands_get_g0", "rsrec_bands_g0_shape", "rsrec_bands_dos", "rsrec_bands_fermi",
    "rsrec_bands_magnetic_moments", "rsrec_bands_moments", "rsrec_bands_band_energy",
    "rsrec_recur_b_ij_green", "rsrec_cheb_recur_ij_green", "rsrec_intersite_gf", "rsrec_conductivity_cumulative", "rsrec_spin_diag_launch_count",
    "rsrec_phase_timing", "rsrec_phase_count", "rsrec_phase_label", "rsrec_phase_read", "rsrec_host_phase_read",
    "rsrec_comm_unique_id", "rsrec_comm_init", "rsrec_comm_destroy", "rsrec_comm_info", "rsrec_shard_range",
    "rsrec_allreduce", "rsrec_allgather_units", "rsrec_lanczos_block_sharded", "rsrec_cheb_moments_random_sum",
]


class RsrecError(RuntimeError):
    """Raised where the Fortran host would call g_logger%fatal (reference logger.f90:186-193)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"rsrec error {code}: {msg}")
        self.code = code


def _prefer_bundled_nccl():
    """librsrec.so opens libnccl.so.2 with dlopen on the first rsrec_comm_* call.  In a Python process that ALSO imports torch
    the library with that SONAME must be the one torch was built against (the nvidia-nccl wheel): if the older system NCCL
    is loaded first, `import torch` later resolves its NCCL symbols against it and fails (undefined ncclDevCommCreate).
    RSREC_NCCL_LIB (honoured by csrc/comm_nccl.cuh before the SONAME search) is pointed at the wheel's file, found without
    importing torch.  A Fortran / C++ host has no such wheel and uses the system library."""
    if os.environ.get("RSREC_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["RSREC_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def load():
    """Load librsrec.so (never builds implicitly on a GPU box: the .so travels in-tree).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB):
        raise RuntimeError(f"{_build.LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    _prefer_bundled_nccl()
    L = C.CDLL(_build.LIB)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    L.rsrec_last_error.restype = C.c_char_p
    L.rsrec_create.argtypes = [C.POINTER(vp), i, i, i, i, i, i]
    L.rsrec_destroy.argtypes = [vp]
    L.rsrec_set_lattice.argtypes = [vp, vp, vp]
    L.rsrec_set_hamiltonian.argtypes = [vp, vp, vp, vp, vp, vp, vp, i]
    L.rsrec_set_operator.argtypes = [vp, i, vp, vp]
    L.rsrec_lanczos_block.argtypes = [vp, i, vp, vp, vp, vp, i, vp, vp]
    L.rsrec_lanczos_scalar.argtypes = [vp, i, vp, i, vp, vp]
    L.rsrec_zsqr.argtypes = [vp, vp, i, i]
    L.rsrec_cheb_moments.argtypes = [vp, i, vp, vp, vp, vp, i, d, d, vp]
    L.rsrec_cheb_moments_random.argtypes = [vp, i, vp, i, d, d, vp]
    L.rsrec_kubo_moments.argtypes = [vp, i, i, vp, vp, i, d, d, vp]
    L.rsrec_ham_vec_matmul.argtypes = [vp, vp, vp, d, d]
    L.rsrec_velo_vec_matmul.argtypes = [vp, i, vp, vp]
    L.rsrec_cheb_begin_random.argtypes = [vp, i, vp, i, d, d]
    L.rsrec_cheb_begin_sites.argtypes = [vp, i, vp, vp, vp, vp, i, d, d]
    L.rsrec_cheb_run_steps.argtypes = [vp, i]
    L.rsrec_cheb_end.argtypes = [vp, vp]
    L.rsrec_synchronize.argtypes = [vp]
    L.rsrec_stream.argtypes = [vp]
    L.rsrec_stream.restype = vp
    L.rsrec_launch_count.argtypes = [vp]
    L.rsrec_launch_count.restype = C.c_longlong
    L.rsrec_spin_diag_launch_count.argtypes = [vp]
    L.rsrec_spin_diag_launch_count.restype = C.c_longlong
    L.rsrec_set_kernel_family.argtypes = [vp, i]
    L.rsrec_set_fusion.argtypes = [vp, i, i]
    L.rsrec_h2d_bytes.argtypes = [vp]
    L.rsrec_h2d_bytes.restype = C.c_longlong
    L.rsrec_d2h_bytes.argtypes = [vp]
    L.rsrec_d2h_bytes.restype = C.c_longlong
    L.rsrec_profile.argtypes = [vp, i]
    L.rsrec_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i)]
    L.rsrec_bpopt.argtypes = [vp, i, i, vp, vp, vp, vp, vp]
    L.rsrec_get_terminf.argtypes = [vp, vp, vp, i, i, vp, vp, vp, vp]
    L.rsrec_bgreen.argtypes = [vp, vp, vp, i, vp, i, i, i, vp, vp, d, d, i, vp]
    L.rsrec_block_green.argtypes = [vp, vp, vp, i, i, vp, i, i, vp]
    L.rsrec_chebyshev_green.argtypes = [vp, vp, i, i, vp, i, d, d, vp, vp]
    L.rsrec_density.argtypes = [vp, vp, vp, i, i, i, vp, i, vp, vp, vp]
    L.rsrec_sgreen.argtypes = [vp, vp, vp, i, i, i, vp, i, vp, vp, vp]
    L.rsrec_conductivity_integrand.argtypes = [vp, vp, i, i, vp, i, d, d, i, vp, vp]
    L.rsrec_recur_b_green.argtypes = [vp, i, vp, i, vp, i, i, vp, vp, vp]
    L.rsrec_cheb_recur_green.argtypes = [vp, i, vp, i, d, d, vp, i, vp, vp, vp]
    L.rsrec_kubo_conductivity.argtypes = [vp, i, i, vp, vp, i, d, d, vp, i, vp, vp, vp]
    L.rsrec_create_ll_map.argtypes = [vp, i, i, vp]
    L.rsrec_orbital_moments.argtypes = [vp, i, vp, vp, d, i, d, d, vp]
    L.rsrec_build_nn.argtypes = [i, i, vp, vp, i, vp, d, vp, vp, vp, d, i, vp, C.POINTER(i)]
    L.rsrec_build_hamiltonian.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp]
    L.rsrec_rotate_to_local_axis.argtypes = [vp, vp]
    L.rsrec_rotate_from_local_axis.argtypes = [vp]
    L.rsrec_lanczos_block_local_axis.argtypes = [vp, i, vp, vp, i, vp, vp]
    L.rsrec_recur_b_ij_green.argtypes = [vp, i, vp, vp, vp, vp, i, vp, i, i, vp, vp, vp]
    L.rsrec_cheb_recur_ij_green.argtypes = [vp, i, vp, vp, vp, vp, i, d, d, vp, i, vp, vp, vp]
    L.rsrec_intersite_gf.argtypes = [vp, i, vp, vp, i, vp, vp, vp]
    L.rsrec_conductivity_cumulative.argtypes = [vp, vp, vp, i, i, i, d, i, vp]
    L.rsrec_bands_set_g0.argtypes = [vp, vp, i, i]
    L.rsrec_bands_get_g0.argtypes = [vp, vp]
    L.rsrec_bands_g0_shape.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
    L.rsrec_bands_dos.argtypes = [vp, vp, vp, vp]
    L.rsrec_bands_fermi.argtypes = [vp, vp, i, d, d, d, i, C.POINTER(d), C.POINTER(i), C.POINTER(d), C.POINTER(i)]
    L.rsrec_bands_magnetic_moments.argtypes = [vp, vp, d, d, i, d, vp, vp]
    L.rsrec_bands_moments.argtypes = [vp, i, vp, d, d, i, d, vp, vp, vp]
    L.rsrec_bands_band_energy.argtypes = [vp, vp, i, vp, d, d, i, d, vp]
    L.rsrec_set_positions.argtypes = [vp, vp]
    L.rsrec_phase_timing.argtypes = [vp, i]
    L.rsrec_phase_label.argtypes = [i]
    L.rsrec_phase_label.restype = C.c_char_p
    L.rsrec_phase_read.argtypes = [vp, vp, vp]
    L.rsrec_host_phase_read.argtypes = [vp, vp]
    L.rsrec_comm_unique_id.argtypes = [vp]
    L.rsrec_comm_init.argtypes = [vp, i, i, vp]
    L.rsrec_comm_destroy.argtypes = [vp]
    L.rsrec_comm_info.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.rsrec_shard_range.argtypes = [i, i, i, C.POINTER(i), C.POINTER(i)]
    L.rsrec_allreduce.argtypes = [vp, vp, C.c_longlong, i]
    L.rsrec_allgather_units.argtypes = [vp, vp, vp, C.c_longlong, i]
    L.rsrec_lanczos_block_sharded.argtypes = [vp, i, vp, vp, vp, vp, i, vp, vp]
    L.rsrec_cheb_moments_random_sum.argtypes = [vp, i, vp, i, d, d, vp]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise RsrecError(rc, load().rsrec_last_error().decode())
