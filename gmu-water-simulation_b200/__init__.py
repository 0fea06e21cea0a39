"""gmu-water-simulation_b200 — B200-native SPH time step behind the reference's simulator interface.

Python side: ctypes bindings of the two shared libraries (no compute happens in Python):

* ``libsph_cuda.so``  — the C ABI of ``include/sph_cuda.h`` (hand-written sm_100a kernels);
* ``libsph_host.so``  — the headless C++ simulator core (``CBaseParticleSimulator`` /
  ``CCUDAParticleSimulator``) behind the facade of ``include/sph_host.h``.

There is no CPU fallback: if the libraries are missing the import of :mod:`binding` symbols raises.
"""
from .binding import (  # noqa: F401
    DAM_BREAK,
    FOUNTAIN,
    PARTICLE_DTYPE,
    Simulator,
    SphContext,
    SphError,
    build,
    comm_local_id,
    comm_unique_id,
    simulation_type,
    slab_face_shift,
    slab_plan,
    cuda_lib,
    declared_symbols,
    device_count,
    device_name,
    host_lib,
    kernel_source_hash,
    make_config,
    particles_from_arrays,
)
