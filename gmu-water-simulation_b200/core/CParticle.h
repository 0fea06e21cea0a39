// CParticle.h — particle record and physical constants shared with the device code.
//
// CParticle::Physics is the reference's 80-byte, 16-aligned AoS record (include/CParticle.h:19-43,
// resources/kernels/sph_common.cl:29-39); it is the layout of the host mirror m_clParticles and of
// sph_particle at the C ABI.  The Qt3D sphere entity the reference wraps around each record
// (src/CParticle.cpp) belongs to the optional viewer and is not part of the headless core.
#pragma once

#include <cstdint>

typedef float cl_float;
typedef int32_t cl_int;
typedef uint32_t cl_uint;
struct alignas(16) cl_float3 { float x, y, z, w; };
struct alignas(16) cl_int3 { int32_t x, y, z, w; };

class CParticle {
public:
    struct alignas(16) Physics {
        cl_float3 position;
        cl_float3 velocity;
        cl_float3 acceleration;
        cl_int3 grid_position;
        cl_float density;
        cl_float pressure;
        cl_uint id;
        cl_uint cell_id;

        Physics(float x, float y, float z, cl_uint particleId, cl_float3 initialVelocity = {0, 0, 0, 0})
            : position{x, y, z, 0.f}, velocity(initialVelocity), acceleration{0.f, 0.f, 0.f, 0.f},
              grid_position{0, 0, 0, 0}, density(0.f), pressure(0.f), id(particleId), cell_id(0) {}
    };

    // include/CParticle.h:80-84 — must equal the constants the kernels use
    static constexpr float h = 0.0457f;
    static constexpr float viscosity = 3.5f;
    static constexpr float mass = 0.02f;
    static constexpr float gas_stiffness = 3.0f;
    static constexpr float rest_density = 998.29f;
};

static_assert(sizeof(CParticle::Physics) == 80, "CParticle::Physics must stay 80 bytes");
