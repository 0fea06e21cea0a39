/* minimal_step.c — the C ABI of include/sph_cuda.h from plain C: a 16 000-particle dam break, 100 steps, read-back.
 *
 *   gcc -std=c99 -Iinclude examples/minimal_step.c -Lgmu-water-simulation_b200 -lsph_cuda \
 *       -Wl,-rpath,$PWD/gmu-water-simulation_b200 -lm -o minimal_step && ./minimal_step
 *
 * Without a CUDA device sph_create fails with SPH_ERR_CUDA (there is no CPU fallback) and the program says so. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sph_cuda.h"

#define CHECK(call)                                                                 \
    do {                                                                            \
        int rc_ = (call);                                                           \
        if (rc_ != SPH_OK) {                                                        \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, sph_last_error(ctx)); \
            sph_destroy(ctx);                                                       \
            return 1;                                                               \
        }                                                                           \
    } while (0)

int main(void) {
    sph_context *ctx = NULL;
    const float box = 0.9f, h = 0.0457f, half = h / 2.0f;

    /* the reference's dam-break lattice (src/CBaseParticleSimulator.cpp:44-55): fp32 loop accumulators */
    uint32_t cap = 20000, n = 0;
    sph_particle *p = (sph_particle *)calloc(cap, sizeof *p);
    for (float y = 0; y < box; y += half)
        for (float x = 0; x < box / 4.0; x += half)
            for (float z = 0; z < box; z += half) {
                if (n == cap) return 2;
                p[n].position[0] = x - box / 2.0f;
                p[n].position[1] = y - box / 2.0f;
                p[n].position[2] = z - box / 2.0f;
                p[n].id = n;
                ++n;
            }

    sph_config cfg;
    sph_config_init(&cfg, box, box, box, n);
    CHECK(sph_create(&cfg, &ctx));
    CHECK(sph_upload_particles(ctx, p, n));
    double ms = 0.0;
    CHECK(sph_step(ctx, 100, &ms));
    uint32_t got = 0;
    CHECK(sph_download_particles(ctx, p, cap, &got));
    double ymax = -1e30;
    for (uint32_t i = 0; i < got; ++i) ymax = fmax(ymax, p[i].position[1]);
    printf("%u particles, 100 steps in %.3f ms on the device (%.3g particle-steps/s), fill height %.3f m\n", got, ms,
           100.0 * got / (ms * 1e-3), ymax + box / 2.0);
    sph_destroy(ctx);
    free(p);
    return 0;
}
