// CPU emulation of the AlignFrames warp stage (TEST INFRASTRUCTURE ONLY): the product's pixel functors
// (shinestacker_b200/csrc/ss_al_kernels.cuh) and launch sequence (ss_al_engine.h) compiled with g++ against a
// host backend that runs every "kernel" as plain loops.
// Build: g++ -O2 -ffp-contract=off -DSS_HOST_EMU -shared -fPIC
#define SS_HOST_EMU 1
#include <stdlib.h>
#include "../../shinestacker_b200/csrc/ss_al_engine.h"
#include "emu_backend.h"

typedef AlEngine<DmHostBackend> Emu;

extern "C" {
void *emu_al_create(int w, int h, int bps)
{
    Emu *e = new Emu();
    if (e->create(w, h, bps)) { e->destroy(); delete e; return nullptr; }
    return e;
}
void emu_al_destroy(void *p) { Emu *e = (Emu *)p; e->destroy(); delete e; }
int emu_al_warp(void *p, const void *src, size_t spitch, const double *m, int transform, int border_mode,
                const double *bval, const int *kernel, int taps, void *dst, size_t dpitch)
{
    return ((Emu *)p)->warp(src, spitch, 0, m, transform, border_mode, bval, kernel, taps, dst, dpitch);
}
int emu_al_ecc_set_reference(void *p, const void *ref, size_t pitch) { return ((Emu *)p)->ecc_set_reference(ref, pitch, 0); }
int emu_al_ecc_estimate(void *p, const void *mov, size_t pitch, int finest, int max_iters, double eps, const double *init_m,
                        double *m_out, double *info)
{
    return ((Emu *)p)->ecc_estimate(mov, pitch, 0, finest, max_iters, eps, init_m, m_out, info);
}
const char *emu_al_error(void *p) { return ((Emu *)p)->err; }
}
