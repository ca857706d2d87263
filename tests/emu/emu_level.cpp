// CPU emulation of the fused level kernel's tile code (TEST INFRASTRUCTURE ONLY): compiles
// shinestacker_b200/csrc/ss_kernels.cuh with g++ and runs the tiles sequentially, so the halo /
// border / tail-column index arithmetic can be checked against the oracle without a GPU.
// Build: g++ -O1 -mfma -ffp-contract=off -DSS_HOST_EMU -shared -fPIC
#define SS_HOST_EMU 1
#include "../../shinestacker_b200/csrc/ss_kernels.cuh"
#include <stdlib.h>
#include <string.h>

extern "C" int emu_level(const void *in, int bytes_per_sample /*1,2: interleaved; 4: planar f32*/,
                         int in_pitch_bytes, int n_frames, int h, int w, const float *k25, float *gnext,
                         int gnext_pitch, float *best_e, int32_t *best_idx, float *fused, int frame_index0,
                         int init)
{
    SsLevelParams p;
    memset(&p, 0, sizeof(p));
    p.h = h; p.w = w; p.h2 = (h + 1) / 2; p.w2 = (w + 1) / 2;
    p.in = in;
    p.in_pitch = in_pitch_bytes;
    if (bytes_per_sample == 4) {
        p.in_plane_stride = (long long)h * in_pitch_bytes; p.in_frame_stride = 3 * p.in_plane_stride;
    } else {
        p.in_frame_stride = (long long)h * p.in_pitch;
    }
    p.gnext = gnext; p.gnext_pitch = gnext_pitch; p.gnext_plane_stride = (long long)p.h2 * gnext_pitch;
    p.gnext_frame_stride = 3 * p.gnext_plane_stride;
    p.best_e = best_e; p.best_idx = best_idx; p.fused = fused;
    p.n_frames = n_frames; p.frame_index0 = frame_index0; p.init = init; p.groups = 1; p.group_stride = 0;
    memcpy(p.k25, k25, sizeof(p.k25));
    p.skip_factor = getenv("SS_EMU_NO_SKIP") ? 0.f : ss_skip_factor(k25);
    float *smem = (float *)malloc(SS_SMEM_BYTES);
    float *scratch = (float *)malloc(sizeof(float) * 3 * SS_TX * SS_TY);
    p.lap_scratch = scratch;
    // poison shared memory so that any read of a never-written slot shows up as NaN
    for (int ty = 0; ty < (h + SS_TY - 1) / SS_TY; ++ty)
        for (int tx = 0; tx < (w + SS_TX - 1) / SS_TX; ++tx) {
            memset(smem, 0xff, SS_SMEM_BYTES);
            if (bytes_per_sample == 1) ss_level_tile<uint8_t>(p, nullptr, tx, ty, 0, 0, smem);
            else if (bytes_per_sample == 2) ss_level_tile<uint16_t>(p, nullptr, tx, ty, 0, 0, smem);
            else ss_level_tile<float>(p, nullptr, tx, ty, 0, 0, smem);
        }
    free(smem);
    free(scratch);
    return 0;
}

extern "C" long long emu_skipped_threads(int reset)
{
    const long long v = ss_emu_skipped_threads;
    if (reset) ss_emu_skipped_threads = 0;
    return v;
}
