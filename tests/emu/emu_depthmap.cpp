// CPU emulation of the depth-map stacker (TEST INFRASTRUCTURE ONLY): compiles the product's pixel functors
// (shinestacker_b200/csrc/ss_dm_kernels.cuh) and its launch sequence (ss_dm_engine.h) with g++ against a
// host backend that runs every "kernel" as plain loops, so tap order, border handling, state machine and
// buffer plumbing can be checked against the oracle and the golden vectors on a CPU-only box.
// Build: g++ -O1 -ffp-contract=off -DSS_HOST_EMU -shared -fPIC
#define SS_HOST_EMU 1
#include <stdlib.h>
#include "../../shinestacker_b200/csrc/ss_dm_engine.h"

#include "emu_backend.h"

typedef DmEngine<DmHostBackend> Emu;

extern "C" {
void *emu_dm_create(int w, int h, int bps, int map_type, int energy, int kernel_size, int blur_size, int smooth_size,
                    double temperature, int levels, int max_frames)
{
    Emu *e = new Emu();
    DmOptions o = {w, h, bps, map_type, energy, kernel_size, blur_size, smooth_size, temperature, levels, max_frames};
    if (e->create(o)) { fprintf(stderr, "emu_dm_create: %s\n", e->err); e->destroy(); delete e; return nullptr; }
    return e;
}
void emu_dm_destroy(void *p) { Emu *e = (Emu *)p; e->destroy(); delete e; }
int emu_dm_reset(void *p) { return ((Emu *)p)->reset(); }
int emu_dm_push(void *p, const void *bgr, size_t pitch) { return ((Emu *)p)->push_frame(bgr, pitch, 0); }
int emu_dm_finish_begin(void *p) { return ((Emu *)p)->finish_begin(); }
int emu_dm_blend(void *p, int i) { return ((Emu *)p)->blend_frame(i); }
int emu_dm_finish_end(void *p, void *out, size_t pitch) { return ((Emu *)p)->finish_end(out, pitch); }
int emu_dm_rerun(void *p) { return ((Emu *)p)->rerun_resident(); }
int emu_dm_get_plane(void *p, int what, int frame, float *out) { return ((Emu *)p)->get_plane(what, frame, out); }
int emu_dm_shard_local_max(void *p, float *out) { return ((Emu *)p)->shard_local_max(out); }
int emu_dm_shard_finish_local(void *p, float gmax) { return ((Emu *)p)->shard_finish_local(gmax); }
int emu_dm_shard_count(void *p) { return 2 + ((Emu *)p)->L; }
int emu_dm_shard_ptrs(void *p, void **ptrs) { return ((Emu *)p)->shard_ptrs(ptrs); }
int emu_dm_shard_connect(void *p, void *const *ptrs, int world, int rank, int root) { return ((Emu *)p)->shard_connect(ptrs, world, rank, root); }
int emu_dm_shard_softmax_rel(void *p) { return ((Emu *)p)->shard_softmax_rel(); }
int emu_dm_shard_finish_weights(void *p) { return ((Emu *)p)->shard_finish_weights(); }
int emu_dm_shard_reduce_band(void *p) { return ((Emu *)p)->shard_reduce_band(); }
int emu_dm_shard_finish_end(void *p, void *out, size_t pitch) { return ((Emu *)p)->shard_finish_end(out, pitch); }
const char *emu_dm_error(void *p) { return ((Emu *)p)->err; }
long long emu_dm_launches(void *p) { return ((Emu *)p)->be.launches; }
}
