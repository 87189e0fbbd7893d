// Internal declarations shared by the host-side sources of libfidget_cuda (include/fidget_cuda.h):
// error reporting, device buffers, the context / tape / evaluator objects and the helpers that more
// than one translation unit uses.
#pragma once
#include <algorithm>
#include <atomic>
#include <memory>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/fidget_cuda.h"
#include "kernels.cuh"
#include "effects.cuh"

using namespace fdev;

extern thread_local std::string g_err;   // defined in capi.cu
inline int32_t fail(int32_t code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(FC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() { return static_cast<T*>(p); }
};

inline bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Pinned (page-locked, mapped) host memory can be written by kernels directly over PCIe.
// Measured on B200 (profiles/r01_prospero4096.md): SM stores over PCIe reach well under half the
// bandwidth of a DMA copy (3.07 ms vs 2.07 ms end to end for a 67 MB image), so this is opt-in
// (FIDGET_B200_ZEROCOPY=1); the default stages the image in HBM and copies it with the DMA engine.
// Returns the device alias of `p` or null.
inline void* pinned_device_alias(const void* p) {
    if (!p) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// Tuning knobs from the environment.  FIDGET_B200_ENV_LIVE=1 re-reads them on every call (tests flip knobs
// between renders); otherwise each (name) is read once per process -- no getenv on the render path.
inline int env_int(const char* name, int dflt) {
    struct Slot { const char* name; int value; bool set; };
    static Slot cache[32];
    static std::mutex mu;
    static const bool live = [] { const char* v = getenv("FIDGET_B200_ENV_LIVE"); return v && *v && atoi(v) != 0; }();
    auto read = [&](int d) { const char* v = getenv(name); return v && *v ? atoi(v) : d; };
    if (live) return read(dflt);
    std::lock_guard<std::mutex> g(mu);
    for (auto& sl : cache) {
        if (sl.name == name) return sl.set ? sl.value : dflt;
        if (!sl.name) {
            const char* v = getenv(name);
            sl.name = name;
            sl.set = v && *v;
            sl.value = sl.set ? atoi(v) : 0;
            return sl.set ? sl.value : dflt;
        }
    }
    return read(dflt);
}

struct fc_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr;        // fills are painted here, concurrently with the next levels
    cudaEvent_t ev_fork[MAX_LEVELS] = {}, ev_join = nullptr;
    uint64_t arena_bytes = 1ull << 30;
    uint32_t epoch = 0;                       // ready mark of the current render's job / fill records
    // render scratch
    DevBuf arena, jobs[MAX_LEVELS + 1], fills[MAX_LEVELS], choice_scratch, counters, stats, image, heightmap, leaf_tapes, zsort, census, occl;
    DevBuf mesh_leaves, mesh_scratch, mesh_verts, mesh_tris;   // fc_mesh_build: sampler output and the mesh, resident in HBM
    uint32_t mesh_n_verts = 0, mesh_n_tris = 0;
    DevBuf fx_in, fx_out, fx_tmp, fx_tables;  // effects: staged host images, intermediate maps, SSAO tables
    // tile interleave: device list of this rank's XY root tiles (cached on its key), and the
    // tile -> gathered-slot table of fc_tiles_unpack
    DevBuf root_list, tile_slots;
    uint32_t root_list_key[5] = {0, 0, 0, 0, 0}, root_list_n = 0;
    uint32_t tile_slots_key[4] = {0, 0, 0, 0};
    std::vector<cudaEvent_t> events;
    std::mutex mu;
    // tape uploads: released device buffers are reused (no cudaMalloc / cudaFree per tape) and the
    // clauses go through a pinned staging buffer with a stream-ordered copy (no host synchronisation)
    std::vector<std::pair<size_t, uint2*>> tape_pool;
    void* stage = nullptr;
    size_t stage_cap = 0;
    cudaEvent_t stage_ev = nullptr;
    struct { size_t smem; int per_sm, threads; } coop_memo[2] = {};   // level-0 launch shape per DIM (occupancy query cached)
    std::shared_ptr<struct Sched> sched_cache[4];
    unsigned sched_next = 0;
};

struct fc_tape {
    fc_ctx* ctx = nullptr;
    std::atomic<int> refs{1};
    uint2* dev = nullptr;
    size_t dev_cap = 0;       // bytes behind `dev` (a pooled buffer may be larger than the tape)
    bool pooled_ok = true;    // false for tapes whose buffer is not a plain cudaMalloc of their own
    std::vector<uint2> host;  // copy of the device clauses
    fc_tape_info info{};
    int ax[3] = {-1, -1, -1};  // input slots of X, Y, Z
    // cooperative level-0 schedule (null when the tape is unsuitable); shared between tapes
    // created from identical bytecode (re-uploading an unchanged shape every frame is the
    // common interactive pattern)
    std::shared_ptr<struct Sched> sched;
};

struct Sched {
    int device = 0;
    uint64_t hash = 0;
    std::vector<uint2> clauses;
    CoopRec* d_recs = nullptr;
    CoopFwd* d_fwd = nullptr;
    uint32_t* d_wave_start = nullptr;
    uint32_t n_waves = 0, tail_begin = 0, tail_end = 0, n_slots = 0;
    std::vector<CoopSeg> segs;
    ~Sched() {
        cudaSetDevice(device);
        if (d_recs) cudaFree(d_recs);
        if (d_fwd) cudaFree(d_fwd);
        if (d_wave_start) cudaFree(d_wave_start);
    }
};
struct fc_eval {
    fc_ctx* ctx = nullptr;
    DevBuf in, out, choices, simplify, ptrs, tmp;
};

// schedule.cu
void upload_schedule(fc_tape* t);
int coop_blocks(fc_ctx* c, const fc_tape* tape, uint64_t n_roots, LevelParams& p, int dim, int& threads);
// capi.cu
int32_t check_device_errors(fc_ctx* c);
int32_t transcode(const uint32_t* words, size_t n_words, uint8_t reg_count, uint32_t mem_count, uint32_t n_vars,
                  uint32_t n_outputs, std::vector<uint2>& out, uint32_t& n_choices);
// render.cu
int32_t pick_tile_sizes(const uint32_t* ts_in, uint32_t n_in, const uint32_t* dflt, uint32_t n_dflt, uint32_t max_size,
                        std::vector<uint32_t>& ts);
int32_t bind_vars(const fc_tape* t, const float* values, uint32_t n_values, VarBind& vb);
cudaEvent_t get_event(fc_ctx* c, size_t i);
int32_t root_subset(fc_ctx* c, uint32_t roots_x, uint32_t row0, uint32_t row1, uint32_t stride, uint32_t offset,
                    cudaStream_t s, const uint32_t** d_list, uint32_t* n);
