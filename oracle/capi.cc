// ORACLE -- test infrastructure, not product code.
// C ABI used by tests / bench.py (ctypes) to drive the CPU oracle.
#include <cstring>
#include <stdexcept>
#include <string>

#include "../fidget_b200/csrc/host/host_capi.h"
#include "octree.h"
#include <mutex>
#include "effects.h"
#include "vm.h"

using namespace oracle;

struct fh_tape { fhost::TapeData d; };  // same layout as in host_capi.cc
struct orc_tape { TapeP t; };

static thread_local std::string g_err;
#define ORC_TRY(body)                                                 \
    try { body; return 0; }                                           \
    catch (const std::exception& e) { g_err = e.what(); return -1; }  \
    catch (...) { g_err = "unknown error"; return -1; }

struct orc_stats {
    uint64_t evaluated[8], filled_inside[8], filled_outside[8], ambiguous[8], simplified[8];
    uint64_t pixels;
};
static void copy_stats(const TileStats& s, orc_stats* o) {
    if (!o) return;
    memcpy(o->evaluated, s.evaluated, sizeof o->evaluated);
    memcpy(o->filled_inside, s.filled_inside, sizeof o->filled_inside);
    memcpy(o->filled_outside, s.filled_outside, sizeof o->filled_outside);
    memcpy(o->ambiguous, s.ambiguous, sizeof o->ambiguous);
    memcpy(o->simplified, s.simplified, sizeof o->simplified);
    o->pixels = s.pixels;
}
static Mat4 to_mat(const float* m16) {
    Mat4 m;
    memcpy(m.m, m16, sizeof m.m);
    return m;
}

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

int32_t orc_tape_from_fh(const fh_tape* t, orc_tape** out) {
    ORC_TRY(*out = new orc_tape{std::make_shared<const Tape>(t->d)});
}
void orc_tape_free(orc_tape* t) { delete t; }

int32_t orc_tape_info(const orc_tape* t, uint32_t* asm_len, uint32_t* ssa_len, uint32_t* choice_count,
                      uint32_t* slot_count, uint32_t* n_vars) {
    ORC_TRY({
        *asm_len = uint32_t(t->t->size());
        *ssa_len = uint32_t(t->t->d.ssa.tape.size());
        *choice_count = t->t->choice_count();
        *slot_count = t->t->slot_count();
        *n_vars = uint32_t(t->t->n_vars());
    });
}

// vars: [n_vars][2] lower/upper; out: [n_out][2]; choices: [choice_count]
int32_t orc_interval_eval(const orc_tape* t, const float* vars, float* out, uint8_t* choices,
                          uint8_t* simplify) {
    ORC_TRY({
        IntervalEval e;
        std::vector<Interval> v(t->t->n_vars());
        for (size_t i = 0; i < v.size(); ++i) v[i] = Interval(vars[2 * i], vars[2 * i + 1]);
        std::vector<Interval> o(t->t->d.ssa.output_count);
        bool s = e.eval(*t->t, v.data(), o.data());
        for (size_t i = 0; i < o.size(); ++i) { out[2 * i] = o[i].lo; out[2 * i + 1] = o[i].hi; }
        if (choices) memcpy(choices, e.choices.data(), e.choices.size());
        if (simplify) *simplify = s;
    });
}
int32_t orc_point_eval(const orc_tape* t, const float* vars, float* out, uint8_t* choices, uint8_t* simplify) {
    ORC_TRY({
        PointEval e;
        bool s = e.eval(*t->t, vars, out);
        if (choices) memcpy(choices, e.choices.data(), e.choices.size());
        if (simplify) *simplify = s;
    });
}
int32_t orc_float_slice_eval(const orc_tape* t, const float* const* vars, float* const* out, uint64_t n) {
    ORC_TRY({
        FloatSliceEval e;
        e.eval(*t->t, vars, size_t(n), out);
    });
}
// grads are {v,dx,dy,dz} quadruples
int32_t orc_grad_slice_eval(const orc_tape* t, const float* const* vars, float* const* out, uint64_t n) {
    ORC_TRY({
        GradSliceEval e;
        e.eval(*t->t, reinterpret_cast<const Grad* const*>(vars), size_t(n), reinterpret_cast<Grad* const*>(out));
    });
}
int32_t orc_simplify(const orc_tape* t, const uint8_t* choices, size_t n, orc_tape** out) {
    ORC_TRY(*out = new orc_tape{simplify(*t->t, choices, n, t->t->d.n_regs)});
}
// VmData::simplify::<M> with a different register count (vm/data.rs:411-437)
int32_t orc_simplify_n(const orc_tape* t, const uint8_t* choices, size_t n, uint32_t n_regs, orc_tape** out) {
    ORC_TRY(*out = new orc_tape{simplify(*t->t, choices, n, n_regs)});
}
// Bytecode of an oracle tape (e.g. a simplified child), for feeding the GPU path
int32_t orc_tape_bytecode(const orc_tape* t, int32_t repack, uint32_t* words, size_t cap, size_t* n_words,
                          uint8_t* reg_count, uint32_t* mem_count) {
    ORC_TRY({
        fhost::Bytecode bc = fhost::make_bytecode(t->t->d.asm_, t->t->d.n_regs, repack != 0);
        if (n_words) *n_words = bc.words.size();
        if (reg_count) *reg_count = bc.reg_count;
        if (mem_count) *mem_count = bc.mem_count;
        if (words) {
            if (cap < bc.words.size()) throw std::runtime_error("bytecode buffer too small");
            memcpy(words, bc.words.data(), bc.words.size() * 4);
        }
    });
}

void orc_screen_to_world_2d(uint32_t w, uint32_t h, float* m16) { Mat4 m = screen_to_world_2d(w, h); memcpy(m16, m.m, 64); }
void orc_screen_to_world_3d(uint32_t w, uint32_t h, uint32_t d, float* m16) { Mat4 m = screen_to_world_3d(w, h, d); memcpy(m16, m.m, 64); }
void orc_pixel_mat(uint32_t w, uint32_t h, const float* wm9, float* m16) { Mat4 m = pixel_mat(w, h, wm9); memcpy(m16, m.m, 64); }
void orc_mat4_mul(const float* a, const float* b, float* out) { Mat4 m = mat4_mul(to_mat(a), to_mat(b)); memcpy(out, m.m, 64); }
void orc_transform_f32(const float* m16, float x, float y, float z, float* out3) { transform_f32(x, y, z, to_mat(m16), out3); }

int32_t orc_render2d(const orc_tape* t, uint32_t w, uint32_t h, const float* mat16, float z, int32_t pixel_perfect,
                     const uint32_t* tile_sizes, uint32_t n_tile_sizes, int32_t threads, uint32_t first_root,
                     uint32_t n_roots, const float* var_values, uint32_t n_var_values, float* out, orc_stats* stats) {
    ORC_TRY({
        Render2DConfig cfg;
        cfg.width = w; cfg.height = h; cfg.mat = to_mat(mat16); cfg.z = z;
        cfg.pixel_perfect = pixel_perfect != 0;
        if (n_tile_sizes) cfg.tile_sizes.assign(tile_sizes, tile_sizes + n_tile_sizes);
        cfg.threads = threads; cfg.first_root = first_root; cfg.n_roots = n_roots;
        if (n_var_values) cfg.var_values.assign(var_values, var_values + n_var_values);
        TileStats s;
        render2d(t->t, cfg, out, &s);
        copy_stats(s, stats);
    });
}
int32_t orc_render3d(const orc_tape* t, uint32_t w, uint32_t h, uint32_t d, const float* mat16,
                     const uint32_t* tile_sizes, uint32_t n_tile_sizes, int32_t threads, uint32_t first_root,
                     uint32_t n_roots, const float* var_values, uint32_t n_var_values, void* out, orc_stats* stats) {
    ORC_TRY({
        Render3DConfig cfg;
        cfg.width = w; cfg.height = h; cfg.depth = d; cfg.mat = to_mat(mat16);
        if (n_tile_sizes) cfg.tile_sizes.assign(tile_sizes, tile_sizes + n_tile_sizes);
        cfg.threads = threads; cfg.first_root = first_root; cfg.n_roots = n_roots;
        if (n_var_values) cfg.var_values.assign(var_values, var_values + n_var_values);
        TileStats s;
        render3d(t->t, cfg, reinterpret_cast<GeometryPixel*>(out), &s);
        copy_stats(s, stats);
    });
}

// Octree sampler.  leaves: buffer of `cap` OctreeLeaf (348 bytes each) or NULL to count.  The result of
// a counting call is kept, so that the fetching call that follows (same tape, depth, transform) copies it
// instead of sampling again.
static int32_t octree_sample_impl(const orc_tape* t, uint32_t depth, const float* world_to_model16, int threads, void* leaves,
                                  uint64_t cap, uint64_t* n_leaves, uint64_t* stats) {
    ORC_TRY({
        static std::mutex mu;
        static std::vector<OctreeLeaf> last;
        static OctreeStats last_st;
        static const orc_tape* last_t = nullptr;
        static uint32_t last_depth = 0;
        static std::vector<float> last_m;
        std::lock_guard<std::mutex> g(mu);
        std::vector<float> m(world_to_model16 ? world_to_model16 : nullptr, world_to_model16 ? world_to_model16 + 16 : nullptr);
        const bool reuse = leaves && last_t == t && last_depth == depth && last_m == m;
        if (!reuse) {
            OctreeConfig cfg;
            cfg.depth = depth;
            cfg.threads = threads;
            if (world_to_model16) { cfg.has_transform = true; cfg.world_to_model = to_mat(world_to_model16); }
            octree_sample(t->t, cfg, last, &last_st);
            last_t = t; last_depth = depth; last_m = m;
        }
        const OctreeStats& st = last_st;
        if (n_leaves) *n_leaves = last.size();
        if (stats) {
            memcpy(stats, st.evaluated, 16 * 8); memcpy(stats + 16, st.full, 16 * 8);
            memcpy(stats + 32, st.empty, 16 * 8); memcpy(stats + 48, st.ambiguous, 16 * 8);
            stats[64] = st.leaf_empty; stats[65] = st.leaf_full; stats[66] = st.leaf_surface;
            stats[67] = st.float_points; stats[68] = st.grad_points;
        }
        if (leaves) {
            if (cap < last.size()) throw std::runtime_error("leaf buffer too small");
            memcpy(leaves, last.data(), last.size() * sizeof(OctreeLeaf));
            std::vector<OctreeLeaf>().swap(last);   // the fetch consumes the cached result
            last_t = nullptr;
        }
    });
}
int32_t orc_octree_sample(const orc_tape* t, uint32_t depth, const float* world_to_model16, void* leaves, uint64_t cap,
                          uint64_t* n_leaves, uint64_t* stats /* [16*4 + 5] */) {
    return octree_sample_impl(t, depth, world_to_model16, 1, leaves, cap, n_leaves, stats);
}
int32_t orc_octree_sample_mt(const orc_tape* t, uint32_t depth, const float* world_to_model16, int32_t threads, void* leaves,
                             uint64_t cap, uint64_t* n_leaves, uint64_t* stats) {
    return octree_sample_impl(t, depth, world_to_model16, threads, leaves, cap, n_leaves, stats);
}
static_assert(sizeof(OctreeLeaf) == 348, "OctreeLeaf layout");

// fidget-raster effects (oracle/effects.h); all images are host arrays, row-major
void orc_denoise_normals(const void* image, uint32_t w, uint32_t h, void* out) {
    denoise_normals(static_cast<const GeoPixel*>(image), w, h, static_cast<GeoPixel*>(out));
}
void orc_compute_ssao(const void* image, uint32_t w, uint32_t h, uint32_t d, const float* kernel, uint32_t nk,
                      const float* noise, uint32_t nn, float* out) {
    compute_ssao(static_cast<const GeoPixel*>(image), w, h, d, kernel, nk, noise, nn, out);
}
void orc_blur_ssao(const float* ssao, uint32_t w, uint32_t h, float* out) { blur_ssao(ssao, w, h, out); }
void orc_apply_shading(const void* image, uint32_t w, uint32_t h, uint32_t d, const float* ssao, uint8_t* out) {
    apply_shading(static_cast<const GeoPixel*>(image), w, h, d, ssao, out);
}
void orc_normals_to_color(const void* image, uint64_t n, uint8_t* out) {
    normals_to_color(static_cast<const GeoPixel*>(image), n, out);
}
void orc_to_rgba_bitmap(const float* image, uint64_t n, int32_t transparent, uint8_t* out) {
    to_rgba_bitmap(image, n, transparent != 0, out);
}
void orc_to_debug_bitmap(const float* image, uint64_t n, uint8_t* out) { to_debug_bitmap(image, n, out); }
void orc_to_rgba_distance(const float* image, uint64_t n, uint8_t* out) { to_rgba_distance(image, n, out); }

}  // extern "C"
