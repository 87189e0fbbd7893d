// Kernel-side data structures shared between kernels.cu and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_ops.cuh"

namespace fdev {

constexpr int MAX_LEVELS = 16;             // 3D renders use <= 8; the octree sampler uses depth + 1
constexpr int WARPS_PER_BLOCK = 4;          // interval kernels
constexpr int REG_SLOTS = 256;              // register slots of the fast interpreters

// A tape living in the device arena (units of uint2 clauses)
struct TapeRef {
    const uint2* ptr;     // first clause (device memory: a root tape buffer or the arena)
    uint32_t n_ops;       // device clauses
    uint32_t ref_len;     // RegTape::len() of the reference's equivalent tape
    uint32_t n_choices;   // choice clauses (== reference choice_count)
    uint32_t pad;
};

// One tile that survived the previous level ("ambiguous"), plus the tape its
// children are evaluated with.
struct TileJob {
    uint32_t x, y, z;     // corner in pixels/voxels
    uint32_t pad;
    TapeRef tape;
};  // 40 bytes

struct FillRec { uint32_t x, y, value, ready; };  // 2D fill: corner + RawDistancePixel bits + this render's ready mark

// One interval-evaluated 3D tile, recorded for the exact census (FC_FLAG_EXACT_CENSUS): the reference visits
// tiles front to back and skips those whose pixels are all finished (voxel.rs:283-293); which tiles that
// are is decided afterwards from the final heightmap (k_census_3d)
struct CensusRec { uint16_t x, y, z; uint8_t level, flags; };   // flags: 0 outside, 1 inside, 2 ambiguous; | 4 simplified tape kept

struct Stats {
    unsigned long long evaluated[MAX_LEVELS];
    unsigned long long filled_inside[MAX_LEVELS];
    unsigned long long filled_outside[MAX_LEVELS];
    unsigned long long ambiguous[MAX_LEVELS];
    unsigned long long simplified[MAX_LEVELS];
    unsigned long long pixels;
    unsigned long long grads;
    unsigned long long culled[MAX_LEVELS];
};

// Device-side counters of one render call
struct Counters {
    uint32_t n_jobs[MAX_LEVELS + 1];   // n_jobs[l] = tiles queued FOR level l (parents whose children are evaluated at l)
    uint32_t n_fills[MAX_LEVELS];      // 2D fill records produced by level l
    uint32_t cursor[MAX_LEVELS + 2];   // dynamic work cursors (one per kernel)
    uint32_t error;                    // bit 0: arena exhausted, bit 1: list overflow, bit 2: fused kernel watchdog
    uint32_t outstanding;              // fused 2D kernel: interval / pixel jobs queued or running
    uint32_t fill_cursor[MAX_LEVELS];  // fused 2D kernel: fill records painted so far, per level
    uint32_t n_census;                 // exact 3D census: records appended
    uint32_t pad;
    unsigned long long arena_top;      // bump pointer (clauses)
};

// Cooperative level-0 schedule (built on the host at fc_tape_create): the root
// tape in SSA form (value id = position of the defining clause), clauses
// grouped into dependency waves; a trailing run of single-clause waves is the
// serial tail.
constexpr int COOP_THREADS = 256;
constexpr uint32_t COOP_NONE = 0xFFFFu;
struct CoopRec {
    uint32_t x, y;        // the device clause
    uint16_t ia, ib;      // defining positions of the lhs / rhs register operands (COOP_NONE: immediate/unused)
    uint16_t p;           // position of this clause in the tape (== id of the value it defines)
    uint16_t cidx;        // choice index (choice clauses only)
};
// The tail is cut into segments: SERIAL runs are executed by one thread,
// CHAIN runs (>= 8 consecutive min- or max-clauses, each combining the
// previous clause's result with a value computed before the run) are
// evaluated with a block-wide prefix scan: min/max of intervals is exactly
// associative, and the choices follow from the prefix values.
struct CoopSeg { uint32_t begin, end, chain, start_slot; };   // start_slot: slot of the value a chain starts from
// The forward pass addresses values by SLOT, not by defining clause: the host colours the
// values of the schedule so that a slot is reused once every reader of its value has run
// (prospero: 6363 values -> ~2700 slots), which is what lets seven root tiles share an SM.
// A chain value whose only reader is the next clause of the same chain gets no slot at all.
struct CoopFwd {
    uint32_t x, y;        // the device clause
    uint16_t sa, sb;      // slots of the lhs / rhs register operands (COOP_NONE: immediate/unused)
    uint16_t so;          // slot of the result (COOP_NONE: not stored)
    uint16_t cidx;        // choice index (choice clauses only)
};
constexpr int COOP_MAX_SEGS = 16;
struct CoopSched {
    const CoopFwd* fwd;           // forward view of the same records (slots instead of positions)
    uint32_t n_slots;
    const CoopRec* recs;
    const uint32_t* wave_start;   // [n_waves + 1] offsets into recs
    uint32_t n_waves;
    uint32_t tail_begin, tail_end;
    uint32_t n_segs;
    CoopSeg segs[COOP_MAX_SEGS];
};

// Where the renderers feed coordinates: input slots of X, Y, Z (-1 = unused) and the
// values bound to every other input slot (ShapeVars, shape/mod.rs:548-640).
constexpr int MAX_RENDER_VARS = 16;
struct VarBind {
    int x, y, z;
    float values[MAX_RENDER_VARS];
};

struct LevelParams {
    CoopSched sched;
    int level;                 // index into tile sizes
    uint32_t tile;             // edge of the tiles evaluated by this launch
    uint32_t n_axis;           // children per axis of each parent job (level > 0)
    uint32_t is_last;          // children are leaf tiles
    uint32_t pixel_perfect;
    // level 0 enumerates root tiles itself
    uint32_t root_mode;
    uint32_t roots_x, roots_y, roots_z;     // root grid
    uint32_t root_x0, root_y0, root_z0;     // origin of the root grid (pixels)
    // multi-GPU tile interleave: when non-null only the XY root tiles listed here (ty * roots_x + tx inside
    // the band) are evaluated, at every Z layer
    const uint32_t* root_list;
    uint32_t n_root_list;
    TapeRef root_tape;
    uint32_t epoch;                // ready mark of this render's job and fill records (never 0)
    // image
    uint32_t width, height, depth;
    float z2d;
    Mat4 mat;
    // lists
    const TileJob* jobs_in;
    uint32_t cap_in;
    TileJob* jobs_out;
    uint32_t cap_out;
    FillRec* fills;
    uint32_t cap_fills;
    // arena
    uint2* arena;
    unsigned long long arena_cap;   // clauses
    // scratch: per-warp choice words [choice_words][32]
    uint32_t* choice_scratch;
    uint32_t choice_words;          // words per lane
    Counters* ctr;
    Stats* stats;
    // octree sampler (mode 1): coordinates are cells at the finest depth; bounds = coord * cell_h - 1
    uint32_t mode;
    uint32_t has_transform;
    float cell_h;
    // 3D
    unsigned long long* heightmap;  // 3D: width*height keys (depth << 32 | leaf job id + 1), atomicMax
    // 3D occlusion map: per 16 x 16 block of pixels, a lower bound of the depth EVERY pixel of the block already has
    // (raised by interval-proven-inside tiles that cover whole blocks).  A parent whose blocks all reach its top + 1
    // cannot show anything: its children are skipped (cull != 0: this level's parents are made of whole blocks).
    uint32_t* occl;
    uint32_t occl_w, occl_h;        // blocks per row, block rows (tiles may overhang a ragged image: blocks outside are skipped)
    uint32_t cull;
    CensusRec* census;              // exact 3D census records (or null)
    uint32_t cap_census;
    VarBind vb;
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t root_count(const LevelParams& p, bool with_z) {
    return (p.root_list ? p.n_root_list : p.roots_x * p.roots_y) * (with_z ? p.roots_z : 1u);
}
// corner of root tile `idx` (XY-major, then Z layers)
__device__ __forceinline__ void root_corner(const LevelParams& p, uint32_t idx, uint32_t T, uint32_t& cx, uint32_t& cy,
                                            uint32_t& cz) {
    const uint32_t n_xy = p.root_list ? p.n_root_list : p.roots_x * p.roots_y;
    uint32_t xy = idx % n_xy;
    const uint32_t zl = idx / n_xy;
    if (p.root_list) xy = __ldg(p.root_list + xy);
    cx = p.root_x0 + (xy % p.roots_x) * T;
    cy = p.root_y0 + (xy / p.roots_x) * T;
    cz = p.root_z0 + zl * T;
}
#endif

struct PixelParams {
    uint32_t tile;             // leaf tile edge
    uint32_t width, height;
    float z2d;
    Mat4 mat;
    const TileJob* jobs;
    float* out;
    Counters* ctr;
    int list;                  // which n_jobs entry holds the leaf count
    int cursor;
    Stats* stats;
    VarBind vb;
};

struct FillParams {
    uint32_t tile, width, height;
    const FillRec* fills;
    const uint32_t* n_fills;
    float* out;
};

struct VoxelParams {
    uint32_t tile, width, height;
    const uint32_t* order;      // leaf jobs sorted front to back (descending z), or null
    Mat4 mat;
    const TileJob* jobs;
    uint32_t cap_jobs;
    unsigned long long* heightmap;
    Counters* ctr;
    int list, cursor;
    Stats* stats;
    VarBind vb;
};
struct NormalParams {
    uint32_t width, height, depth;
    uint32_t y0, y1;            // rows to finish (a Y band of a sharded render, else 0..height)
    // tile interleave: only these root tiles (ty * roots_x + tx inside the band, edge `root_tile`) are finished
    const uint32_t* root_list;
    uint32_t n_root_list, roots_x, root_tile;
    uint32_t clamp;             // apply the final `depth >= D-1` clamp (voxel.rs:535-546)
    Mat4 mat;
    const TileJob* jobs;        // leaf jobs
    const unsigned long long* heightmap;
    void* out;                  // GeometryPixel[width*height]
    Stats* stats;
    VarBind vb;
};

// One leaf of the Manifold-Dual-Contouring octree (LeafHermiteData, fidget-mesh/src/octree.rs:864-900)
struct OctreeLeaf {
    uint16_t ix, iy, iz;
    uint8_t mask, n_edges;
    uint16_t present, pad;
    float pos[12][3];
    float grad[12][4];   // dx, dy, dz, v
};
struct OctreeLeafParams {
    const TileJob* jobs;
    uint32_t cap_jobs;
    Counters* ctr;
    int list, cursor;
    float cell_h;
    uint32_t has_transform;
    Mat4 mat;
    VarBind vb;
    OctreeLeaf* out;
    TapeRef* out_tapes;         // tape of each emitted leaf (consumed by the gradient pass)
    uint32_t cap_out;
    uint32_t* n_out;            // device counter
    unsigned long long* stats;  // [0] leaf_empty [1] leaf_full [2] leaf_surface [3] float points [4] grad points
};

// launchers (kernels.cu)
void launch_octree_leaf(const OctreeLeafParams& p, int blocks, cudaStream_t s);
void launch_octree_grads(const OctreeLeafParams& p, int blocks, cudaStream_t s);
void launch_interval_level_3d(const LevelParams& p, int blocks, cudaStream_t s);
void launch_voxels_3d(const VoxelParams& p, int blocks, cudaStream_t s);
// Counting sort of the leaf jobs by descending Z layer: hist/offsets are device scratch of n_layers+1 words
void launch_leaf_zsort(const TileJob* jobs, const uint32_t* n_jobs, uint32_t cap, uint32_t z0, uint32_t tile,
                       uint32_t n_layers, uint32_t* hist, uint32_t* order, cudaStream_t s);
void launch_normals_3d(const NormalParams& p, cudaStream_t s);
struct CensusParams {
    const CensusRec* recs;
    const uint32_t* n_recs;
    uint32_t cap;
    uint32_t tile[MAX_LEVELS];
    int last_level;
    const unsigned long long* heightmap;
    uint32_t width;
    Stats* stats;
};
void launch_census_3d(const CensusParams& p, int blocks, cudaStream_t s);
void launch_merge_slabs(const void* const* d_slabs, uint32_t n_slabs, uint32_t n_pixels, uint32_t depth, void* out,
                        cudaStream_t s);
// tile interleave: rank >= 0 packs that rank's tiles of `src` (an image) into `dst` (its chunk);
// rank < 0 unpacks every tile of the gathered chunks in `src` into the image `dst`
void launch_tiles_copy(const void* src, void* dst, uint32_t width, uint32_t height, uint32_t px_bytes, uint32_t T,
                       uint32_t roots_x, uint32_t roots_y, const uint32_t* slots, uint32_t n_ranks, uint32_t per_rank, int rank,
                       cudaStream_t s);
void launch_interval_level_2d(const LevelParams& p, int blocks, cudaStream_t s);
int coop_regs_per_thread(int dim);
int coop_occupancy(int dim, int threads, size_t smem);
size_t coop_smem_bytes(uint32_t n_ops, uint32_t n_choices, uint32_t n_slots);
cudaError_t launch_interval_root_coop_2d(const LevelParams& p, int blocks, int threads, cudaStream_t s);
cudaError_t launch_interval_root_coop_3d(const LevelParams& p, int blocks, int threads, cudaStream_t s);
void launch_pixels_2d(const PixelParams& p, int blocks, cudaStream_t s);
// Fused 2D tail (tail2d.cu): every level after the root level, the leaf pixels and the fills in ONE
// persistent launch that drains a dependency-ordered queue
constexpr int TAIL_MAX_LEVELS = 4;
struct Tail2DParams {
    int n_levels;                      // interval levels handled here (levels 1 .. n_levels of the render)
    LevelParams lv[TAIL_MAX_LEVELS];   // lv[k] = parameters of render level k + 1
    PixelParams px;
    uint32_t fill_tile[TAIL_MAX_LEVELS + 1];        // tile edge of the fill records of render level l
    const FillRec* fills[TAIL_MAX_LEVELS + 1];
    uint32_t fill_cap[TAIL_MAX_LEVELS + 1];
    uint32_t epoch;
    uint32_t paint_fills;              // 1: idle warps paint the fill records; 0: k_fill_2d launches do (after / beside this kernel)
};
cudaError_t launch_tail_2d(const Tail2DParams& p, int sm_count, cudaStream_t s);
int tail_2d_blocks(int sm_count);
void launch_fill_2d(const FillParams& p, int blocks, cudaStream_t s);

// trait-level evaluators
struct BulkParams {
    const uint2* tape;
    uint32_t n_ops;
    uint32_t n_vars, n_outputs;
    uint32_t n_slots;           // 256 + mem_count
    uint64_t n;
    const void* const* vars;    // device array of device pointers (n_vars)
    void* const* outs;          // device array of device pointers (n_outputs)
};
void launch_float_slice(const BulkParams& p, cudaStream_t s);
// TMA-fed persistent bulk evaluators (bulk.cu): float4 columns, i.e. four consecutive f32 points or one
// gradient per thread
struct SliceTmaParams {
    const uint2* tape;
    uint32_t n_ops, n_vars, n_outputs, n_regs;
    uint64_t n;                 // points
    const float4* vars[4];
    float4* outs[2];
};
bool launch_slice_tma(const SliceTmaParams& p, bool grad, int sm_count, cudaStream_t s);
void launch_grad_slice(const BulkParams& p, cudaStream_t s);

struct TracingParams {
    const uint2* tape;
    uint32_t n_ops, n_vars, n_outputs, n_choices, n_slots;
    uint64_t n;
    const float* vars;   // interval: [n][n_vars][2]; point: [n][n_vars]
    float* out;          // interval: [n][n_outputs][2]; point: [n][n_outputs]
    uint8_t* choices;    // [n][n_choices] or null
    uint8_t* simplify;   // [n] or null
};
void launch_interval_batch(const TracingParams& p, cudaStream_t s);
void launch_point_batch(const TracingParams& p, cudaStream_t s);

struct SimplifyParams {
    const uint2* parent;
    uint32_t n_ops, parent_ref_len;
    const uint8_t* choices;   // [n_choices] bytes
    uint32_t n_choices;
    uint2* out;               // capacity n_ops; child occupies the TAIL
    uint32_t* result;         // {n_dev, ref_len, n_choices_child}
};
void launch_simplify_single(const SimplifyParams& p, cudaStream_t s);

}  // namespace fdev
