// ORACLE -- test infrastructure, not product code.
//
// CPU restatement of the reference's VM backend and of the callers that drive
// the tape-evaluation hot path:
//   VmIntervalEval / VmPointEval / VmFloatSliceEval / VmGradSliceEval
//                         <-> fidget-core/src/vm/mod.rs:324-1397
//   VmData::simplify      <-> fidget-core/src/vm/data.rs:123-318
//   RenderHandle          <-> fidget-core/src/render/mod.rs:19-179
//   Transformable         <-> fidget-core/src/shape/mod.rs:894-948
//   screen_to_world       <-> fidget-core/src/render/region.rs:87-108
//   pixel::render         <-> fidget-raster/src/pixel.rs:276-492
//   voxel::render         <-> fidget-raster/src/voxel.rs:216-553
//
// The tape front end (Context, SSA flattening, register allocation) is shared
// with the product's host library (fidget_b200/csrc/host/tape.{h,cc}): it is
// upstream of the hot path and is pinned separately by the reference's own
// tape-shape and bytecode unit tests (tests/test_tape_frontend.py).
#pragma once
#include <memory>
#include <vector>

#include "../fidget_b200/csrc/host/tape.h"
#include "types.h"

namespace oracle {

using fhost::Clause;
using fhost::TapeData;

struct Tape {
    TapeData d;
    std::vector<Clause> eval_order;  // asm reversed (evaluation order)
    explicit Tape(TapeData t);
    size_t size() const { return d.asm_.len(); }  // Function::size()
    uint32_t choice_count() const { return d.ssa.choice_count; }
    uint32_t slot_count() const { return d.asm_.slot_count; }
    size_t n_vars() const { return d.vars.size(); }
};
using TapeP = std::shared_ptr<const Tape>;

// Tracing evaluators; `choices` is resized to choice_count and OR-accumulated
// from Unknown.  Return value: true iff a trace is available (some choice is
// not Both).
struct IntervalEval {
    std::vector<Interval> slots;
    std::vector<uint8_t> choices;
    bool eval(const Tape& t, const Interval* vars, Interval* out);
};
struct PointEval {
    std::vector<float> slots;
    std::vector<uint8_t> choices;
    bool eval(const Tape& t, const float* vars, float* out);
};
struct FloatSliceEval {
    std::vector<std::vector<float>> slots;
    // vars[i] points at n floats; out[o] receives n floats
    void eval(const Tape& t, const float* const* vars, size_t n, float* const* out);
};
struct GradSliceEval {
    std::vector<std::vector<Grad>> slots;
    void eval(const Tape& t, const Grad* const* vars, size_t n, Grad* const* out);
};

// VmData::simplify with the same register count as the parent
TapeP simplify(const Tape& parent, const uint8_t* choices, size_t n_choices, uint32_t n_regs);

// RenderHandle: lazily simplified chain with a one-entry trace cache
struct RenderHandle {
    TapeP shape;
    std::vector<uint8_t> next_trace;
    std::unique_ptr<RenderHandle> next;
    explicit RenderHandle(TapeP s) : shape(std::move(s)) {}
    RenderHandle* simplify(const std::vector<uint8_t>& trace);
};

struct Mat4 { float m[4][4]; };  // row-major m[row][col]
Mat4 mat4_identity();
Mat4 mat4_mul(const Mat4& a, const Mat4& b);
// RegionSize::screen_to_world for 2D (embedded as 4x4, z preserved) and 3D
Mat4 screen_to_world_2d(uint32_t w, uint32_t h);
Mat4 screen_to_world_3d(uint32_t w, uint32_t h, uint32_t d);
// pixel::RenderConfig::mat -> 4x4 (world_to_model is 3x3 row-major)
Mat4 pixel_mat(uint32_t w, uint32_t h, const float world_to_model[9]);

void transform_f32(float x, float y, float z, const Mat4& m, float out[3]);
void transform_interval(Interval x, Interval y, Interval z, const Mat4& m, Interval out[3]);
void transform_grad(Grad x, Grad y, Grad z, const Mat4& m, Grad out[3]);

struct TileStats {  // per-level tile census, for tile-mask parity checks
    uint64_t evaluated[8] = {0}, filled_inside[8] = {0}, filled_outside[8] = {0}, ambiguous[8] = {0};
    uint64_t pixels = 0;  // points evaluated by the bulk evaluator
    uint64_t simplified[8] = {0};  // simplifications kept (shorter than parent)
};

struct Render2DConfig {
    uint32_t width = 0, height = 0;
    Mat4 mat;  // full 4x4 transform (screen -> model)
    float z = 0.0f;
    bool pixel_perfect = false;
    std::vector<uint32_t> tile_sizes = {128, 32, 8};
    int threads = 1;
    // values for non-XYZ variables, indexed by tape input slot (ShapeVars, shape/mod.rs:548-640)
    std::vector<float> var_values;
    // restrict to root tiles [first, first+count) in the reference's
    // enumeration order (x-major); count = 0 means all
    uint32_t first_root = 0, n_roots = 0;
};
// out: width*height RawDistancePixel bit patterns (as float), row-major
void render2d(const TapeP& tape, const Render2DConfig& cfg, float* out, TileStats* stats);

struct GeometryPixel { float normal[3]; uint32_t depth; };
struct Render3DConfig {
    uint32_t width = 0, height = 0, depth = 0;
    Mat4 mat;
    std::vector<uint32_t> tile_sizes = {128, 64, 32, 16, 8};
    int threads = 1;
    std::vector<float> var_values;
    uint32_t first_root = 0, n_roots = 0;
};
void render3d(const TapeP& tape, const Render3DConfig& cfg, GeometryPixel* out, TileStats* stats);

// TileSizesRef::new (fidget-raster/src/lib.rs:59-66)
std::vector<uint32_t> trim_tile_sizes(const std::vector<uint32_t>& ts, uint32_t max_size);

}  // namespace oracle
