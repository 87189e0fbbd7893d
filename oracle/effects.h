// ORACLE -- test infrastructure, not product code.
//
// CPU restatement of fidget-raster's post-processing effects
// (fidget-raster/src/effects.rs:13-547) and GeometryPixel::to_color
// (fidget-raster/src/voxel.rs:136-153).
//
// PARITY UNPINNED BY REFERENCE VECTORS: effects.rs has no tests or fixtures and draws its SSAO
// tables from rand::rng().  tests/test_effects_oracle.py checks this file against the source's
// literal colour tables, closed-form cases and an independent float32 numpy transcription.
#pragma once
#include <cstdint>

namespace oracle {

struct GeoPixel { float normal[3]; uint32_t depth; };   // voxel.rs:126-134

// effects.rs:17-36
void denoise_normals(const GeoPixel* image, uint32_t w, uint32_t h, GeoPixel* out);
// effects.rs:72-95; kernel: 3 floats per sample (columns of the 3xN matrix), noise: 2 floats per entry.
// The reference draws both tables from rand::rng() (effects.rs:385-440), i.e. they differ on
// every call; here they are inputs.
void compute_ssao(const GeoPixel* image, uint32_t w, uint32_t h, uint32_t d, const float* kernel,
                  uint32_t n_kernel, const float* noise, uint32_t n_noise, float* out);
// effects.rs:98-115
void blur_ssao(const float* ssao, uint32_t w, uint32_t h, float* out);
// effects.rs:42-66 with the (already blurred) occlusion map as an input; ssao may be null
void apply_shading(const GeoPixel* image, uint32_t w, uint32_t h, uint32_t d, const float* ssao, uint8_t* out_rgb);
// voxel.rs:136-153
void normals_to_color(const GeoPixel* image, uint64_t n, uint8_t* out_rgb);
// effects.rs:446-467
void to_rgba_bitmap(const float* image, uint64_t n, bool transparent, uint8_t* out_rgba);
// effects.rs:470-497
void to_debug_bitmap(const float* image, uint64_t n, uint8_t* out_rgba);
// effects.rs:504-547
void to_rgba_distance(const float* image, uint64_t n, uint8_t* out_rgba);

}  // namespace oracle
