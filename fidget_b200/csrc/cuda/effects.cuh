// Launchers of the fidget-raster post-processing kernels (effects.cu).  All pointers are device
// pointers; images are row-major.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace fdev {

struct GeoPixel { float normal[3]; uint32_t depth; };   // voxel.rs:126-134, 16 bytes

void launch_denoise_normals(const GeoPixel* img, uint32_t w, uint32_t h, GeoPixel* out, cudaStream_t s);
void launch_compute_ssao(const GeoPixel* img, uint32_t w, uint32_t h, uint32_t d, const float* kernel, uint32_t nk,
                         const float* noise, uint32_t nn, float* out, cudaStream_t s);
void launch_blur_ssao(const float* ssao, uint32_t w, uint32_t h, float* out, cudaStream_t s);
// ssao may be null; blur != 0 applies blur_ssao on the fly to a raw occlusion map
void launch_apply_shading(const GeoPixel* img, uint32_t w, uint32_t h, uint32_t d, const float* ssao, int blur,
                          uint8_t* out_rgb, cudaStream_t s);
void launch_normals_to_color(const GeoPixel* img, uint64_t n, uint8_t* out_rgb, cudaStream_t s);
// mode 0: to_rgba_bitmap, 1: to_rgba_bitmap(transparent), 2: to_debug_bitmap, 3: to_rgba_distance
// inside/outside mask: one byte (255 / 0) per pixel, or one bit per pixel (rows padded to whole bytes)
void launch_to_mask(const float* img, uint32_t w, uint32_t h, uint8_t* out, int one_bit, cudaStream_t s);
void launch_to_rgba(int mode, const float* img, uint64_t n, uint8_t* out_rgba, cudaStream_t s);

}  // namespace fdev
