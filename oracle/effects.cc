// ORACLE -- test infrastructure, not product code.
//
// CPU restatement of fidget-raster/src/effects.rs.  Vector arithmetic follows
// nalgebra's evaluation order for fixed 3-vectors: dot = (a0*b0 + a1*b1) + a2*b2,
// normalize = componentwise division by sqrt(dot(v, v)), mat3 * vec3 accumulates
// column by column (col0*k0, then + col1*k1, then + col2*k2).
#include "effects.h"

#include <cmath>
#include <cstring>

#include "types.h"

namespace oracle {
namespace {

struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 normalize(V3 a) { return a / std::sqrt(dot(a, a)); }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// Rust `f as u8`: saturating, truncating, NaN -> 0
inline uint8_t as_u8(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 255.0f) return 255;
    return uint8_t(int(f));
}
// f32::clamp (NaN passes through)
inline float clampf(float x, float lo, float hi) {
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return x;
}
// ordered_float::OrderedFloat ordering: NaN is the greatest value and equal to itself
inline int ord_cmp(float a, float b) {
    const bool an = a != a, bn = b != b;
    if (an) return bn ? 0 : 1;
    if (bn) return -1;
    return a < b ? -1 : (a > b ? 1 : 0);
}

const int WIN[4][2] = {{0, 0}, {-1, 0}, {0, -1}, {-1, -1}};   // scaled by the radius

// effects.rs:256-326
void denoise_pixel(const GeoPixel* img, int w, int h, int x, int y, int r, float out[3]) {
    const float* n = img[size_t(y) * w + x].normal;
    if (n[2] > 0.0f) { memcpy(out, n, 12); return; }
    bool have = false;
    float best_score = 0.0f;
    V3 best{n[0], n[1], n[2]};
    for (int k = 0; k < 4; ++k) {
        const int xmin = WIN[k][0] * r, ymin = WIN[k][1] * r;
        V3 sum{0, 0, 0};
        int count = 0;
        for (int i = 0; i <= r; ++i)
            for (int j = 0; j <= r; ++j) {
                const int tx = x + xmin + i, ty = y + ymin + j;
                if (tx >= 0 && ty >= 0 && tx < w && ty < h) {
                    const GeoPixel& p = img[size_t(ty) * w + tx];
                    if (p.depth != 0 && p.normal[2] > 0.0f) {
                        sum = sum + V3{p.normal[0], p.normal[1], p.normal[2]};
                        ++count;
                    }
                }
            }
        if (!count) continue;
        const V3 mean = sum / float(count);
        float score = 0.0f;
        for (int i = 0; i <= r; ++i)
            for (int j = 0; j <= r; ++j) {
                const int tx = x + xmin + i, ty = y + ymin + j;
                if (tx >= 0 && ty >= 0 && tx < w && ty < h) {
                    const GeoPixel& p = img[size_t(ty) * w + tx];
                    if (p.depth != 0) score += dot(V3{p.normal[0], p.normal[1], p.normal[2]}, mean);
                }
            }
        // Iterator::max_by_key keeps the LAST of several equal maxima
        if (!have || ord_cmp(best_score, score) <= 0) { best_score = score; best = mean; have = true; }
    }
    out[0] = best.x; out[1] = best.y; out[2] = best.z;
}

// effects.rs:329-381
float blur_pixel(const float* s, int w, int h, int x, int y, int r) {
    bool have = false;
    float best_dev = 0.0f, best_mean = s[size_t(y) * w + x];
    for (int k = 0; k < 4; ++k) {
        const int xmin = WIN[k][0] * r, ymin = WIN[k][1] * r;
        float sum = 0.0f;
        int count = 0;
        for (int i = 0; i <= r; ++i)
            for (int j = 0; j <= r; ++j) {
                const int tx = x + xmin + i, ty = y + ymin + j;
                if (tx >= 0 && ty >= 0 && tx < w && ty < h) {
                    const float v = s[size_t(ty) * w + tx];
                    if (v == v) { sum += v; ++count; }
                }
            }
        if (!count) continue;
        const float mean = sum / float(count);
        float stdev = 0.0f;
        for (int i = 0; i <= r; ++i)
            for (int j = 0; j <= r; ++j) {
                const int tx = x + xmin + i, ty = y + ymin + j;
                if (tx >= 0 && ty >= 0 && tx < w && ty < h) {
                    const float v = s[size_t(ty) * w + tx];
                    if (v == v) { const float e = mean - v; stdev += e * e; }
                }
            }
        const float dev = stdev / float(count);
        // Iterator::min_by_key keeps the FIRST of several equal minima
        if (!have || ord_cmp(best_dev, dev) > 0) { best_dev = dev; best_mean = mean; have = true; }
    }
    return best_mean;
}

// effects.rs:159-253
float ssao_pixel(const GeoPixel* img, uint32_t w, uint32_t h, uint32_t d, uint32_t x, uint32_t y,
                 const float* kernel, uint32_t nk, const float* noise, uint32_t nn) {
    const GeoPixel& px0 = img[size_t(y) * w + x];
    if (px0.depth == 0) return u2f(0x7FC00000u);
    const float fw = float(w), fh = float(h), fd = float(d);
    uint32_t m = w < h ? w : h;
    if (d < m) m = d;
    const float scale_min = float(m);
    const float scale_x = scale_min / fw, scale_y = scale_min / fh, scale_z = scale_min / fd;
    const V3 p{(((float(x) + 0.5f) / fw) - 0.5f) * 2.0f, (((float(y) + 0.5f) / fh) - 0.5f) * 2.0f,
               ((float(px0.depth) / fd) - 0.5f) * 2.0f};
    const V3 n = normalize(V3{px0.normal[0], px0.normal[1], px0.normal[2]});
    const uint32_t ri = rng_mix(y, x) % nn;
    const V3 rvec{noise[2 * ri], noise[2 * ri + 1], 0.0f};
    const V3 tangent = normalize(rvec - n * dot(rvec, n));
    const V3 bitangent = cross(n, tangent);
    const float RADIUS = 0.1f;
    float occlusion = 0.0f;
    for (uint32_t i = 0; i < nk; ++i) {
        const float k0 = kernel[3 * i], k1 = kernel[3 * i + 1], k2 = kernel[3 * i + 2];
        V3 off{tangent.x * k0, tangent.y * k0, tangent.z * k0};
        off = V3{bitangent.x * k1 + off.x, bitangent.y * k1 + off.y, bitangent.z * k1 + off.z};
        off = V3{n.x * k2 + off.x, n.y * k2 + off.y, n.z * k2 + off.z};
        off = off * RADIUS;
        off.x *= scale_x;
        off.y *= scale_y;
        off.z *= scale_z;
        const V3 sp = off + p;
        const float px = ((sp.x / 2.0f) + 0.5f) * fw;
        const float py = ((sp.y / 2.0f) + 0.5f) * fh;
        uint32_t actual_h = 0;
        if (px < fw && py < fh && px > 0.0f && py > 0.0f) actual_h = img[size_t(uint32_t(py)) * w + uint32_t(px)].depth;
        const float actual_z = ((float(actual_h) / fd) - 0.5f) * 2.0f;
        const float dz = sp.z - actual_z;
        if (dz < RADIUS) {
            occlusion += (sp.z <= actual_z) ? 1.0f : 0.0f;
        } else if (dz < RADIUS * 2.0f && sp.z <= actual_z) {
            const float t = (RADIUS - (dz - RADIUS)) / RADIUS;
            occlusion += t * t;
        }
    }
    return 1.0f - (occlusion / float(nk));
}

// effects.rs:118-154
void shade_pixel(const GeoPixel* img, uint32_t w, uint32_t h, uint32_t d, const float* ssao, uint32_t x, uint32_t y,
                 uint8_t out[3]) {
    const GeoPixel& g = img[size_t(y) * w + x];
    const V3 n = normalize(V3{g.normal[0], g.normal[1], g.normal[2]});
    const V3 p{2.0f * (float(x) / float(w) - 0.5f), 2.0f * (float(y) / float(h) - 0.5f),
               2.0f * (float(g.depth) / float(d) - 0.5f)};
    const float lights[3][4] = {{5.0f, -5.0f, 10.0f, 0.5f}, {-5.0f, 0.0f, 10.0f, 0.15f}, {0.0f, -5.0f, 10.0f, 0.15f}};
    float accum = 0.2f;
    for (auto& l : lights) {
        const V3 dir = normalize(V3{l[0], l[1], l[2]} - p);
        accum += rmax(dot(dir, n), 0.0f) * l[3];
    }
    if (ssao) accum *= ssao[size_t(y) * w + x] * 0.6f + 0.4f;
    accum = clampf(accum, 0.0f, 1.0f);
    const uint8_t c = as_u8(accum * 255.0f);
    out[0] = out[1] = out[2] = c;
}

const uint32_t KEY = 0xF6u << 9, KEY_MASK = 0xFFu << 9;   // pixel.rs:180-181
inline bool is_distance(float f) { return f == f || (f2u(f) & KEY_MASK) != KEY; }   // pixel.rs:197-203

}  // namespace

void denoise_normals(const GeoPixel* image, uint32_t w, uint32_t h, GeoPixel* out) {
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            GeoPixel o{{0, 0, 0}, image[size_t(y) * w + x].depth};
            if (o.depth > 0) denoise_pixel(image, int(w), int(h), int(x), int(y), 2, o.normal);
            out[size_t(y) * w + x] = o;
        }
}

void compute_ssao(const GeoPixel* image, uint32_t w, uint32_t h, uint32_t d, const float* kernel, uint32_t n_kernel,
                  const float* noise, uint32_t n_noise, float* out) {
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x)
            out[size_t(y) * w + x] = ssao_pixel(image, w, h, d, x, y, kernel, n_kernel, noise, n_noise);
}

void blur_ssao(const float* ssao, uint32_t w, uint32_t h, float* out) {
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            const float v = ssao[size_t(y) * w + x];
            out[size_t(y) * w + x] = (v != v) ? u2f(0x7FC00000u) : blur_pixel(ssao, int(w), int(h), int(x), int(y), 2);
        }
}

void apply_shading(const GeoPixel* image, uint32_t w, uint32_t h, uint32_t d, const float* ssao, uint8_t* out) {
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            uint8_t* o = out + (size_t(y) * w + x) * 3;
            if (image[size_t(y) * w + x].depth > 0) shade_pixel(image, w, h, d, ssao, x, y, o);
            else o[0] = o[1] = o[2] = 0;
        }
}

void normals_to_color(const GeoPixel* image, uint64_t n, uint8_t* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const float dx = image[i].normal[0], dy = image[i].normal[1], dz = image[i].normal[2];
        const float s = std::sqrt((dx * dx + dy * dy) + dz * dz);
        uint8_t* o = out + i * 3;
        if (s != 0.0f) {
            const float scale = 255.0f / s;
            o[0] = as_u8(std::fabs(dx) * scale);
            o[1] = as_u8(std::fabs(dy) * scale);
            o[2] = as_u8(std::fabs(dz) * scale);
        } else o[0] = o[1] = o[2] = 0;
    }
}

void to_rgba_bitmap(const float* image, uint64_t n, bool transparent, uint8_t* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const float f = image[i];
        const bool inside = is_distance(f) ? (f < 0.0f) : ((f2u(f) & 1u) == 1u);   // pixel.rs:188-193
        const uint8_t v = inside ? 255 : 0, a = (inside || !transparent) ? 255 : 0;
        uint8_t* o = out + i * 4;
        o[0] = o[1] = o[2] = v;
        o[3] = a;
    }
}

void to_debug_bitmap(const float* image, uint64_t n, uint8_t* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const float f = image[i];
        uint8_t* o = out + i * 4;
        o[3] = 255;
        if (is_distance(f)) {
            o[0] = o[1] = o[2] = (f < 0.0f) ? 255 : 0;
        } else {
            const uint32_t bits = f2u(f);
            const uint8_t v = (bits & 1u) ? 255 : 50, depth = uint8_t(bits >> 1);
            o[0] = (depth == 0 || depth > 2) ? v : 0;
            o[1] = (depth == 1 || depth > 2) ? v : 0;
            o[2] = (depth == 2) ? v : 0;
        }
    }
}

void to_rgba_distance(const float* image, uint64_t n, uint8_t* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const float f = image[i];
        uint8_t* o = out + i * 4;
        o[3] = 255;
        if (!is_distance(f)) {
            const bool inside = f2u(f) & 1u;
            o[0] = inside ? 184 : 217; o[1] = inside ? 235 : 144; o[2] = inside ? 255 : 72;
        } else if (f != f) {
            o[0] = 255; o[1] = 0; o[2] = 0;
        } else {
            const float af = std::fabs(f);
            const float rgb[3] = {1.0f - std::copysign(0.1f, f), 1.0f - std::copysign(0.4f, f), 1.0f - std::copysign(0.7f, f)};
            const float dim = 1.0f - std::exp(-4.0f * af);
            const float bands = 0.8f + 0.2f * std::cos(140.0f * f);
            auto smoothstep = [](float e0, float e1, float x) {
                const float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
                return t * t * (3.0f - 2.0f * t);
            };
            auto mix = [](float x, float y, float a) { return x * (1.0f - a) + y * a; };
            for (int c = 0; c < 3; ++c) {
                float v = rgb[c] * dim * bands;
                v = mix(v, 1.0f, 1.0f - smoothstep(0.0f, 0.015f, af));
                v = mix(v, 1.0f, 1.0f - smoothstep(0.0f, 0.005f, af));
                o[c] = as_u8(clampf(v, 0.0f, 1.0f) * 255.0f);
            }
        }
    }
}

}  // namespace oracle
