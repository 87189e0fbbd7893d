// ORACLE -- test infrastructure, not product code.
//
// CPU restatement of Fidget's numeric types for the tape-evaluation hot path:
//   Interval  <-> fidget-core/src/types/interval.rs:12-744
//   Grad      <-> fidget-core/src/types/grad.rs:4-416
//   f32 ops   <-> fidget-core/src/types/float.rs:66-142
//   Choice    <-> fidget-core/src/vm/choice.rs:13-29
//   rng       <-> fidget-core/src/rng/mod.rs:8-33
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use anything under oracle/.
//
// Deviations from the reference, all outside its defined behaviour:
//  * `Interval::new` panics on lower > upper (interval.rs:39-45); the oracle
//    stores whatever was computed instead of aborting.
//  * Rust's f32::min/max leave the sign of zero unspecified when comparing
//    +0 and -0; rmin/rmax below pick IEEE-754-2019 minimumNumber /
//    maximumNumber (-0 < +0), which is also what CUDA's fminf/fmaxf do.
// Must be compiled with -ffp-contract=off (Rust never fuses a*b+c).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace oracle {

enum Choice : uint8_t { C_UNKNOWN = 0, C_LEFT = 1, C_RIGHT = 2, C_BOTH = 3 };

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// f32::min / f32::max: the other operand if one is NaN
static inline float rmin(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == b) return std::signbit(a) ? a : b;
    return a < b ? a : b;
}
static inline float rmax(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == b) return std::signbit(a) ? b : a;
    return a > b ? a : b;
}

// rng/mod.rs:8-33
static inline uint32_t rng_hash(uint32_t v) {
    uint32_t state = v * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28) + 4)) ^ state) * 277803737u;
    return (word >> 22) ^ word;
}
static inline float rng_rand(uint32_t seed) { return u2f((rng_hash(seed) >> 9) | 0x3f800000u) - 1.0f; }
static inline uint32_t rng_mix(uint32_t a, uint32_t b) { return rng_hash(a + rng_hash(b)); }

// f32::rem_euclid / div_euclid (Rust std)
static inline float rem_euclid(float a, float b) {
    float r = std::fmod(a, b);
    return r < 0.0f ? r + std::fabs(b) : r;
}
static inline float div_euclid(float a, float b) {
    float q = std::trunc(a / b);
    if (std::fmod(a, b) < 0.0f) return b > 0.0f ? q - 1.0f : q + 1.0f;
    return q;
}

////////////////////////////////////////////////////////////////////////////
// f32 with choices (float.rs:66-142)
struct FC { float v; Choice c; };
static inline float f_compare(float a, float b) {
    if (a < b) return -1.0f;
    if (a > b) return 1.0f;
    if (a == b) return 0.0f;
    return NAN;
}
static inline FC f_max_choice(float a, float b) {
    if (a > b) return {a, C_LEFT};
    if (b > a) return {b, C_RIGHT};
    return {(a != a || b != b) ? NAN : b, C_BOTH};
}
static inline FC f_min_choice(float a, float b) {
    if (a < b) return {a, C_LEFT};
    if (b < a) return {b, C_RIGHT};
    return {(a != a || b != b) ? NAN : b, C_BOTH};
}
static inline FC f_and_choice(float a, float b) { return a == 0.0f ? FC{a, C_LEFT} : FC{b, C_RIGHT}; }
static inline FC f_or_choice(float a, float b) { return a != 0.0f ? FC{a, C_LEFT} : FC{b, C_RIGHT}; }
static inline float f_not(float a) { return a == 0.0f ? 1.0f : 0.0f; }

////////////////////////////////////////////////////////////////////////////
struct Interval {
    float lo, hi;
    Interval() : lo(0), hi(0) {}
    Interval(float l, float h) : lo(l), hi(h) {}
    Interval(float f) : lo(f), hi(f) {}
    bool has_nan() const { return lo != lo || hi != hi; }
    bool contains(float v) const { return v >= lo && v <= hi; }
    float width() const { return hi - lo; }
    static Interval nan() { return Interval(NAN, NAN); }
};
struct IC { Interval v; Choice c; };

static const float PI_F = 3.14159265358979323846f;
static const float TAU_F = 6.28318530717958647692f;

static inline Interval i_abs(Interval a) {
    if (a.lo < 0.0f) {
        if (a.hi > 0.0f) return Interval(0.0f, rmax(a.hi, -a.lo));
        return Interval(-a.hi, -a.lo);
    }
    return a;
}
static inline Interval i_square(Interval a) {
    if (a.hi < 0.0f) return Interval(a.hi * a.hi, a.lo * a.lo);
    if (a.lo > 0.0f) return Interval(a.lo * a.lo, a.hi * a.hi);
    if (a.has_nan()) return Interval::nan();
    float m = rmax(std::fabs(a.lo), std::fabs(a.hi));
    return Interval(0.0f, m * m);
}
static inline int i_quadrant(float angle) {
    return int(uint8_t(rem_euclid(std::floor(angle * 2.0f / PI_F), 4.0f)));
}
static inline Interval i_compare(Interval l, Interval r) {
    if (l.has_nan() || r.has_nan()) return Interval::nan();
    if (l.hi < r.lo) return Interval(-1.0f);
    if (l.lo > r.hi) return Interval(1.0f);
    if (l.lo == l.hi && r.lo == r.hi && l.lo == r.lo) return Interval(0.0f, 0.0f);
    return Interval(-1.0f, 1.0f);
}
static inline Interval i_sin(Interval a) {
    if (a.has_nan()) return Interval::nan();
    if (a.width() >= TAU_F) return Interval(-1.0f, 1.0f);
    if (a.lo == a.hi) return Interval(std::sin(a.lo));
    int ql = i_quadrant(a.lo), qu = i_quadrant(a.hi);
    float d = a.width();
    float sl = std::sin(a.lo), su = std::sin(a.hi);
    if (ql == qu && d >= PI_F) return Interval(-1.0f, 1.0f);
    if ((ql == 1 && qu == 1) || (ql == 2 && qu == 2)) return Interval(su, sl);
    if ((ql == 0 && qu == 0) || (ql == 3 && qu == 3)) return Interval(sl, su);
    if (ql == 3 && qu == 0) return d >= PI_F ? Interval(-1.0f, 1.0f) : Interval(sl, su);
    if (ql == 1 && qu == 2) return d >= PI_F ? Interval(-1.0f, 1.0f) : Interval(su, sl);
    if ((ql == 0 || ql == 3) && (qu == 1 || qu == 2)) return Interval(rmin(sl, su), 1.0f);
    if ((ql == 1 || ql == 2) && (qu == 3 || qu == 0)) return Interval(-1.0f, rmax(sl, su));
    return Interval(-1.0f, 1.0f);  // (Q0,Q3) | (Q2,Q1)
}
static inline Interval i_cos(Interval a) {
    if (a.has_nan()) return Interval::nan();
    if (a.width() >= TAU_F) return Interval(-1.0f, 1.0f);
    if (a.lo == a.hi) return Interval(std::cos(a.lo));
    int ql = i_quadrant(a.lo), qu = i_quadrant(a.hi);
    float d = a.width();
    float cl = std::cos(a.lo), cu = std::cos(a.hi);
    if (ql == qu && d >= PI_F) return Interval(-1.0f, 1.0f);
    if ((ql == 2 && qu == 2) || (ql == 3 && qu == 3)) return Interval(cl, cu);
    if ((ql == 0 && qu == 0) || (ql == 1 && qu == 1)) return Interval(cu, cl);
    if (ql == 2 && qu == 3) return d >= PI_F ? Interval(-1.0f, 1.0f) : Interval(cl, cu);
    if (ql == 0 && qu == 1) return d >= PI_F ? Interval(-1.0f, 1.0f) : Interval(cu, cl);
    if ((ql == 2 || ql == 3) && (qu == 0 || qu == 1)) return Interval(rmin(cl, cu), 1.0f);
    if ((ql == 0 || ql == 1) && (qu == 2 || qu == 3)) return Interval(-1.0f, rmax(cl, cu));
    return Interval(-1.0f, 1.0f);  // (Q3,Q2) | (Q1,Q0)
}
static inline Interval i_tan(Interval a) {
    float size = a.hi - a.lo;
    if (size >= PI_F) return Interval::nan();
    if (a.lo == a.hi) return Interval(std::tan(a.lo));
    float l = std::tan(a.lo), u = std::tan(a.hi);
    return u >= l ? Interval(l, u) : Interval::nan();
}
static inline Interval i_asin(Interval a) {
    if (a.lo < -1.0f || a.hi > 1.0f) return Interval::nan();
    if (a.lo == a.hi) return Interval(std::asin(a.lo));
    return Interval(std::asin(a.lo), std::asin(a.hi));
}
static inline Interval i_acos(Interval a) {
    if (a.lo < -1.0f || a.hi > 1.0f) return Interval::nan();
    if (a.lo == a.hi) return Interval(std::acos(a.lo));
    return Interval(std::acos(a.hi), std::acos(a.lo));
}
static inline Interval i_atan(Interval a) { return Interval(std::atan(a.lo), std::atan(a.hi)); }
static inline Interval i_exp(Interval a) { return Interval(std::exp(a.lo), std::exp(a.hi)); }
static inline Interval i_ln(Interval a) {
    if (a.lo <= 0.0f) return Interval::nan();
    return Interval(std::log(a.lo), std::log(a.hi));
}
static inline Interval i_sqrt(Interval a) {
    if (a.lo < 0.0f) return Interval::nan();
    return Interval(std::sqrt(a.lo), std::sqrt(a.hi));
}
static inline Interval i_recip(Interval a) {
    if (a.lo > 0.0f || a.hi < 0.0f) return Interval(1.0f / a.hi, 1.0f / a.lo);
    return Interval::nan();
}
static inline IC i_min_choice(Interval a, Interval b) {
    if (a.has_nan() || b.has_nan()) return {Interval::nan(), C_BOTH};
    Choice c = a.hi < b.lo ? C_LEFT : (b.hi < a.lo ? C_RIGHT : C_BOTH);
    return {Interval(rmin(a.lo, b.lo), rmin(a.hi, b.hi)), c};
}
static inline IC i_max_choice(Interval a, Interval b) {
    if (a.has_nan() || b.has_nan()) return {Interval::nan(), C_BOTH};
    Choice c = a.lo > b.hi ? C_LEFT : (b.lo > a.hi ? C_RIGHT : C_BOTH);
    return {Interval(rmax(a.lo, b.lo), rmax(a.hi, b.hi)), c};
}
static inline IC i_and_choice(Interval a, Interval b) {
    if (a.has_nan() || b.has_nan()) return {Interval::nan(), C_BOTH};
    if (a.lo == 0.0f && a.hi == 0.0f) return {Interval(0.0f), C_LEFT};
    if (!a.contains(0.0f)) return {b, C_RIGHT};
    return {Interval(rmin(b.lo, 0.0f), rmax(b.hi, 0.0f)), C_BOTH};
}
static inline IC i_or_choice(Interval a, Interval b) {
    if (a.has_nan() || b.has_nan()) return {Interval::nan(), C_BOTH};
    if (!a.contains(0.0f)) return {a, C_LEFT};
    if (a.lo == 0.0f && a.hi == 0.0f) return {b, C_RIGHT};
    return {Interval(rmin(a.lo, b.lo), rmax(a.hi, b.hi)), C_BOTH};
}
static inline Interval i_rem_euclid(Interval a, Interval o) {
    if (a.has_nan() || o.has_nan() || o.contains(0.0f)) return Interval::nan();
    if (o.lo == o.hi && o.lo > 0.0f) {
        float x = a.lo / o.lo, y = a.hi / o.lo;
        if (x != std::floor(x) && std::floor(x) == std::floor(y))
            return Interval(rem_euclid(a.lo, o.lo), rem_euclid(a.hi, o.lo));
        return Interval(0.0f, i_abs(o).hi);
    }
    return Interval(0.0f, i_abs(o).hi);
}
static inline Interval i_floor(Interval a) { return Interval(std::floor(a.lo), std::floor(a.hi)); }
static inline Interval i_ceil(Interval a) { return Interval(std::ceil(a.lo), std::ceil(a.hi)); }
static inline Interval i_round(Interval a) { return Interval(std::round(a.lo), std::round(a.hi)); }
static inline Interval i_not(Interval a) {
    if (!a.contains(0.0f) && !a.has_nan()) return Interval(0.0f, 0.0f);
    if (a.lo == 0.0f && a.hi == 0.0f) return Interval(1.0f, 1.0f);
    return Interval(0.0f, 1.0f);
}
static inline Interval i_atan2(Interval y, Interval x) {
    if (y.has_nan() || x.has_nan()) return Interval::nan();
    if (y.lo <= 0.0f && y.hi >= 0.0f && x.lo < 0.0f) return Interval(-PI_F, PI_F);
    float lower = INFINITY, upper = -INFINITY;
    auto update = [&](float yy, float xx) {
        float v = std::atan2(yy, xx);
        lower = rmin(lower, v);
        upper = rmax(upper, v);
    };
    if (y.lo >= 0.0f) {
        if (x.lo >= 0.0f) { update(y.hi, x.lo); update(y.lo, x.hi); }
        else if (x.hi <= 0.0f) { update(y.lo, x.lo); update(y.hi, x.hi); }
        else { update(y.lo, x.lo); update(y.lo, x.hi); }
    } else if (y.hi <= 0.0f) {
        if (x.lo >= 0.0f) { update(y.lo, x.lo); update(y.hi, x.hi); }
        else if (x.hi <= 0.0f) { update(y.hi, x.lo); update(y.lo, x.hi); }
        else { update(y.hi, x.lo); update(y.hi, x.hi); }
    } else {
        update(y.lo, x.lo);
        update(y.hi, x.lo);
    }
    return Interval(lower, upper);
}
static inline Interval i_mix(Interval a, Interval b) {
    if (a.has_nan() || b.has_nan() || f2u(a.lo) != f2u(a.hi) || f2u(b.lo) != f2u(b.hi))
        return Interval::nan();
    return Interval(u2f(rng_mix(f2u(a.lo), f2u(b.lo))));
}
static inline Interval i_rand(Interval a) {
    if (a.has_nan() || f2u(a.lo) != f2u(a.hi)) return Interval(0.0f, 1.0f);
    return Interval(rng_rand(f2u(a.lo)));
}
static inline Interval i_add(Interval a, Interval b) { return Interval(a.lo + b.lo, a.hi + b.hi); }
static inline Interval i_sub(Interval a, Interval b) { return Interval(a.lo - b.hi, a.hi - b.lo); }
static inline Interval i_neg(Interval a) { return Interval(-a.hi, -a.lo); }
static inline Interval i_mul(Interval a, Interval b) {
    if (a.has_nan() || b.has_nan()) return Interval::nan();
    float o0 = a.lo * b.lo, o1 = a.lo * b.hi, o2 = a.hi * b.lo, o3 = a.hi * b.hi;
    float lo = o0, hi = o0;
    lo = rmin(lo, o1); hi = rmax(hi, o1);
    lo = rmin(lo, o2); hi = rmax(hi, o2);
    lo = rmin(lo, o3); hi = rmax(hi, o3);
    return Interval(lo, hi);
}
static inline Interval i_mul_f(Interval a, float k) {
    if (a.has_nan() || k != k) return Interval::nan();
    if (k < 0.0f) return Interval(a.hi * k, a.lo * k);
    return Interval(a.lo * k, a.hi * k);
}
static inline Interval i_div(Interval a, Interval b) {
    if (a.has_nan()) return Interval::nan();
    if (b.lo > 0.0f || b.hi < 0.0f) {
        float o0 = a.lo / b.lo, o1 = a.lo / b.hi, o2 = a.hi / b.lo, o3 = a.hi / b.hi;
        float lo = o0, hi = o0;
        lo = rmin(lo, o1); hi = rmax(hi, o1);
        lo = rmin(lo, o2); hi = rmax(hi, o2);
        lo = rmin(lo, o3); hi = rmax(hi, o3);
        return Interval(lo, hi);
    }
    return Interval::nan();
}

////////////////////////////////////////////////////////////////////////////
struct Grad {
    float v, dx, dy, dz;
    Grad() : v(0), dx(0), dy(0), dz(0) {}
    Grad(float v_) : v(v_), dx(0), dy(0), dz(0) {}
    Grad(float v_, float x, float y, float z) : v(v_), dx(x), dy(y), dz(z) {}
};
static inline Grad g_add(Grad a, Grad b) { return Grad(a.v + b.v, a.dx + b.dx, a.dy + b.dy, a.dz + b.dz); }
static inline Grad g_sub(Grad a, Grad b) { return Grad(a.v - b.v, a.dx - b.dx, a.dy - b.dy, a.dz - b.dz); }
static inline Grad g_neg(Grad a) { return Grad(-a.v, -a.dx, -a.dy, -a.dz); }
static inline Grad g_mul(Grad a, Grad b) {
    return Grad(a.v * b.v, a.v * b.dx + b.v * a.dx, a.v * b.dy + b.v * a.dy, a.v * b.dz + b.v * a.dz);
}
static inline Grad g_mul_f(Grad a, float k) { return Grad(a.v * k, a.dx * k, a.dy * k, a.dz * k); }
static inline Grad g_div(Grad a, Grad b) {
    float d = b.v * b.v;
    return Grad(a.v / b.v, (b.v * a.dx - a.v * b.dx) / d, (b.v * a.dy - a.v * b.dy) / d,
                (b.v * a.dz - a.v * b.dz) / d);
}
static inline Grad g_abs(Grad a) { return a.v < 0.0f ? g_neg(a) : a; }
static inline Grad g_sqrt(Grad a) {
    float v = std::sqrt(a.v);
    return Grad(v, a.dx / (2.0f * v), a.dy / (2.0f * v), a.dz / (2.0f * v));
}
static inline Grad g_sin(Grad a) {
    float c = std::cos(a.v);
    return Grad(std::sin(a.v), a.dx * c, a.dy * c, a.dz * c);
}
static inline Grad g_cos(Grad a) {
    float s = -std::sin(a.v);
    return Grad(std::cos(a.v), a.dx * s, a.dy * s, a.dz * s);
}
static inline Grad g_tan(Grad a) {
    float c0 = std::cos(a.v);
    float c = c0 * c0;
    return Grad(std::tan(a.v), a.dx / c, a.dy / c, a.dz / c);
}
static inline Grad g_asin(Grad a) {
    float r = std::sqrt(1.0f - a.v * a.v);
    return Grad(std::asin(a.v), a.dx / r, a.dy / r, a.dz / r);
}
static inline Grad g_acos(Grad a) {
    float r = std::sqrt(1.0f - a.v * a.v);
    return Grad(std::acos(a.v), -a.dx / r, -a.dy / r, -a.dz / r);
}
static inline Grad g_atan(Grad a) {
    float r = a.v * a.v + 1.0f;
    return Grad(std::atan(a.v), a.dx / r, a.dy / r, a.dz / r);
}
static inline Grad g_exp(Grad a) {
    float v = std::exp(a.v);
    return Grad(v, v * a.dx, v * a.dy, v * a.dz);
}
static inline Grad g_ln(Grad a) { return Grad(std::log(a.v), a.dx / a.v, a.dy / a.v, a.dz / a.v); }
static inline Grad g_min(Grad a, Grad b) {
    if (a.v != a.v || b.v != b.v) return Grad(NAN);
    return a.v < b.v ? a : b;
}
static inline Grad g_max(Grad a, Grad b) {
    if (a.v != a.v || b.v != b.v) return Grad(NAN);
    return a.v > b.v ? a : b;
}
static inline Grad g_rem_euclid(Grad a, Grad b) {
    float e = div_euclid(a.v, b.v);
    return Grad(rem_euclid(a.v, b.v), a.dx - b.dx * e, a.dy - b.dy * e, a.dz - b.dz * e);
}
static inline Grad g_and(Grad a, Grad b) { return a.v == 0.0f ? a : b; }
static inline Grad g_or(Grad a, Grad b) { return a.v != 0.0f ? a : b; }
static inline Grad g_atan2(Grad y, Grad x) {
    float d = x.v * x.v + y.v * y.v;
    return Grad(std::atan2(y.v, x.v), (x.v * y.dx - y.v * x.dx) / d, (x.v * y.dy - y.v * x.dy) / d,
                (x.v * y.dz - y.v * x.dz) / d);
}
static inline Grad g_compare(Grad a, Grad b) { return Grad(f_compare(a.v, b.v)); }

}  // namespace oracle
