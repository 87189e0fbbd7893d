// sm_100a kernels for the tape-evaluation hot path.
//
//  k_interval_level  -- K1: one warp per parent tile, one lane per child tile;
//                       walks the (warp-uniform) parent tape, classifies each
//                       child (fill inside / fill outside / ambiguous), then
//                       runs the reverse liveness pass that compacts a child
//                       tape into the arena (VmData::simplify semantics,
//                       fidget-core/src/vm/data.rs:123-318) and queues the
//                       ambiguous children for the next level
//                       (pixel.rs:316-398 / voxel.rs:275-357).
//  k_pixels_2d       -- K2: one warp per leaf tile, two pixels per lane
//                       (pixel.rs:400-440 + VmFloatSliceEval, vm/mod.rs:800).
//  k_fill_2d         -- paints interval-proven tiles (pixel.rs:345-369).
//  k_float_slice / k_grad_slice / k_interval_batch / k_point_batch /
//  k_simplify_single -- the trait-level evaluators behind fc_*_eval.
//
// All interpreters keep the tape's VM registers in per-thread local memory
// (L1-resident, lane-interleaved, so a warp's access to one register is one
// 128/256-byte line) and read tape clauses with warp-uniform 8-byte loads.
#include <algorithm>
#include <cstdio>

#include "kernels.cuh"

namespace fdev {

#define FULL 0xffffffffu

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// INPUT clause -> value: the axes get coordinates, other slots their bound value
template <class T, class F>
__device__ __forceinline__ T pick_input(const VarBind& vb, uint32_t i, T X, T Y, T Z, F from_float) {
    const int k = int(i);
    if (k == vb.x) return X;
    if (k == vb.y) return Y;
    if (k == vb.z) return Z;
    return from_float(vb.values[k & (MAX_RENDER_VARS - 1)]);
}

struct Dec {
    uint32_t op, form, out, lhs, rhs;
    Dec() = default;
    __device__ __forceinline__ explicit Dec(uint32_t x) {
        uint32_t dop = x & 0xffu;
        op = dop >> 2;
        form = dop & 3u;
        out = (x >> 8) & 0xffu;
        lhs = (x >> 16) & 0xffu;
        rhs = x >> 24;
    }
};

// ---------------------------------------------------------------------------
// Interval interpreter.  `Input` maps a variable index to an interval,
// `Sink` receives one choice per choice clause in evaluation order, `Out`
// receives (output index, value).
template <class Input, class Sink, class Out>
__device__ __forceinline__ void run_interval(const uint2* __restrict__ tape, uint32_t n_ops, itv* slots,
                                             Input input, Sink& sink, Out out_fn) {
    if (n_ops == 0) return;
    uint2 w = __ldg(tape);
    for (uint32_t i = 0; i < n_ops; ++i) {
        uint2 nxt = __ldg(tape + (i + 1 < n_ops ? i + 1 : i));
        Dec d(w.x);
        float imm = __uint_as_float(w.y);
        itv sl = slots[d.lhs], sr = slots[d.rhs];
        itv a = d.form == F_IR ? iv1(imm) : sl;
        itv b = d.form == F_RI ? iv1(imm) : sr;
        itv r;
        if (d.op >= OP_MIN) {
            if (d.op == OP_MEM) {
                if (d.form == F_RI) slots[d.out] = slots[MEM_BASE + w.y];
                else slots[MEM_BASE + w.y] = sl;
                w = nxt;
                continue;
            }
            uint32_t c;
            r = iv_choice_op(d.op, a, b, c);
            sink.push(c);
        } else if (d.op >= OP_ADD) {
            if (d.op == OP_MUL && d.form == F_RI) r = iv_mul_f(sl, imm);
            else r = iv_binary(d.op, a, b);
        } else if (d.op >= OP_NEG) {
            r = iv_unary(d.op, sl);
        } else if (d.op == OP_COPY) {
            r = d.form == F_RI ? iv1(imm) : sl;
        } else if (d.op == OP_INPUT) {
            r = input(w.y);
        } else {  // OP_OUTPUT
            out_fn(w.y, sl);
            w = nxt;
            continue;
        }
        slots[d.out] = r;
        w = nxt;
    }
}

// Two-points-per-lane f32 interpreter
__device__ __forceinline__ float2 f32x2_unary(uint32_t op, float2 a) {
    switch (op) {
        case OP_NEG: return make_float2(-a.x, -a.y);
        case OP_ABS: return make_float2(fabsf(a.x), fabsf(a.y));
        case OP_SQRT: return make_float2(sqrtf(a.x), sqrtf(a.y));
        case OP_SQUARE: return make_float2(a.x * a.x, a.y * a.y);
        default: return make_float2(f32_unary(op, a.x), f32_unary(op, a.y));
    }
}
__device__ __forceinline__ float2 f32x2_binary(uint32_t op, float2 a, float2 b) {
    switch (op) {
        case OP_ADD: return make_float2(a.x + b.x, a.y + b.y);
        case OP_SUB: return make_float2(a.x - b.x, a.y - b.y);
        case OP_MUL: return make_float2(a.x * b.x, a.y * b.y);
        case OP_MIN: return make_float2(f_min(a.x, b.x), f_min(a.y, b.y));
        case OP_MAX: return make_float2(f_max(a.x, b.x), f_max(a.y, b.y));
        default: return make_float2(f32_binary(op, a.x, b.x), f32_binary(op, a.y, b.y));
    }
}

template <class Input>
__device__ __forceinline__ float2 run_f32x2(const uint2* __restrict__ tape, uint32_t n_ops, float2* slots,
                                            Input input) {
    float2 result = make_float2(nanf_(), nanf_());
    if (n_ops == 0) return result;
    uint2 w = __ldg(tape);
    for (uint32_t i = 0; i < n_ops; ++i) {
        uint2 nxt = __ldg(tape + (i + 1 < n_ops ? i + 1 : i));
        Dec d(w.x);
        float imm = __uint_as_float(w.y);
        float2 sl = slots[d.lhs], sr = slots[d.rhs];
        float2 a = d.form == F_IR ? make_float2(imm, imm) : sl;
        float2 b = d.form == F_RI ? make_float2(imm, imm) : sr;
        float2 r;
        if (d.op >= OP_ADD) {
            if (d.op == OP_MEM) {
                if (d.form == F_RI) slots[d.out] = slots[MEM_BASE + w.y];
                else slots[MEM_BASE + w.y] = sl;
                w = nxt;
                continue;
            }
            r = f32x2_binary(d.op, a, b);
        } else if (d.op >= OP_NEG) {
            r = f32x2_unary(d.op, sl);
        } else if (d.op == OP_COPY) {
            r = d.form == F_RI ? make_float2(imm, imm) : sl;
        } else if (d.op == OP_INPUT) {
            r = input(w.y);
        } else {
            if (w.y == 0) result = sl;
            w = nxt;
            continue;
        }
        slots[d.out] = r;
        w = nxt;
    }
    return result;
}

// ---------------------------------------------------------------------------
// Choice storage for the level kernel: 2 bits per choice, 16 per word, words
// interleaved across the 32 lanes of the warp.
struct ChoicePacker {
    uint32_t* base;  // already offset by lane; stride 32
    uint32_t acc = 0, ci = 0;
    bool any_nonboth = false;
    __device__ __forceinline__ void push(uint32_t c) {
        acc |= c << ((ci & 15u) * 2u);
        any_nonboth |= (c != 3u);
        ++ci;
        if ((ci & 15u) == 0u) {
            base[((ci >> 4) - 1u) * 32u] = acc;
            acc = 0;
        }
    }
    __device__ __forceinline__ void finish() {
        if (ci & 15u) base[(ci >> 4) * 32u] = acc;
    }
};
struct ChoiceUnpacker {
    const uint32_t* base;
    uint32_t ci;       // choices remaining
    uint32_t cached_word = 0xffffffffu, cur = 0;
    __device__ __forceinline__ uint32_t pop() {
        --ci;
        uint32_t wi = ci >> 4;
        if (wi != cached_word) {
            cur = base[wi * 32u];
            cached_word = wi;
        }
        return (cur >> ((ci & 15u) * 2u)) & 3u;
    }
};
struct ByteChoiceSource {
    const uint8_t* base;
    uint32_t ci;
    __device__ __forceinline__ uint32_t pop() { return base[--ci] & 3u; }
};

// Reverse liveness pass + compaction (VmData::simplify on a register tape that
// keeps the parent's register assignment).  Writes the child tape backwards,
// ending at `wend`.  `live` is this warp's [8][32] bitset in shared memory.
template <class ChoiceSrc>
__device__ __forceinline__ void simplify_lane(const uint2* __restrict__ tape, uint32_t n_ops, bool active,
                                              uint32_t (*live)[32], int lane, ChoiceSrc& cs, uint2* wend,
                                              uint32_t& n_dev, uint32_t& ref_len, uint32_t& n_choices) {
#pragma unroll
    for (int k = 0; k < 8; ++k) live[k][lane] = 0;
    auto test = [&](uint32_t r) { return (live[r >> 5][lane] >> (r & 31u)) & 1u; };
    auto set = [&](uint32_t r) { live[r >> 5][lane] |= 1u << (r & 31u); };
    auto clear = [&](uint32_t r) { live[r >> 5][lane] &= ~(1u << (r & 31u)); };
    uint2* wp = wend;
    uint32_t ref = 0, nch = 0;
    uint2 nxt = n_ops ? __ldg(tape + (n_ops - 1)) : make_uint2(0, 0);
    for (int i = int(n_ops) - 1; i >= 0; --i) {
        const uint2 w = nxt;
        if (i > 0) nxt = __ldg(tape + (i - 1));   // prefetch: the clause stream is the latency chain here
        Dec d(w.x);
        uint32_t c = 3u;
        bool is_choice = op_is_choice(d.op);
        if (is_choice) c = cs.pop();
        if (!active) continue;
        if (d.op == OP_OUTPUT) {
            set(d.lhs);
            *--wp = w;
            ++ref;
            continue;
        }
        if (!test(d.out)) continue;
        clear(d.out);
        if (is_choice && c != 3u) {
            if (c == 2u && d.form == F_RI) {
                *--wp = make_uint2(enc(OP_COPY, F_RI, d.out, 0xff, 0xff), w.y);
                ++ref;
            } else {
                // F_RI keeps its register in lhs; F_RR left = lhs, right = rhs
                uint32_t src = (c == 1u) ? d.lhs : d.rhs;
                if (src == d.out) {
                    set(d.out);
                } else {
                    uint32_t was = test(src);
                    set(src);
                    *--wp = make_uint2(enc(OP_COPY, was ? F_RR : F_ALIAS, d.out, src, 0xff), 0xFF000000u);
                    ref += was;
                }
            }
            continue;
        }
        if (d.op == OP_COPY && d.form != F_RI) {
            uint32_t src = d.lhs;
            if (src == d.out) { set(d.out); continue; }
            uint32_t was = test(src);
            set(src);
            uint32_t nf = (d.form == F_ALIAS || !was) ? F_ALIAS : F_RR;
            *--wp = make_uint2(enc(OP_COPY, nf, d.out, src, 0xff), w.y);
            ref += (nf == F_RR);
            continue;
        }
        *--wp = w;
        ++ref;
        if (is_choice) ++nch;
        if (d.op == OP_INPUT || d.op == OP_COPY) continue;
        if (d.op < OP_ADD) set(d.lhs);
        else {
            if (d.form != F_IR) set(d.lhs);
            if (d.form != F_RI) set(d.rhs);
        }
    }
    n_dev = uint32_t(wend - wp);
    ref_len = ref;
    n_choices = nch;
}

// ---------------------------------------------------------------------------
// K1: interval level kernel.  DIM = 2: pixel::render tiles (fill records are
// painted later by k_fill_2d).  DIM = 3: voxel::render tiles; an
// interval-proven-inside tile raises the heightmap to its top + 1
// (voxel.rs:310-317), heightmap entries are (depth << 32 | leaf job id + 1).
template <int DIM>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_interval_level(const __grid_constant__ LevelParams p) {
    __shared__ uint32_t live_s[WARPS_PER_BLOCK][8][32];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * WARPS_PER_BLOCK + wib;
    uint32_t* cs = p.choice_scratch + size_t(gw) * p.choice_words * 32u + lane;
    itv slots[REG_SLOTS];

    const uint32_t n_roots = p.roots_x * p.roots_y * (DIM == 3 ? p.roots_z : 1u);
    const uint32_t n_jobs = p.root_mode ? (n_roots + 31u) / 32u : min(p.ctr->n_jobs[p.level], p.cap_in);
    const uint32_t T = p.tile;

    for (;;) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(&p.ctr->cursor[p.level], 1u);
        j = __shfl_sync(FULL, j, 0);
        if (j >= n_jobs) break;

        TapeRef tr;
        uint32_t px = 0, py = 0, pz = 0, nchild;
        if (p.root_mode) {
            tr = p.root_tape;
            nchild = min(32u, n_roots - j * 32u);
        } else {
            const TileJob* job = p.jobs_in + j;
            px = job->x;
            py = job->y;
            pz = job->z;
            tr = job->tape;
            nchild = p.n_axis * p.n_axis * (DIM == 3 ? p.n_axis : 1u);
        }
        const uint2* tape = tr.ptr;

        for (uint32_t chunk = 0; chunk * 32u < nchild; ++chunk) {
            const uint32_t c = chunk * 32u + lane;
            const bool valid = c < nchild;
            uint32_t cx, cy, cz = 0;
            if (p.root_mode) {
                uint32_t idx = j * 32u + (valid ? c : 0u);
                cx = p.root_x0 + (idx % p.roots_x) * T;
                cy = p.root_y0 + ((idx / p.roots_x) % p.roots_y) * T;
                if (DIM == 3) cz = p.root_z0 + (idx / (p.roots_x * p.roots_y)) * T;
            } else {
                uint32_t cc = valid ? c : 0u;
                cx = px + (cc % p.n_axis) * T;
                cy = py + ((cc / p.n_axis) % p.n_axis) * T;
                if (DIM == 3) cz = pz + (cc / (p.n_axis * p.n_axis)) * T;
            }
            // Region in screen coordinates -> model space (pixel.rs:325-342, voxel.rs:291-306)
            itv X = iv(float(cx), float(cx) + float(T));
            itv Y = iv(float(cy), float(cy) + float(T));
            itv Z = DIM == 3 ? iv(float(cz), float(cz) + float(T)) : iv(p.z2d, p.z2d);
            itv vx, vy, vz;
            if (DIM == 3 && p.mode == 1u) {
                // octree cell bounds in world space (CellBounds::child, cell.rs:155-166): dyadic, exact in f32
                const float h = p.cell_h;
                X = iv(float(cx) * h - 1.0f, float(cx + T) * h - 1.0f);
                Y = iv(float(cy) * h - 1.0f, float(cy + T) * h - 1.0f);
                Z = iv(float(cz) * h - 1.0f, float(cz + T) * h - 1.0f);
                if (p.has_transform) xform_iv(p.mat, X, Y, Z, vx, vy, vz);
                else { vx = X; vy = Y; vz = Z; }
            } else {
                xform_iv(p.mat, X, Y, Z, vx, vy, vz);
            }

            ChoicePacker pk;
            pk.base = cs;
            itv r = iv_nan();
            run_interval(
                tape, tr.n_ops, slots,
                [&](uint32_t i) { return pick_input(p.vb, i, vx, vy, vz, [](float f) { return iv1(f); }); }, pk,
                [&](uint32_t oi, itv v) { if (oi == 0) r = v; });
            pk.finish();

            const bool fill_in = valid && !p.pixel_perfect && r.y < 0.0f;
            const bool fill_out = valid && !p.pixel_perfect && !fill_in && r.x > 0.0f;
            const bool amb = valid && !fill_in && !fill_out;

            if (DIM == 3) {
                // full tile: depth = max(depth, top + 1) over its footprint (voxel.rs:310-317)
                uint32_t m = p.mode == 1u ? 0u : __ballot_sync(FULL, fill_in);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t fx = __shfl_sync(FULL, cx, src), fy = __shfl_sync(FULL, cy, src),
                                   fz = __shfl_sync(FULL, cz, src);
                    const unsigned long long key = (unsigned long long)(fz + T + 1u) << 32;
                    for (uint32_t q = lane; q < T * T; q += 32u) {
                        const uint32_t x = fx + q % T, y = fy + q / T;
                        if (x < p.width && y < p.height) atomicMax(&p.heightmap[size_t(y) * p.width + x], key);
                    }
                }
                if (p.stats) {
                    uint32_t mv = __ballot_sync(FULL, valid), mi = __ballot_sync(FULL, fill_in),
                             mo = __ballot_sync(FULL, fill_out), ma = __ballot_sync(FULL, amb);
                    if (lane == 0) {
                        atomicAdd(&p.stats->evaluated[p.level], (unsigned long long)__popc(mv));
                        if (mi) atomicAdd(&p.stats->filled_inside[p.level], (unsigned long long)__popc(mi));
                        if (mo) atomicAdd(&p.stats->filled_outside[p.level], (unsigned long long)__popc(mo));
                        if (ma) atomicAdd(&p.stats->ambiguous[p.level], (unsigned long long)__popc(ma));
                    }
                }
            } else {
                uint32_t m = __ballot_sync(FULL, fill_in || fill_out);
                if (m) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&p.ctr->n_fills[p.level], uint32_t(__popc(m)));
                    base = __shfl_sync(FULL, base, 0);
                    if (fill_in || fill_out) {
                        uint32_t slot = base + __popc(m & lanemask_lt());
                        if (slot < p.cap_fills) {
                            FillRec fr;
                            fr.x = cx;
                            fr.y = cy;
                            fr.value = 0x7FC00000u | (uint32_t(p.level & 0xff) << 1) | (fill_in ? 1u : 0u) | (0xF6u << 9);
                            p.fills[slot] = fr;
                        } else {
                            atomicOr(&p.ctr->error, 2u);
                        }
                    }
                }
                if (p.stats) {
                    uint32_t mv = __ballot_sync(FULL, valid), mi = __ballot_sync(FULL, fill_in),
                             mo = __ballot_sync(FULL, fill_out), ma = __ballot_sync(FULL, amb);
                    if (lane == 0) {
                        atomicAdd(&p.stats->evaluated[p.level], (unsigned long long)__popc(mv));
                        if (mi) atomicAdd(&p.stats->filled_inside[p.level], (unsigned long long)__popc(mi));
                        if (mo) atomicAdd(&p.stats->filled_outside[p.level], (unsigned long long)__popc(mo));
                        if (ma) atomicAdd(&p.stats->ambiguous[p.level], (unsigned long long)__popc(ma));
                    }
                }
            }

            // simplification (render/mod.rs:96-152: keep the child only if it is shorter)
            TapeRef child = tr;
            const bool need = amb && pk.any_nonboth;
            const uint32_t mneed = __ballot_sync(FULL, need);
            if (mneed) {
                const uint32_t total = __popc(mneed);
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&p.ctr->arena_top, (unsigned long long)total * tr.n_ops);
                base = __shfl_sync(FULL, base, 0);
                if (base + (unsigned long long)total * tr.n_ops > p.arena_cap) {
                    if (lane == 0) atomicOr(&p.ctr->error, 1u);
                } else {
                    const uint32_t rank = __popc(mneed & lanemask_lt());
                    unsigned long long end = base + (unsigned long long)(rank + 1u) * tr.n_ops;
                    ChoiceUnpacker cu;
                    cu.base = cs;
                    cu.ci = tr.n_choices;
                    uint32_t n_dev, ref_len, nch;
                    simplify_lane(tape, tr.n_ops, need, live_s[wib], lane, cu, p.arena + end, n_dev, ref_len, nch);
                    bool keep = need && ref_len < tr.ref_len;
                    if (keep) {
                        child.ptr = p.arena + (end - n_dev);
                        child.n_ops = n_dev;
                        child.ref_len = ref_len;
                        child.n_choices = nch;
                    }
                    if (p.stats) {
                        uint32_t mk = __ballot_sync(FULL, keep);
                        if (lane == 0 && mk) atomicAdd(&p.stats->simplified[p.level], (unsigned long long)__popc(mk));
                    }
                }
            }

            // queue ambiguous children for the next level
            const uint32_t mamb = __ballot_sync(FULL, amb);
            if (mamb) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&p.ctr->n_jobs[p.level + 1], uint32_t(__popc(mamb)));
                base = __shfl_sync(FULL, base, 0);
                if (amb) {
                    uint32_t slot = base + __popc(mamb & lanemask_lt());
                    if (slot < p.cap_out) {
                        TileJob o;
                        o.x = cx;
                        o.y = cy;
                        o.z = cz;
                        o.pad = 0;
                        o.tape = child;
                        p.jobs_out[slot] = o;
                    } else {
                        atomicOr(&p.ctr->error, 2u);
                    }
                }
            }
        }
    }
}

void launch_interval_level_2d(const LevelParams& p, int blocks, cudaStream_t s) {
    k_interval_level<2><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(p);
}
void launch_interval_level_3d(const LevelParams& p, int blocks, cudaStream_t s) {
    k_interval_level<3><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(p);
}

// Four-points-per-lane f32 interpreter (leaf voxels): decode, dispatch and register-file traffic
// are amortised over four points.
__device__ __forceinline__ float4 f32x4_unary(uint32_t op, float4 a) {
    switch (op) {
        case OP_NEG: return make_float4(-a.x, -a.y, -a.z, -a.w);
        case OP_ABS: return make_float4(fabsf(a.x), fabsf(a.y), fabsf(a.z), fabsf(a.w));
        case OP_SQRT: return make_float4(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z), sqrtf(a.w));
        case OP_SQUARE: return make_float4(a.x * a.x, a.y * a.y, a.z * a.z, a.w * a.w);
        default: return make_float4(f32_unary(op, a.x), f32_unary(op, a.y), f32_unary(op, a.z), f32_unary(op, a.w));
    }
}
__device__ __forceinline__ float4 f32x4_binary(uint32_t op, float4 a, float4 b) {
    switch (op) {
        case OP_ADD: return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        case OP_SUB: return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
        case OP_MUL: return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
        case OP_MIN: return make_float4(f_min(a.x, b.x), f_min(a.y, b.y), f_min(a.z, b.z), f_min(a.w, b.w));
        case OP_MAX: return make_float4(f_max(a.x, b.x), f_max(a.y, b.y), f_max(a.z, b.z), f_max(a.w, b.w));
        default: return make_float4(f32_binary(op, a.x, b.x), f32_binary(op, a.y, b.y), f32_binary(op, a.z, b.z),
                                    f32_binary(op, a.w, b.w));
    }
}
template <class Input>
__device__ __forceinline__ float4 run_f32x4(const uint2* __restrict__ tape, uint32_t n_ops, float4* slots, Input input) {
    float4 result = make_float4(nanf_(), nanf_(), nanf_(), nanf_());
    if (n_ops == 0) return result;
    uint2 w = __ldg(tape);
    for (uint32_t i = 0; i < n_ops; ++i) {
        const uint2 nxt = __ldg(tape + (i + 1 < n_ops ? i + 1 : i));
        Dec d(w.x);
        const float imm = __uint_as_float(w.y);
        const float4 sl = slots[d.lhs], sr = slots[d.rhs];
        const float4 im = make_float4(imm, imm, imm, imm);
        const float4 a = d.form == F_IR ? im : sl;
        const float4 b = d.form == F_RI ? im : sr;
        float4 r;
        if (d.op >= OP_ADD) {
            if (d.op == OP_MEM) {
                if (d.form == F_RI) slots[d.out] = slots[MEM_BASE + w.y];
                else slots[MEM_BASE + w.y] = sl;
                w = nxt;
                continue;
            }
            r = f32x4_binary(d.op, a, b);
        } else if (d.op >= OP_NEG) {
            r = f32x4_unary(d.op, sl);
        } else if (d.op == OP_COPY) {
            r = d.form == F_RI ? im : sl;
        } else if (d.op == OP_INPUT) {
            r = input(w.y);
        } else {
            if (w.y == 0) result = sl;
            w = nxt;
            continue;
        }
        slots[d.out] = r;
        w = nxt;
    }
    return result;
}

// ---------------------------------------------------------------------------
// K2 (3D): leaf voxels.  One warp per leaf tile; each lane owns two XY columns
// and walks Z front to back (k descending), two points per tape pass; the
// warp stops as soon as every column has hit the surface (voxel.rs:359-447).
__global__ void __launch_bounds__(128) k_voxels_3d(const __grid_constant__ VoxelParams p) {
    const int lane = threadIdx.x & 31;
    float4 slots[REG_SLOTS];
    const uint32_t n_jobs = min(p.ctr->n_jobs[p.list], p.cap_jobs);
    const uint32_t T = p.tile, ncol = T * T;
    unsigned long long shaded = 0;
    for (;;) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(&p.ctr->cursor[p.cursor], 1u);
        j = __shfl_sync(FULL, j, 0);
        if (j >= n_jobs) break;
        if (p.order) j = p.order[j];   // front-to-back: tiles behind a finished column find it done and exit early
        const TileJob* job = p.jobs + j;
        const uint32_t cx = job->x, cy = job->y, cz = job->z;
        const TapeRef tr = job->tape;
        const uint2* tape = tr.ptr;
        const unsigned long long id = (unsigned long long)(j + 1u);
        for (uint32_t base = 0; base < ncol; base += 64u) {
            const uint32_t c0 = base + lane, c1 = c0 + 32u;
            const bool v0 = c0 < ncol, v1 = c1 < ncol;
            const uint32_t i0 = (v0 ? c0 : 0u) % T, j0 = (v0 ? c0 : 0u) / T;
            const uint32_t i1 = (v1 ? c1 : 0u) % T, j1 = (v1 ? c1 : 0u) / T;
            const uint32_t gx0 = cx + i0, gy0 = cy + j0, gx1 = cx + i1, gy1 = cy + j1;
            const bool in0 = v0 && gx0 < p.width && gy0 < p.height, in1 = v1 && gx1 < p.width && gy1 < p.height;
            // columns already at or above this tile's top are skipped (voxel.rs:376-381)
            const uint32_t zmax = cz + T;
            bool done0 = !in0 || uint32_t(p.heightmap[size_t(gy0) * p.width + gx0] >> 32) >= zmax;
            bool done1 = !in1 || uint32_t(p.heightmap[size_t(gy1) * p.width + gx1] >> 32) >= zmax;
            // two Z levels per tape pass: (column 0, column 1) x (k, k - 1)
            for (int k = int(T) - 1; k >= 0; k -= 2) {
                if (__all_sync(FULL, done0 && done1)) break;
                const int k2 = k > 0 ? k - 1 : 0;
                float xa, ya, za, xb, yb, zb, xc, yc, zc, xd, yd, zd;
                xform_f32(p.mat, float(gx0), float(gy0), float(cz + uint32_t(k)), xa, ya, za);
                xform_f32(p.mat, float(gx1), float(gy1), float(cz + uint32_t(k)), xb, yb, zb);
                xform_f32(p.mat, float(gx0), float(gy0), float(cz + uint32_t(k2)), xc, yc, zc);
                xform_f32(p.mat, float(gx1), float(gy1), float(cz + uint32_t(k2)), xd, yd, zd);
                const float4 X = make_float4(xa, xb, xc, xd), Y = make_float4(ya, yb, yc, yd), Z = make_float4(za, zb, zc, zd);
                const float4 r = run_f32x4(tape, tr.n_ops, slots, [&](uint32_t i) {
                    return pick_input(p.vb, i, X, Y, Z, [](float f) { return make_float4(f, f, f, f); });
                });
                const unsigned long long key_hi = ((unsigned long long)(cz + uint32_t(k) + 1u) << 32) | id;
                const unsigned long long key_lo = ((unsigned long long)(cz + uint32_t(k2) + 1u) << 32) | id;
                const bool two = k > 0;
                if (!done0) {
                    shaded += two ? 2 : 1;
                    if (r.x < 0.0f) { atomicMax(&p.heightmap[size_t(gy0) * p.width + gx0], key_hi); done0 = true; }
                    else if (two && r.z < 0.0f) { atomicMax(&p.heightmap[size_t(gy0) * p.width + gx0], key_lo); done0 = true; }
                }
                if (!done1) {
                    shaded += two ? 2 : 1;
                    if (r.y < 0.0f) { atomicMax(&p.heightmap[size_t(gy1) * p.width + gx1], key_hi); done1 = true; }
                    else if (two && r.w < 0.0f) { atomicMax(&p.heightmap[size_t(gy1) * p.width + gx1], key_lo); done1 = true; }
                }
            }
        }
    }
    if (p.stats) {
        for (int o = 16; o > 0; o >>= 1) shaded += __shfl_xor_sync(FULL, shaded, o);
        if (lane == 0 && shaded) atomicAdd(&p.stats->pixels, shaded);
    }
}
void launch_voxels_3d(const VoxelParams& p, int blocks, cudaStream_t s) { k_voxels_3d<<<blocks, 128, 0, s>>>(p); }

// Front-to-back ordering of the leaf tiles (the reference walks Z descending, voxel.rs:244-263,
// 335-351): a counting sort by Z layer, front layer first.
__global__ void k_zsort_hist(const TileJob* jobs, const uint32_t* n_jobs, uint32_t cap, uint32_t z0, uint32_t tile,
                             uint32_t n_layers, uint32_t* hist) {
    const uint32_t n = min(*n_jobs, cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t layer = min((jobs[i].z - z0) / tile, n_layers - 1u);
        atomicAdd(&hist[n_layers - 1u - layer], 1u);
    }
}
__global__ void k_zsort_scan(uint32_t n_layers, uint32_t* hist) {   // single thread: n_layers <= 4096
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n_layers; ++i) {
        const uint32_t c = hist[i];
        hist[i] = acc;
        acc += c;
    }
}
__global__ void k_zsort_scatter(const TileJob* jobs, const uint32_t* n_jobs, uint32_t cap, uint32_t z0, uint32_t tile,
                                uint32_t n_layers, uint32_t* hist, uint32_t* order) {
    const uint32_t n = min(*n_jobs, cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t layer = min((jobs[i].z - z0) / tile, n_layers - 1u);
        order[atomicAdd(&hist[n_layers - 1u - layer], 1u)] = i;
    }
}
void launch_leaf_zsort(const TileJob* jobs, const uint32_t* n_jobs, uint32_t cap, uint32_t z0, uint32_t tile,
                       uint32_t n_layers, uint32_t* hist, uint32_t* order, cudaStream_t s) {
    cudaMemsetAsync(hist, 0, size_t(n_layers) * 4, s);
    k_zsort_hist<<<296, 256, 0, s>>>(jobs, n_jobs, cap, z0, tile, n_layers, hist);
    k_zsort_scan<<<1, 1, 0, s>>>(n_layers, hist);
    k_zsort_scatter<<<296, 256, 0, s>>>(jobs, n_jobs, cap, z0, tile, n_layers, hist, order);
}

// Gradient interpreter (VmGradSliceEval, vm/mod.rs:1097-1396)
template <class Input>
__device__ __forceinline__ grd run_grad(const uint2* __restrict__ tape, uint32_t n_ops, grd* slots, Input input) {
    grd result = gr1(nanf_());
    for (uint32_t i = 0; i < n_ops; ++i) {
        const uint2 w = __ldg(tape + i);
        Dec d(w.x);
        const float imm = __uint_as_float(w.y);
        const grd sl = slots[d.lhs], sr = slots[d.rhs];
        const grd a = d.form == F_IR ? gr1(imm) : sl;
        const grd b = d.form == F_RI ? gr1(imm) : sr;
        grd r;
        if (d.op == OP_MEM) {
            if (d.form == F_RI) slots[d.out] = slots[MEM_BASE + w.y];
            else slots[MEM_BASE + w.y] = sl;
            continue;
        } else if (d.op >= OP_ADD) {
            if (d.op == OP_MUL && d.form == F_RI) r = gr_mul_f(sl, imm);
            else r = gr_binary(d.op, a, b);
        } else if (d.op >= OP_NEG) r = gr_unary(d.op, sl);
        else if (d.op == OP_COPY) r = d.form == F_RI ? gr1(imm) : sl;
        else if (d.op == OP_INPUT) r = input(w.y);
        else { if (w.y == 0) result = sl; continue; }
        slots[d.out] = r;
    }
    return result;
}

// K3: normals + final image.  One thread per pixel; the gradient is evaluated
// at the surface voxel (x, y, depth - 1) with the tape of the
// leaf tile that found it (voxel.rs:449-481); lanes of a warp that share a
// leaf tile run its tape together.
__global__ void __launch_bounds__(128) k_normals_3d(const __grid_constant__ NormalParams p) {
    grd slots[REG_SLOTS];
    const int lane = threadIdx.x & 31;
    // 8x4 pixel patch per warp
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t patches_x = (p.width + 7u) / 8u, patches_y = (p.height + 3u) / 4u;
    if (warp >= patches_x * patches_y) return;
    const uint32_t x = (warp % patches_x) * 8u + (lane & 7), y = (warp / patches_x) * 4u + (lane >> 3);
    const bool inb = x < p.width && y < p.height;
    const unsigned long long key = inb ? p.heightmap[size_t(y) * p.width + x] : 0ull;
    const uint32_t depth = uint32_t(key >> 32), id = uint32_t(key);
    grd g = gr(0.0f, 0.0f, 0.0f, 0.0f);
    bool pending = inb && id != 0u;
    unsigned long long n = 0;
    for (;;) {
        const uint32_t m = __ballot_sync(FULL, pending);
        if (!m) break;
        const uint32_t lead_id = __shfl_sync(FULL, id, __ffs(m) - 1);
        const bool mine = pending && id == lead_id;
        const TileJob* job = p.jobs + (lead_id - 1u);
        const TapeRef tr = job->tape;
        grd gx, gy, gz;
        xform_gr(p.mat, gr(float(x), 1.0f, 0.0f, 0.0f), gr(float(y), 0.0f, 1.0f, 0.0f),
                 gr(float(depth - 1u), 0.0f, 0.0f, 1.0f), gx, gy, gz);
        const grd r = run_grad(tr.ptr, tr.n_ops, slots, [&](uint32_t i) {
            return pick_input(p.vb, i, gx, gy, gz, [](float f) { return gr1(f); });
        });
        if (mine) { g = r; pending = false; ++n; }
    }
    if (inb) {
        float4 o;
        if (p.clamp && depth >= p.depth - 1u) {   // voxel.rs:535-546
            o = make_float4(0.0f, 0.0f, 1.0f, __uint_as_float(p.depth));
        } else {
            o = make_float4(g.y, g.z, g.w, __uint_as_float(depth));
        }
        reinterpret_cast<float4*>(p.out)[size_t(y) * p.width + x] = o;
    }
    if (p.stats) {
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(FULL, n, o);
        if (lane == 0 && n) atomicAdd(&p.stats->grads, n);
    }
}
void launch_normals_3d(const NormalParams& p, cudaStream_t s) {
    const uint64_t warps = uint64_t((p.width + 7u) / 8u) * ((p.height + 3u) / 4u);
    k_normals_3d<<<unsigned((warps + 3) / 4), 128, 0, s>>>(p);
}

// Multi-GPU: per-pixel merge of Z-ordered slab images; the highest slab with
// the greatest depth wins, then the final clamp is applied.
__global__ void k_merge_slabs(const float4* const* slabs, uint32_t n_slabs, uint32_t n_pixels, uint32_t depth,
                              float4* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    float4 best = slabs[0][i];
    for (uint32_t s = 1; s < n_slabs; ++s) {
        const float4 c = slabs[s][i];
        if (__float_as_uint(c.w) >= __float_as_uint(best.w)) best = c;
    }
    if (__float_as_uint(best.w) >= depth - 1u) best = make_float4(0.0f, 0.0f, 1.0f, __uint_as_float(depth));
    out[i] = best;
}
void launch_merge_slabs(const void* const* d_slabs, uint32_t n_slabs, uint32_t n_pixels, uint32_t depth, void* out,
                        cudaStream_t s) {
    k_merge_slabs<<<(n_pixels + 255) / 256, 256, 0, s>>>(reinterpret_cast<const float4* const*>(d_slabs), n_slabs,
                                                         n_pixels, depth, reinterpret_cast<float4*>(out));
}

// ---------------------------------------------------------------------------
// K2: leaf pixels (2D)
__global__ void __launch_bounds__(128) k_pixels_2d(const __grid_constant__ PixelParams p) {
    const int lane = threadIdx.x & 31;
    float2 slots[REG_SLOTS];
    const uint32_t n_jobs = min(p.ctr->n_jobs[p.list], 0xffffffffu);
    const uint32_t T = p.tile, npix = T * T;
    unsigned long long shaded = 0;
    for (;;) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(&p.ctr->cursor[p.cursor], 1u);
        j = __shfl_sync(FULL, j, 0);
        if (j >= n_jobs) break;
        const TileJob* job = p.jobs + j;
        const uint32_t cx = job->x, cy = job->y;
        const TapeRef tr = job->tape;
        const uint2* tape = tr.ptr;
        for (uint32_t base = 0; base < npix; base += 64u) {
            uint32_t p0 = base + lane, p1 = p0 + 32u;
            bool v0 = p0 < npix, v1 = p1 < npix;
            uint32_t i0 = (v0 ? p0 : 0u) % T, j0 = (v0 ? p0 : 0u) / T;
            uint32_t i1 = (v1 ? p1 : 0u) % T, j1 = (v1 ? p1 : 0u) / T;
            float x0, y0, z0, x1, y1, z1;
            xform_f32(p.mat, float(cx + i0), float(cy + j0), p.z2d, x0, y0, z0);
            xform_f32(p.mat, float(cx + i1), float(cy + j1), p.z2d, x1, y1, z1);
            const float2 X = make_float2(x0, x1), Y = make_float2(y0, y1), Z = make_float2(z0, z1);
            float2 r = run_f32x2(tape, tr.n_ops, slots, [&](uint32_t i) {
                return pick_input(p.vb, i, X, Y, Z, [](float f) { return make_float2(f, f); });
            });
            // RawDistancePixel::from(f32): canonical NaN (pixel.rs:234-240)
            if (r.x != r.x) r.x = nanf_();
            if (r.y != r.y) r.y = nanf_();
            uint32_t gx0 = cx + i0, gy0 = cy + j0, gx1 = cx + i1, gy1 = cy + j1;
            if (v0 && gx0 < p.width && gy0 < p.height) p.out[size_t(gy0) * p.width + gx0] = r.x;
            if (v1 && gx1 < p.width && gy1 < p.height) p.out[size_t(gy1) * p.width + gx1] = r.y;
            shaded += (v0 ? 1 : 0) + (v1 ? 1 : 0);
        }
    }
    if (p.stats) {
        for (int o = 16; o > 0; o >>= 1) shaded += __shfl_xor_sync(FULL, shaded, o);
        if (lane == 0 && shaded) atomicAdd(&p.stats->pixels, shaded);
    }
}

void launch_pixels_2d(const PixelParams& p, int blocks, cudaStream_t s) { k_pixels_2d<<<blocks, 128, 0, s>>>(p); }

// ---------------------------------------------------------------------------
// Fill painter: one warp per (record, 1024-pixel unit)
__global__ void __launch_bounds__(256) k_fill_2d(const __grid_constant__ FillParams p) {
    const uint32_t n = *p.n_fills;
    const uint32_t T = p.tile;
    const uint32_t tile_px = T * T;
    const uint32_t unit_px = min(tile_px, 1024u);
    const uint32_t units = tile_px / unit_px;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const bool vec_ok = (p.width % 4u == 0u) && ((reinterpret_cast<uintptr_t>(p.out) & 15u) == 0u) && (T % 4u == 0u);
    const unsigned long long total = (unsigned long long)n * units;
    for (unsigned long long w = warp; w < total; w += n_warps) {
        const uint32_t rec = uint32_t(w / units), u = uint32_t(w % units);
        const FillRec fr = p.fills[rec];
        const float v = __uint_as_float(fr.value);
        const uint32_t first = u * unit_px;
        for (uint32_t q = lane * 4u; q < unit_px; q += 128u) {
            uint32_t pix = first + q;
            uint32_t x = fr.x + pix % T, y = fr.y + pix / T;
            if (y >= p.height) continue;
            float* dst = p.out + size_t(y) * p.width + x;
            if (vec_ok && x + 3u < p.width) {
                *reinterpret_cast<float4*>(dst) = make_float4(v, v, v, v);
            } else {
                for (uint32_t k = 0; k < 4u; ++k)
                    if (x + k < p.width && (pix + k) / T == pix / T) dst[k] = v;
            }
        }
    }
}

void launch_fill_2d(const FillParams& p, int blocks, cudaStream_t s) { k_fill_2d<<<blocks, 256, 0, s>>>(p); }

// ---------------------------------------------------------------------------
// Trait-level evaluators
template <int NSLOTS>
__global__ void __launch_bounds__(128) k_float_slice(const __grid_constant__ BulkParams p) {
    float slots[NSLOTS];
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t idx = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < p.n; idx += stride) {
        const uint2* tape = p.tape;
        for (uint32_t i = 0; i < p.n_ops; ++i) {
            uint2 w = __ldg(tape + i);
            Dec d(w.x);
            float imm = __uint_as_float(w.y);
            float sl = slots[d.lhs], sr = slots[d.rhs];
            float a = d.form == F_IR ? imm : sl;
            float b = d.form == F_RI ? imm : sr;
            float r;
            if (d.op == OP_MEM) {
                if (d.form == F_RI) slots[d.out] = slots[MEM_BASE + w.y];
                else slots[MEM_BASE + w.y] = sl;
                continue;
            } else if (d.op >= OP_ADD) r = f32_binary(d.op, a, b);
            else if (d.op >= OP_NEG) r = f32_unary(d.op, sl);
            else if (d.op == OP_COPY) r = d.form == F_RI ? imm : sl;
            else if (d.op == OP_INPUT) r = static_cast<const float*>(p.vars[w.y])[idx];
            else {
                static_cast<float*>(p.outs[w.y])[idx] = sl;
                continue;
            }
            slots[d.out] = r;
        }
    }
}

template <int NSLOTS>
__global__ void __launch_bounds__(128) k_grad_slice(const __grid_constant__ BulkParams p) {
    grd slots[NSLOTS];
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t idx = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < p.n; idx += stride) {
        const uint2* tape = p.tape;
        for (uint32_t i = 0; i < p.n_ops; ++i) {
            uint2 w = __ldg(tape + i);
            Dec d(w.x);
            float imm = __uint_as_float(w.y);
            grd sl = slots[d.lhs], sr = slots[d.rhs];
            grd a = d.form == F_IR ? gr1(imm) : sl;
            grd b = d.form == F_RI ? gr1(imm) : sr;
            grd r;
            if (d.op == OP_MEM) {
                if (d.form == F_RI) slots[d.out] = slots[MEM_BASE + w.y];
                else slots[MEM_BASE + w.y] = sl;
                continue;
            } else if (d.op >= OP_ADD) {
                if (d.op == OP_MUL && d.form == F_RI) r = gr_mul_f(sl, imm);
                else r = gr_binary(d.op, a, b);
            } else if (d.op >= OP_NEG) r = gr_unary(d.op, sl);
            else if (d.op == OP_COPY) r = d.form == F_RI ? gr1(imm) : sl;
            else if (d.op == OP_INPUT) r = static_cast<const grd*>(p.vars[w.y])[idx];
            else {
                static_cast<grd*>(p.outs[w.y])[idx] = sl;
                continue;
            }
            slots[d.out] = r;
        }
    }
}

static int bulk_blocks(uint64_t n) {
    uint64_t b = (n + 127) / 128;
    return int(b < 1 ? 1 : (b > 148ull * 16 ? 148ull * 16 : b));
}
void launch_float_slice(const BulkParams& p, cudaStream_t s) {
    if (p.n == 0) return;
    if (p.n_slots <= 256) k_float_slice<256><<<bulk_blocks(p.n), 128, 0, s>>>(p);
    else k_float_slice<2048><<<bulk_blocks(p.n), 128, 0, s>>>(p);
}
void launch_grad_slice(const BulkParams& p, cudaStream_t s) {
    if (p.n == 0) return;
    if (p.n_slots <= 256) k_grad_slice<256><<<bulk_blocks(p.n), 128, 0, s>>>(p);
    else k_grad_slice<2048><<<bulk_blocks(p.n), 128, 0, s>>>(p);
}

struct ByteChoiceSink {
    uint8_t* base;  // may be null
    uint32_t ci = 0;
    bool any_nonboth = false;
    __device__ __forceinline__ void push(uint32_t c) {
        if (base) base[ci] = uint8_t(c);
        ++ci;
        any_nonboth |= (c != 3u);
    }
};

template <int NSLOTS>
__global__ void __launch_bounds__(128) k_interval_batch(const __grid_constant__ TracingParams p) {
    itv slots[NSLOTS];
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t idx = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < p.n; idx += stride) {
        const float* v = p.vars + idx * p.n_vars * 2;
        float* o = p.out + idx * p.n_outputs * 2;
        ByteChoiceSink sink;
        sink.base = p.choices ? p.choices + idx * p.n_choices : nullptr;
        run_interval(
            p.tape, p.n_ops, slots, [&](uint32_t i) { return iv(v[2 * i], v[2 * i + 1]); }, sink,
            [&](uint32_t oi, itv val) { o[2 * oi] = val.x; o[2 * oi + 1] = val.y; });
        if (p.simplify) p.simplify[idx] = sink.any_nonboth ? 1 : 0;
    }
}

template <int NSLOTS>
__global__ void __launch_bounds__(128) k_point_batch(const __grid_constant__ TracingParams p) {
    float slots[NSLOTS];
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t idx = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < p.n; idx += stride) {
        const float* v = p.vars + idx * p.n_vars;
        float* o = p.out + idx * p.n_outputs;
        ByteChoiceSink sink;
        sink.base = p.choices ? p.choices + idx * p.n_choices : nullptr;
        for (uint32_t i = 0; i < p.n_ops; ++i) {
            uint2 w = __ldg(p.tape + i);
            Dec d(w.x);
            float imm = __uint_as_float(w.y);
            float sl = slots[d.lhs], sr = slots[d.rhs];
            float a = d.form == F_IR ? imm : sl;
            float b = d.form == F_RI ? imm : sr;
            float r;
            if (d.op == OP_MEM) {
                if (d.form == F_RI) slots[d.out] = slots[MEM_BASE + w.y];
                else slots[MEM_BASE + w.y] = sl;
                continue;
            } else if (d.op >= OP_MIN) {
                r = f32_binary(d.op, a, b);
                sink.push(f32_choice(d.op, a, b));
            } else if (d.op >= OP_ADD) r = f32_binary(d.op, a, b);
            else if (d.op >= OP_NEG) r = f32_unary(d.op, sl);
            else if (d.op == OP_COPY) r = d.form == F_RI ? imm : sl;
            else if (d.op == OP_INPUT) r = v[w.y];
            else { o[w.y] = sl; continue; }
            slots[d.out] = r;
        }
        if (p.simplify) p.simplify[idx] = sink.any_nonboth ? 1 : 0;
    }
}

void launch_interval_batch(const TracingParams& p, cudaStream_t s) {
    if (p.n == 0) return;
    if (p.n_slots <= 256) k_interval_batch<256><<<bulk_blocks(p.n), 128, 0, s>>>(p);
    else k_interval_batch<2048><<<bulk_blocks(p.n), 128, 0, s>>>(p);
}
void launch_point_batch(const TracingParams& p, cudaStream_t s) {
    if (p.n == 0) return;
    if (p.n_slots <= 256) k_point_batch<256><<<bulk_blocks(p.n), 128, 0, s>>>(p);
    else k_point_batch<2048><<<bulk_blocks(p.n), 128, 0, s>>>(p);
}

__global__ void k_simplify_single(const __grid_constant__ SimplifyParams p) {
    __shared__ uint32_t live_s[8][32];
    const int lane = threadIdx.x;
    ByteChoiceSource src;
    src.base = p.choices;
    src.ci = p.n_choices;
    uint32_t n_dev = 0, ref_len = 0, nch = 0;
    simplify_lane(p.parent, p.n_ops, lane == 0, live_s, lane, src, p.out + p.n_ops, n_dev, ref_len, nch);
    if (lane == 0) {
        p.result[0] = n_dev;
        p.result[1] = ref_len;
        p.result[2] = nch;
    }
}
void launch_simplify_single(const SimplifyParams& p, cudaStream_t s) { k_simplify_single<<<1, 32, 0, s>>>(p); }

}  // namespace fdev

// ---------------------------------------------------------------------------
// K1-root: cooperative level-0 kernel.  The root tape is long (prospero: 6363
// clauses) and there are few root tiles (1024 at 4096^2), so one lane per
// tile is latency-bound.  Here one CTA evaluates one root tile: the tape's
// clauses run wave by wave (all clauses of a wave are independent), values
// live in shared memory indexed by the defining clause, long min/max chains
// in the tail are evaluated with a block-wide prefix scan, and the reverse
// liveness pass + compaction are parallel too.  Results are identical to
// k_interval_level_2d (same per-clause arithmetic, same simplify rules).
namespace fdev {

// shared memory of one root tile: forward values by slot, overlaid by the reverse pass's
// last_use words (one per clause; bits 16.. hold the emit code), then the 2-bit choices
size_t coop_smem_bytes(uint32_t n_ops, uint32_t n_choices, uint32_t n_slots) {
    return std::max(size_t(n_slots) * 8, size_t(n_ops) * 4) + size_t((n_choices + 15) / 16 + 1) * 4 + 16;
}

struct Fwd {
    uint32_t x, y, sa, sb, so, cidx;
    __device__ __forceinline__ explicit Fwd(const uint4 q)
        : x(q.x), y(q.y), sa(q.z & 0xffffu), sb(q.z >> 16), so(q.w & 0xffffu), cidx(q.w >> 16) {}
};
__device__ __forceinline__ Fwd load_fwd(const CoopFwd* f, uint32_t i) {
    return Fwd(__ldg(reinterpret_cast<const uint4*>(f) + i));
}
struct Rec {
    uint32_t x, y, ia, ib, p, cidx;
    __device__ __forceinline__ explicit Rec(const uint4 q)
        : x(q.x), y(q.y), ia(q.z & 0xffffu), ib(q.z >> 16), p(q.w & 0xffffu), cidx(q.w >> 16) {}
};
__device__ __forceinline__ Rec load_rec(const CoopRec* recs, uint32_t i) {
    return Rec(__ldg(reinterpret_cast<const uint4*>(recs) + i));
}

template <int DIM>
__global__ void __launch_bounds__(COOP_THREADS)
k_interval_root_coop(const __grid_constant__ LevelParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t n = p.root_tape.n_ops, nch = p.root_tape.n_choices;
    const uint32_t cw = (nch + 15u) / 16u + 1u;
    itv* vals = reinterpret_cast<itv*>(smem_raw);                 // forward: values by slot
    const size_t data_bytes = max(size_t(p.sched.n_slots) * 8, size_t(n) * 4);
    uint32_t* chs = reinterpret_cast<uint32_t*>(smem_raw + data_bytes);
    // reverse: one word per clause, overlaying vals: bits 0..15 = 1 + position of the last live
    // reader (0: dead), bits 16.. = emit code
    uint32_t* last_use = reinterpret_cast<uint32_t*>(smem_raw);
    __shared__ itv s_res;
    __shared__ uint32_t s_tile, s_nonboth, s_warp_tot[COOP_THREADS / 32], s_ref, s_nch;
    __shared__ unsigned long long s_base;
    __shared__ float s_agg_lo[COOP_THREADS / 32], s_agg_hi[COOP_THREADS / 32];
    __shared__ uint8_t s_agg_f[COOP_THREADS / 32], s_agg_u[COOP_THREADS / 32];

    const uint32_t tid = threadIdx.x, T = p.tile, NT = blockDim.x;   // NT <= COOP_THREADS, a multiple of 32
    const CoopRec* __restrict__ recs = p.sched.recs;
    const CoopFwd* __restrict__ fwd = p.sched.fwd;
    const uint32_t* __restrict__ ws = p.sched.wave_start;
    const uint32_t n_roots = p.roots_x * p.roots_y * (DIM == 3 ? p.roots_z : 1u);
    const uint2* __restrict__ tape = p.root_tape.ptr;

    for (;;) {
        if (tid == 0) {
            s_tile = atomicAdd(&p.ctr->cursor[0], 1u);
            s_nonboth = 0;
            s_ref = 0;
            s_nch = 0;
            s_res = iv_nan();
        }
        for (uint32_t i = tid; i < cw; i += NT) chs[i] = 0;
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_roots) break;
        const uint32_t cx = p.root_x0 + (tile % p.roots_x) * T, cy = p.root_y0 + ((tile / p.roots_x) % p.roots_y) * T;
        const uint32_t cz = DIM == 3 ? p.root_z0 + (tile / (p.roots_x * p.roots_y)) * T : 0u;
        itv vx, vy, vz;
        xform_iv(p.mat, iv(float(cx), float(cx) + float(T)), iv(float(cy), float(cy) + float(T)),
                 DIM == 3 ? iv(float(cz), float(cz) + float(T)) : iv(p.z2d, p.z2d), vx, vy, vz);
        auto put_choice = [&](uint32_t cidx, uint32_t c) {
            atomicOr(&chs[cidx >> 4], c << ((cidx & 15u) * 2u));
            if (c != 3u) s_nonboth = 1u;
        };
        auto get_choice = [&](uint32_t cidx) { return (chs[cidx >> 4] >> ((cidx & 15u) * 2u)) & 3u; };
        auto exec = [&](const Fwd& rc, itv sl, itv sr) -> itv {
            Dec d(rc.x);
            float imm = __uint_as_float(rc.y);
            itv a = d.form == F_IR ? iv1(imm) : sl;
            itv b = d.form == F_RI ? iv1(imm) : sr;
            itv r;
            if (d.op >= OP_MIN) {
                uint32_t c;
                r = iv_choice_op(d.op, a, b, c);
                put_choice(rc.cidx, c);
            } else if (d.op >= OP_ADD) {
                if (d.op == OP_MUL && d.form == F_RI) r = iv_mul_f(sl, imm);
                else r = iv_binary(d.op, a, b);
            } else if (d.op >= OP_NEG) {
                r = iv_unary(d.op, sl);
            } else if (d.op == OP_COPY) {
                r = d.form == F_RI ? iv1(imm) : sl;
            } else if (d.op == OP_INPUT) {
                r = pick_input(p.vb, rc.y, vx, vy, vz, [](float f) { return iv1(f); });
            } else {
                if (rc.y == 0) s_res = sl;
                r = sl;
            }
            return r;
        };
        auto ld = [&](uint32_t id) { return id != COOP_NONE ? vals[id] : iv_nan(); };

        // ---- forward: dependency waves ----
        {
            // each thread keeps the NEXT record it will execute in registers, so the
            // global (L2) latency of the schedule stream overlaps the current clause
            uint32_t w = 0, i = tid;   // recs of wave w are [ws[w], ws[w+1]); wave 0 starts at 0
            const uint32_t n_waves = p.sched.n_waves, wave_end_all = p.sched.tail_begin;
            uint32_t e = n_waves ? ws[1] : 0;
            auto advance = [&]() {     // move (w, i) to this thread's next record, crossing waves
                while (w < n_waves && i >= e) {
                    ++w;
                    if (w < n_waves) { i = e + tid; e = ws[w + 1]; }
                }
            };
            advance();
            uint4 q = (w < n_waves) ? __ldg(reinterpret_cast<const uint4*>(fwd) + i) : make_uint4(0, 0, 0, 0);
            uint32_t cur_w = 0;
            while (cur_w < n_waves) {
                // run everything this thread owns in wave cur_w
                while (w == cur_w) {
                    const Fwd rc(q);
                    i += NT;
                    advance();
                    if (w < n_waves) q = __ldg(reinterpret_cast<const uint4*>(fwd) + i);
                    const itv r = exec(rc, ld(rc.sa), ld(rc.sb));
                    if (rc.so != COOP_NONE) vals[rc.so] = r;
                }
                __syncthreads();
                ++cur_w;
            }
            (void)wave_end_all;
        }
        // ---- forward: tail segments ----
        for (uint32_t sgi = 0; sgi < p.sched.n_segs; ++sgi) {
            const uint32_t b = p.sched.segs[sgi].begin, e = p.sched.segs[sgi].end;
            if (!p.sched.segs[sgi].chain) {
                if (tid == 0) {
                    uint32_t last_s = COOP_NONE;
                    itv last_r = iv_nan();
                    for (uint32_t i = b; i < e; ++i) {
                        const Fwd rc = load_fwd(fwd, i);
                        const itv sl = (rc.sa == last_s && last_s != COOP_NONE) ? last_r : ld(rc.sa);
                        const itv sr = (rc.sb == last_s && last_s != COOP_NONE) ? last_r : ld(rc.sb);
                        const itv r = exec(rc, sl, sr);
                        if (rc.so != COOP_NONE) vals[rc.so] = r;
                        last_s = rc.so;
                        last_r = r;
                    }
                }
            } else {
                // m_i = OP(m_{i-1}, s_i): prefix scan over the sides
                const uint32_t m = e - b, ch = (m + NT - 1) / NT;
                const uint32_t c0 = min(e, b + tid * ch), c1 = min(e, c0 + ch);
                // (forward view: the operand that is the previous chain value is marked COOP_NONE,
                //  the value the chain starts from sits in segs[].start_slot)
                const Fwd first = load_fwd(fwd, b);
                const bool is_min = (Dec(first.x).op == OP_MIN);
                const uint32_t start_slot = p.sched.segs[sgi].start_slot;
                auto comb = [&](float& lo, float& hi, uint32_t& f, itv s) {
                    f |= uint32_t(iv_has_nan(s));
                    lo = is_min ? fminf(lo, s.x) : fmaxf(lo, s.x);
                    hi = is_min ? fminf(hi, s.y) : fmaxf(hi, s.y);
                };
                const float ident = is_min ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
                float alo = ident, ahi = ident;
                uint32_t af = 0;
                for (uint32_t i = c0; i < c1; ++i) {
                    const Fwd rc = load_fwd(fwd, i);
                    comb(alo, ahi, af, vals[rc.sa == COOP_NONE ? rc.sb : rc.sa]);
                }
                // exclusive block scan of the per-thread aggregates (warp shuffles + one smem hop)
                float xlo = alo, xhi = ahi;
                uint32_t xf = af;
                const uint32_t ln = tid & 31u, wp = tid >> 5;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float vlo = __shfl_up_sync(FULL, xlo, o), vhi = __shfl_up_sync(FULL, xhi, o);
                    const uint32_t vf = __shfl_up_sync(FULL, xf, o);
                    if (ln >= uint32_t(o)) {
                        xlo = is_min ? fminf(vlo, xlo) : fmaxf(vlo, xlo);
                        xhi = is_min ? fminf(vhi, xhi) : fmaxf(vhi, xhi);
                        xf |= vf;
                    }
                }
                if (ln == 31u) { s_agg_lo[wp] = xlo; s_agg_hi[wp] = xhi; s_agg_f[wp] = uint8_t(xf); }
                // exclusive within the warp
                float elo = __shfl_up_sync(FULL, xlo, 1), ehi = __shfl_up_sync(FULL, xhi, 1);
                uint32_t ef = __shfl_up_sync(FULL, xf, 1);
                if (ln == 0u) { elo = ident; ehi = ident; ef = 0; }
                __syncthreads();
                if (c0 < c1) {
                    const itv start = vals[start_slot];
                    float lo = start.x, hi = start.y;
                    uint32_t f = uint32_t(iv_has_nan(start));
                    for (uint32_t k = 0; k < wp; ++k) {
                        f |= s_agg_f[k];
                        lo = is_min ? fminf(lo, s_agg_lo[k]) : fmaxf(lo, s_agg_lo[k]);
                        hi = is_min ? fminf(hi, s_agg_hi[k]) : fmaxf(hi, s_agg_hi[k]);
                    }
                    f |= ef;
                    lo = is_min ? fminf(lo, elo) : fmaxf(lo, elo);
                    hi = is_min ? fminf(hi, ehi) : fmaxf(hi, ehi);
                    for (uint32_t i = c0; i < c1; ++i) {
                        const Fwd rc = load_fwd(fwd, i);
                        const bool prev_is_lhs = (rc.sa == COOP_NONE);
                        const itv s = vals[prev_is_lhs ? rc.sb : rc.sa];
                        const itv mprev = f ? iv_nan() : iv(lo, hi);
                        uint32_t c;
                        const itv r = prev_is_lhs ? iv_choice_op(is_min ? OP_MIN : OP_MAX, mprev, s, c)
                                                  : iv_choice_op(is_min ? OP_MIN : OP_MAX, s, mprev, c);
                        comb(lo, hi, f, s);
                        put_choice(rc.cidx, c);
                        if (rc.so != COOP_NONE) vals[rc.so] = r;
                    }
                }
            }
            __syncthreads();
        }

        const itv r = s_res;
        const bool fill_in = !p.pixel_perfect && r.y < 0.0f;
        const bool fill_out = !p.pixel_perfect && !fill_in && r.x > 0.0f;
        const bool amb = !fill_in && !fill_out;
        if (DIM == 3 && fill_in) {   // voxel.rs:310-317
            const unsigned long long key = (unsigned long long)(cz + T + 1u) << 32;
            for (uint32_t q = tid; q < T * T; q += NT) {
                const uint32_t x = cx + q % T, y = cy + q / T;
                if (x < p.width && y < p.height) atomicMax(&p.heightmap[size_t(y) * p.width + x], key);
            }
        }
        if (tid == 0) {
            if (DIM == 2 && !amb) {
                uint32_t slot = atomicAdd(&p.ctr->n_fills[0], 1u);
                if (slot < p.cap_fills) {
                    FillRec fr;
                    fr.x = cx;
                    fr.y = cy;
                    fr.value = 0x7FC00000u | (fill_in ? 1u : 0u) | (0xF6u << 9);
                    p.fills[slot] = fr;
                } else atomicOr(&p.ctr->error, 2u);
            }
            if (p.stats) {
                atomicAdd(&p.stats->evaluated[0], 1ull);
                if (fill_in) atomicAdd(&p.stats->filled_inside[0], 1ull);
                if (fill_out) atomicAdd(&p.stats->filled_outside[0], 1ull);
                if (amb) atomicAdd(&p.stats->ambiguous[0], 1ull);
            }
        }
        if (!amb) { __syncthreads(); continue; }

        TapeRef child = p.root_tape;
        if (s_nonboth) {   // uniform: written before the last barrier
            // ---- R1: reverse liveness; last_use[v] = 1 + position of the last live clause reading v ----
            for (uint32_t i = tid; i < n; i += NT) last_use[i] = 0;
            __syncthreads();
            auto r1 = [&](const Rec& rc) {
                Dec d(rc.x);
                if (d.op != OP_OUTPUT && last_use[rc.p] == 0u) return;
                const uint32_t mark = rc.p + 1u;
                bool use_a = rc.ia != COOP_NONE, use_b = rc.ib != COOP_NONE;
                if (d.op >= OP_MIN) {
                    uint32_t c = get_choice(rc.cidx);
                    if (c == 1u) use_b = false;
                    else if (c == 2u) use_a = false;
                }
                if (use_a) atomicMax(&last_use[rc.ia], mark);
                if (use_b) atomicMax(&last_use[rc.ib], mark);
            };
            for (uint32_t sgi = p.sched.n_segs; sgi > 0; --sgi) {
                const uint32_t b = p.sched.segs[sgi - 1].begin, e = p.sched.segs[sgi - 1].end;
                if (!p.sched.segs[sgi - 1].chain) {
                    if (tid == 0)
                        for (uint32_t i = e; i > b; --i) r1(load_rec(recs, i - 1));
                } else {
                    // live_i = ext_i | (uses_prev_{i+1} & live_{i+1}), suffix scan over the run
                    const uint32_t m = e - b, ch = (m + NT - 1) / NT;
                    const uint32_t c0 = min(e, b + tid * ch), c1 = min(e, c0 + ch);
                    const uint32_t prev_first = load_rec(recs, b - 1).p;
                    auto uses_prev = [&](const Rec& rc, uint32_t prevp) {
                        uint32_t c = get_choice(rc.cidx);
                        return c == 3u || c == (rc.ia == prevp ? 1u : 2u);
                    };
                    // F(x) = O | (U & x), x = (uses_prev & live) of the element after the chunk
                    uint32_t O = 0, U = 1;
                    for (uint32_t i = c1; i > c0; --i) {
                        const Rec rc = load_rec(recs, i - 1);
                        const uint32_t prevp = i - 1 > b ? load_rec(recs, i - 2).p : prev_first;
                        const uint32_t ext = last_use[rc.p] != 0u, up = uses_prev(rc, prevp);
                        // y_i = up_i & live_i, live_i = ext_i | y_{i+1}
                        O = up & (ext | O);
                        U = up & U;
                    }
                    // inclusive suffix scan of F = (O, U) under composition (earlier o later)
                    uint32_t xO = O, xU = U;
                    const uint32_t ln = tid & 31u, wp = tid >> 5;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t vO = __shfl_down_sync(FULL, xO, o), vU = __shfl_down_sync(FULL, xU, o);
                        if (ln + uint32_t(o) < 32u) { xO = xO | (xU & vO); xU = xU & vU; }
                    }
                    if (ln == 0u) { s_agg_f[wp] = uint8_t(xO); s_agg_u[wp] = uint8_t(xU); }
                    // composition of the chunks AFTER this thread inside the warp
                    uint32_t eO = __shfl_down_sync(FULL, xO, 1), eU = __shfl_down_sync(FULL, xU, 1);
                    if (ln == 31u) { eO = 0; eU = 1; }
                    __syncthreads();
                    if (c0 < c1) {
                        uint32_t y = 0;   // (uses_prev & live) of the element right after this chunk
                        for (uint32_t k = (NT >> 5); k > wp + 1u; --k) y = s_agg_f[k - 1] | (s_agg_u[k - 1] & y);
                        y = eO | (eU & y);
                        for (uint32_t i = c1; i > c0; --i) {
                            const Rec rc = load_rec(recs, i - 1);
                            const uint32_t prevp = i - 1 > b ? load_rec(recs, i - 2).p : prev_first;
                            const uint32_t live = (last_use[rc.p] != 0u) | y;
                            const uint32_t c = get_choice(rc.cidx);
                            const bool prev_is_lhs = (rc.ia == prevp);
                            const bool up = c == 3u || c == (prev_is_lhs ? 1u : 2u);
                            const bool us = c == 3u || c == (prev_is_lhs ? 2u : 1u);
                            if (live) {
                                if (up) atomicMax(&last_use[prevp], rc.p + 1u);
                                if (us) atomicMax(&last_use[prev_is_lhs ? rc.ib : rc.ia], rc.p + 1u);
                            }
                            y = up & live;
                        }
                    }
                }
                __syncthreads();
            }
            for (uint32_t w = p.sched.n_waves; w > 0; --w) {
                const uint32_t e = ws[w];
                for (uint32_t i = ws[w - 1] + tid; i < e; i += NT) r1(load_rec(recs, i));
                __syncthreads();
            }
            // ---- R2a: what each clause turns into ----
            // 0 none, 1 as is, 7 as is (choice kept), 2/3 copy lhs real/alias, 4/5 copy rhs real/alias, 6 copy imm
            for (uint32_t i = tid; i < p.sched.tail_end; i += NT) {
                const Rec rc = load_rec(recs, i);
                Dec d(rc.x);
                const uint32_t pos = rc.p;
                uint32_t code;
                if (d.op != OP_OUTPUT && (last_use[pos] & 0xffffu) == 0u) code = 0;
                else if (d.op >= OP_MIN) {
                    uint32_t c = get_choice(rc.cidx);
                    if (c == 3u) code = 7;
                    else if (c == 2u && d.form == F_RI) code = 6;
                    else {
                        const bool use_rhs = (c == 2u);
                        const uint32_t src_reg = use_rhs ? d.rhs : d.lhs, src_def = use_rhs ? rc.ib : rc.ia;
                        if (src_reg == d.out) code = 0;
                        else code = (use_rhs ? 4u : 2u) + ((last_use[src_def] & 0xffffu) > pos + 1u ? 0u : 1u);
                    }
                } else if (d.op == OP_COPY && d.form != F_RI) {
                    if (d.lhs == d.out) code = 0;
                    else if (d.form == F_ALIAS) code = 3;
                    else code = (last_use[rc.ia] & 0xffffu) > pos + 1u ? 2u : 3u;
                } else code = 1;
                last_use[pos] = (last_use[pos] & 0xffffu) | (code << 16);   // only this thread writes word `pos`
            }
            __syncthreads();
            // ---- R2b: scan in tape order, then write the compacted child ----
            const uint32_t chunk = (n + NT - 1) / NT;
            const uint32_t b0 = min(n, tid * chunk), b1 = min(n, b0 + chunk);
            uint32_t my_dev = 0, my_ref = 0, my_nch = 0;
            for (uint32_t q = b0; q < b1; ++q) {
                uint32_t c = last_use[q] >> 16;
                my_dev += (c != 0u);
                my_ref += (c != 0u && c != 3u && c != 5u);
                my_nch += (c == 7u);
            }
            uint32_t incl = my_dev;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(FULL, incl, o);
                if ((tid & 31) >= uint32_t(o)) incl += v;
            }
            if ((tid & 31) == 31) s_warp_tot[tid >> 5] = incl;
            if (my_ref) atomicAdd(&s_ref, my_ref);
            if (my_nch) atomicAdd(&s_nch, my_nch);
            __syncthreads();
            uint32_t warp_off = 0, n_dev = 0;
            for (uint32_t k = 0; k < (NT >> 5); ++k) {
                if (k < (tid >> 5)) warp_off += s_warp_tot[k];
                n_dev += s_warp_tot[k];
            }
            const uint32_t ref_len = s_ref, nch_c = s_nch;
            const bool keep = ref_len < p.root_tape.ref_len;   // render/mod.rs:125-129
            if (keep) {
                if (tid == 0) {
                    unsigned long long base = atomicAdd(&p.ctr->arena_top, (unsigned long long)n_dev);
                    if (base + n_dev > p.arena_cap) {
                        atomicOr(&p.ctr->error, 1u);
                        base = ~0ull;
                    }
                    s_base = base;
                }
                __syncthreads();
                const unsigned long long base = s_base;
                if (base != ~0ull) {
                    uint2* dst = p.arena + base + warp_off + (incl - my_dev);
                    for (uint32_t q = b0; q < b1; ++q) {
                        uint32_t c = last_use[q] >> 16;
                        if (!c) continue;
                        uint2 w = __ldg(tape + q);
                        if (c != 1u && c != 7u) {
                            Dec d(w.x);
                            if (c == 6u) w = make_uint2(enc(OP_COPY, F_RI, d.out, 0xff, 0xff), w.y);
                            else {
                                uint32_t src = (c >= 4u) ? d.rhs : d.lhs;
                                w = make_uint2(enc(OP_COPY, (c & 1u) ? F_ALIAS : F_RR, d.out, src, 0xff), 0xFF000000u);
                            }
                        }
                        *dst++ = w;
                    }
                    child.ptr = p.arena + base;
                    child.n_ops = n_dev;
                    child.ref_len = ref_len;
                    child.n_choices = nch_c;
                    if (tid == 0 && p.stats) atomicAdd(&p.stats->simplified[0], 1ull);
                }
            }
        }
        if (tid == 0) {
            uint32_t slot = atomicAdd(&p.ctr->n_jobs[1], 1u);
            if (slot < p.cap_out) {
                TileJob o;
                o.x = cx;
                o.y = cy;
                o.z = cz;
                o.pad = 0;
                o.tape = child;
                p.jobs_out[slot] = o;
            } else atomicOr(&p.ctr->error, 2u);
        }
        __syncthreads();
    }
}

template <int DIM>
static cudaError_t launch_coop(const LevelParams& p, int blocks, int threads, cudaStream_t s) {
    size_t smem = coop_smem_bytes(p.root_tape.n_ops, p.root_tape.n_choices, p.sched.n_slots);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_interval_root_coop<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        // many small CTAs per SM: ask for the largest shared-memory carve-out
        cudaFuncSetAttribute(k_interval_root_coop<DIM>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = smem;
    }
    k_interval_root_coop<DIM><<<blocks, threads, smem, s>>>(p);
    return cudaGetLastError();
}
int coop_occupancy(int dim, int threads, size_t smem) {
    int n = 0;
    if (dim == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_interval_root_coop<3>, threads, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_interval_root_coop<2>, threads, smem);
    return n;
}
int coop_regs_per_thread(int dim) {
    static int regs[2] = {0, 0};
    int& r = regs[dim == 3];
    if (!r) {
        cudaFuncAttributes a{};
        cudaError_t e = dim == 3 ? cudaFuncGetAttributes(&a, k_interval_root_coop<3>) : cudaFuncGetAttributes(&a, k_interval_root_coop<2>);
        r = e == cudaSuccess ? a.numRegs : 64;
    }
    return r;
}
cudaError_t launch_interval_root_coop_2d(const LevelParams& p, int blocks, int threads, cudaStream_t s) { return launch_coop<2>(p, blocks, threads, s); }
cudaError_t launch_interval_root_coop_3d(const LevelParams& p, int blocks, int threads, cudaStream_t s) { return launch_coop<3>(p, blocks, threads, s); }

}  // namespace fdev

// ---------------------------------------------------------------------------
// Octree sampler leaves (OctreeBuilder::leaf, fidget-mesh/src/octree.rs:590-808):
// 8 corner samples -> corner mask; for every edge whose corners differ, 4 rounds
// of 16-ary search from the inside corner to the outside one; the intersection
// is the midpoint of the final bracket.  One warp per leaf; a pass handles four
// edges (two half-warps x two points per lane), `frac` comes from a ballot.
namespace fdev {

__device__ __forceinline__ float lerp_u16(float lo, float hi, uint32_t p) {
    const float frac = float(p) / 65535.0f;   // CellBounds::pos (cell.rs:183-192), Interval::lerp
    return lo * (1.0f - frac) + hi * frac;
}

struct EdgeState { uint32_t s[3], e[3]; };

// Edge `index` (= 4 t + 2 [start & v] + [start & u], types.rs:208-219) of a cell with corner `mask`
__device__ __forceinline__ bool edge_setup(uint32_t index, uint32_t mask, EdgeState& st) {
    const uint32_t t = index >> 2, su = index & 1u, sv = (index >> 1) & 1u;
    const uint32_t u = (t + 1u) % 3u, v = (t + 2u) % 3u;
    const uint32_t c0 = (su << u) | (sv << v), c1 = c0 | (1u << t);
    const bool in0 = (mask >> c0) & 1u, in1 = (mask >> c1) & 1u;
    if (in0 == in1) return false;
    st.s[u] = st.e[u] = su ? 65535u : 0u;
    st.s[v] = st.e[v] = sv ? 65535u : 0u;
    st.s[t] = in0 ? 0u : 65535u;   // the search runs inside -> outside
    st.e[t] = in0 ? 65535u : 0u;
    return true;
}

__global__ void __launch_bounds__(128) k_octree_leaf(const __grid_constant__ OctreeLeafParams p) {
    const int lane = threadIdx.x & 31;
    float2 slots[REG_SLOTS];
    const uint32_t n_jobs = min(p.ctr->n_jobs[p.list], p.cap_jobs);
    unsigned long long n_empty = 0, n_full = 0, n_surf = 0, n_pts = 0;
    for (;;) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(&p.ctr->cursor[p.cursor], 1u);
        j = __shfl_sync(FULL, j, 0);
        if (j >= n_jobs) break;
        const TileJob* job = p.jobs + j;
        const uint32_t cx = job->x, cy = job->y, cz = job->z;
        const TapeRef tr = job->tape;
        const float h = p.cell_h;
        const float lo[3] = {float(cx) * h - 1.0f, float(cy) * h - 1.0f, float(cz) * h - 1.0f};
        const float hi[3] = {float(cx + 1u) * h - 1.0f, float(cy + 1u) * h - 1.0f, float(cz + 1u) * h - 1.0f};
        auto eval2 = [&](float x0, float y0, float z0, float x1, float y1, float z1) -> float2 {
            if (p.has_transform) {
                xform_f32(p.mat, x0, y0, z0, x0, y0, z0);
                xform_f32(p.mat, x1, y1, z1, x1, y1, z1);
            }
            const float2 X = make_float2(x0, x1), Y = make_float2(y0, y1), Z = make_float2(z0, z1);
            return run_f32x2(tr.ptr, tr.n_ops, slots, [&](uint32_t i) {
                return pick_input(p.vb, i, X, Y, Z, [](float f) { return make_float2(f, f); });
            });
        };
        // corners (CellBounds::corner: bit i of the corner index selects the upper bound on axis i)
        const int c = lane & 7;
        const float2 cv = eval2((c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2],
                                lo[0], lo[1], lo[2]);
        if (lane == 0) n_pts += 8;
        const uint32_t mask = __ballot_sync(FULL, cv.x < 0.0f) & 0xffu;
        if (mask == 0u) { ++n_empty; continue; }
        if (mask == 255u) { ++n_full; continue; }
        ++n_surf;
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(p.n_out, 1u);
        slot = __shfl_sync(FULL, slot, 0);
        if (slot >= p.cap_out) {
            if (lane == 0) atomicOr(&p.ctr->error, 2u);
            continue;
        }
        OctreeLeaf* L = p.out + slot;
        if (lane == 0) p.out_tapes[slot] = tr;
        // active edges, ascending undirected index
        uint32_t active = 0;
        for (uint32_t e = 0; e < 12u; ++e) {
            EdgeState tmp;
            if (edge_setup(e, mask, tmp)) active |= 1u << e;
        }
        const uint32_t ne = __popc(active);
        if (lane == 0) {
            n_pts += 64ull * ne;
            L->ix = uint16_t(cx); L->iy = uint16_t(cy); L->iz = uint16_t(cz);
            L->mask = uint8_t(mask); L->n_edges = uint8_t(ne);
            L->present = uint16_t(active); L->pad = 0;
        }
        const int half = lane >> 4, jj = lane & 15;
        for (uint32_t pass = 0; pass * 4u < ne; ++pass) {
            // this lane follows edges k0 (component x) and k1 (component y) of the pass
            const uint32_t k0 = pass * 4u + uint32_t(half), k1 = k0 + 2u;
            auto nth = [&](uint32_t k) {   // index of the k-th set bit of `active`
                uint32_t m = active;
                for (uint32_t q = 0; q < k; ++q) m &= m - 1u;
                return uint32_t(__ffs(m) - 1);
            };
            const bool v0 = k0 < ne, v1 = k1 < ne;
            const uint32_t e0 = v0 ? nth(k0) : nth(0), e1 = v1 ? nth(k1) : nth(0);
            EdgeState s0, s1;
            edge_setup(e0, mask, s0);
            edge_setup(e1, mask, s1);
            for (int round = 0; round < 4; ++round) {
                uint32_t q0[3], q1[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    q0[a] = (s0.s[a] * uint32_t(15 - jj) + s0.e[a] * uint32_t(jj)) / 15u;
                    q1[a] = (s1.s[a] * uint32_t(15 - jj) + s1.e[a] * uint32_t(jj)) / 15u;
                }
                const float2 v = eval2(lerp_u16(lo[0], hi[0], q0[0]), lerp_u16(lo[1], hi[1], q0[1]), lerp_u16(lo[2], hi[2], q0[2]),
                                       lerp_u16(lo[0], hi[0], q1[0]), lerp_u16(lo[1], hi[1], q1[1]), lerp_u16(lo[2], hi[2], q1[2]));
                const uint32_t b0 = (__ballot_sync(FULL, v.x >= 0.0f) >> (16 * half)) & 0xffffu;
                const uint32_t b1 = (__ballot_sync(FULL, v.y >= 0.0f) >> (16 * half)) & 0xffffu;
                auto narrow = [&](EdgeState& st, uint32_t bits) {
                    uint32_t frac = bits ? uint32_t(__ffs(bits) - 1) : 15u;
                    if (frac == 0u) frac = 1u;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const uint32_t na = (st.s[a] * (16u - frac) + st.e[a] * (frac - 1u)) / 15u;
                        const uint32_t nb = (st.s[a] * (15u - frac) + st.e[a] * frac) / 15u;
                        st.s[a] = na & 0xffffu;
                        st.e[a] = nb & 0xffffu;
                    }
                };
                narrow(s0, b0);
                narrow(s1, b1);
            }
            if (jj == 0) {
                if (v0) for (int a = 0; a < 3; ++a) L->pos[e0][a] = lerp_u16(lo[a], hi[a], ((s0.s[a] + s0.e[a]) / 2u) & 0xffffu);
                if (v1) for (int a = 0; a < 3; ++a) L->pos[e1][a] = lerp_u16(lo[a], hi[a], ((s1.s[a] + s1.e[a]) / 2u) & 0xffffu);
            }
        }
    }
    if (p.stats) {
        if (lane == 0) {
            if (n_empty) atomicAdd(&p.stats[0], n_empty);
            if (n_full) atomicAdd(&p.stats[1], n_full);
            if (n_surf) atomicAdd(&p.stats[2], n_surf);
            if (n_pts) atomicAdd(&p.stats[3], n_pts);
        }
    }
}
void launch_octree_leaf(const OctreeLeafParams& p, int blocks, cudaStream_t s) { k_octree_leaf<<<blocks, 128, 0, s>>>(p); }

// Gradients at the intersections (octree.rs:780-808): one warp per surface leaf, one lane per edge,
// with the tape k_octree_leaf recorded for that leaf.
__global__ void __launch_bounds__(128) k_octree_grads(const __grid_constant__ OctreeLeafParams p) {
    grd slots[REG_SLOTS];
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n = min(*p.n_out, p.cap_out);
    unsigned long long n_pts = 0;
    for (uint32_t i = warp; i < n; i += n_warps) {
        OctreeLeaf* L = p.out + i;
        const TapeRef tr = p.out_tapes[i];
        const uint32_t active = L->present;
        const bool mine = lane < 12 && ((active >> lane) & 1u);
        const int e = mine ? lane : (__ffs(active) - 1);
        grd gx = gr(L->pos[e][0], 1.0f, 0.0f, 0.0f), gy = gr(L->pos[e][1], 0.0f, 1.0f, 0.0f),
            gz = gr(L->pos[e][2], 0.0f, 0.0f, 1.0f);
        if (p.has_transform) xform_gr(p.mat, gx, gy, gz, gx, gy, gz);
        const grd r = run_grad(tr.ptr, tr.n_ops, slots, [&](uint32_t k) {
            return pick_input(p.vb, k, gx, gy, gz, [](float f) { return gr1(f); });
        });
        if (mine) {
            L->grad[e][0] = r.y; L->grad[e][1] = r.z; L->grad[e][2] = r.w; L->grad[e][3] = r.x;
            ++n_pts;
        }
    }
    if (p.stats) {
        for (int o = 16; o > 0; o >>= 1) n_pts += __shfl_xor_sync(FULL, n_pts, o);
        if (lane == 0 && n_pts) atomicAdd(&p.stats[4], n_pts);
    }
}
void launch_octree_grads(const OctreeLeafParams& p, int blocks, cudaStream_t s) { k_octree_grads<<<blocks, 128, 0, s>>>(p); }

}  // namespace fdev
