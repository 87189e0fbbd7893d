// Device-side interpreters and helpers shared by the kernels of the hot path: clause decoding,
// the interval / f32 (two and four points per lane) / gradient tape walks, choice packing and the
// reverse liveness pass that compacts a child tape (VmData::simplify semantics,
// fidget-core/src/vm/data.rs:123-318).
//
// All interpreters keep the tape's VM registers in per-thread local memory (L1-resident,
// lane-interleaved, so a warp's access to one register is one 128/256-byte line) and read tape
// clauses with warp-uniform 8-byte loads.
#pragma once
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

// Clause loads: tapes that were complete before the launch go through the read-only path (ld.global.nc);
// in the fused 2D kernel a tape may have been written by another SM during the same launch, so it is
// read with plain loads (its 128-byte lines are exclusive to it and were never cached before).
template <bool NC>
__device__ __forceinline__ uint2 ld_clause(const uint2* p) {
    return NC ? __ldg(p) : *p;
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
template <bool NC = true, class Input, class Sink, class Out>
__device__ __forceinline__ void run_interval(const uint2* __restrict__ tape, uint32_t n_ops, itv* slots,
                                             Input input, Sink& sink, Out out_fn) {
    if (n_ops == 0) return;
    uint2 w = ld_clause<NC>(tape);
    for (uint32_t i = 0; i < n_ops; ++i) {
        uint2 nxt = ld_clause<NC>(tape + (i + 1 < n_ops ? i + 1 : i));
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

template <bool NC = true, class Input>
__device__ __forceinline__ float2 run_f32x2(const uint2* __restrict__ tape, uint32_t n_ops, float2* slots,
                                            Input input) {
    float2 result = make_float2(nanf_(), nanf_());
    if (n_ops == 0) return result;
    uint2 w = ld_clause<NC>(tape);
    for (uint32_t i = 0; i < n_ops; ++i) {
        uint2 nxt = ld_clause<NC>(tape + (i + 1 < n_ops ? i + 1 : i));
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
template <bool NC = true, class ChoiceSrc>
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
    uint2 nxt = n_ops ? ld_clause<NC>(tape + (n_ops - 1)) : make_uint2(0, 0);
    for (int i = int(n_ops) - 1; i >= 0; --i) {
        const uint2 w = nxt;
        if (i > 0) nxt = ld_clause<NC>(tape + (i - 1));   // prefetch: the clause stream is the latency chain here
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

}  // namespace fdev
