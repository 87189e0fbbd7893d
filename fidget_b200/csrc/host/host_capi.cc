// C ABI of the host-side tape front end; see host_capi.h.
#include "host_capi.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

#include "tape.h"

using namespace fhost;

struct fh_context { Context ctx; };
struct fh_tape { TapeData d; };

static thread_local std::string g_err;
#define FH_TRY(body)                                            \
    try { body; return 0; }                                     \
    catch (const std::exception& e) { g_err = e.what(); return -1; } \
    catch (...) { g_err = "unknown error"; return -1; }

extern "C" {

const char* fh_last_error(void) { return g_err.c_str(); }

int32_t fh_context_new(fh_context** out) { FH_TRY(*out = new fh_context()); }
void fh_context_free(fh_context* ctx) { delete ctx; }

int32_t fh_context_from_text(fh_context* ctx, const char* text, uint32_t* root) {
    FH_TRY(*root = ctx->ctx.from_text(text));
}
int32_t fh_constant(fh_context* ctx, float v, uint32_t* node) { FH_TRY(*node = ctx->ctx.constant(v)); }
int32_t fh_var(fh_context* ctx, int32_t kind, uint32_t* node, uint64_t* var_id) {
    FH_TRY({
        if (kind < 0 || kind > 3) throw std::runtime_error("bad var kind");
        if (kind == 3) {
            *node = ctx->ctx.fresh_var();
            if (var_id) *var_id = ctx->ctx.get(*node).var.id;
        } else {
            *node = ctx->ctx.var(Var{Var::Kind(kind), 0});
            if (var_id) *var_id = 0;
        }
    });
}
static void check(const fh_context* ctx, uint32_t n) {
    if (n >= ctx->ctx.size()) throw std::runtime_error("bad node");
}
int32_t fh_unary(fh_context* ctx, uint8_t opcode, uint32_t a, uint32_t* node) {
    FH_TRY({
        check(ctx, a);
        if (!is_unary(opcode)) throw std::runtime_error("not a unary opcode");
        *node = ctx->ctx.unary(opcode, a);
    });
}
int32_t fh_binary(fh_context* ctx, uint8_t opcode, uint32_t a, uint32_t b, uint32_t* node) {
    FH_TRY({
        check(ctx, a);
        check(ctx, b);
        Context& c = ctx->ctx;
        switch (opcode) {
            case OP_ADD: *node = c.add(a, b); break;
            case OP_SUB: *node = c.sub(a, b); break;
            case OP_MUL: *node = c.mul(a, b); break;
            case OP_DIV: *node = c.div(a, b); break;
            case OP_ATAN2: *node = c.atan2(a, b); break;
            case OP_COMPARE: *node = c.compare(a, b); break;
            case OP_MIX: *node = c.mix(a, b); break;
            case OP_MOD: *node = c.modulo(a, b); break;
            case OP_MIN: *node = c.min(a, b); break;
            case OP_MAX: *node = c.max(a, b); break;
            case OP_AND: *node = c.and_(a, b); break;
            case OP_OR: *node = c.or_(a, b); break;
            default: throw std::runtime_error("not a binary opcode");
        }
    });
}
int32_t fh_context_len(const fh_context* ctx, uint32_t* n) { FH_TRY(*n = uint32_t(ctx->ctx.size())); }

int32_t fh_tape_build(const fh_context* ctx, const uint32_t* roots, uint32_t n_roots, uint32_t n_regs,
                      fh_tape** out) {
    FH_TRY({
        std::vector<Node> r(roots, roots + n_roots);
        for (Node n : r) check(ctx, n);
        *out = new fh_tape{TapeData::build(ctx->ctx, r, n_regs)};
    });
}
void fh_tape_free(fh_tape* t) { delete t; }

int32_t fh_tape_get_info(const fh_tape* t, fh_tape_info* info) {
    FH_TRY({
        info->ssa_len = uint32_t(t->d.ssa.tape.size());
        info->asm_len = uint32_t(t->d.asm_.tape.size());
        info->slot_count = t->d.asm_.slot_count;
        info->choice_count = t->d.ssa.choice_count;
        info->output_count = t->d.ssa.output_count;
        info->n_vars = uint32_t(t->d.vars.size());
        info->n_regs = t->d.n_regs;
        info->var_x = t->d.vars.get(Var{Var::X, 0});
        info->var_y = t->d.vars.get(Var{Var::Y, 0});
        info->var_z = t->d.vars.get(Var{Var::Z, 0});
    });
}
int32_t fh_tape_var(const fh_tape* t, uint32_t i, int32_t* kind, uint64_t* id) {
    FH_TRY({
        if (i >= t->d.vars.size()) throw std::runtime_error("bad var index");
        *kind = int32_t(t->d.vars.order[i].kind);
        *id = t->d.vars.order[i].id;
    });
}

int32_t fh_tape_bytecode(const fh_tape* t, int32_t repack, uint32_t* words, size_t cap, size_t* n_words,
                         uint8_t* reg_count, uint32_t* mem_count) {
    FH_TRY({
        Bytecode bc = make_bytecode(t->d.asm_, t->d.n_regs, repack != 0);
        if (n_words) *n_words = bc.words.size();
        if (reg_count) *reg_count = bc.reg_count;
        if (mem_count) *mem_count = bc.mem_count;
        if (words) {
            if (cap < bc.words.size()) throw std::runtime_error("bytecode buffer too small");
            memcpy(words, bc.words.data(), bc.words.size() * 4);
        }
    });
}

int32_t fh_tape_serialize(const fh_tape* t, uint8_t* buf, size_t cap, size_t* n_bytes) {
    FH_TRY({
        Bytecode bc = make_bytecode(t->d.asm_, t->d.n_regs, true);
        fh_tape_info info;
        if (fh_tape_get_info(t, &info)) throw std::runtime_error("tape info");
        const size_t need = 4 + 6 * 4 + 3 * 4 + 8 + bc.words.size() * 4;
        if (n_bytes) *n_bytes = need;
        if (buf) {
            if (cap < need) throw std::runtime_error("serialize: buffer too small");
            uint8_t* q = buf;
            auto put32 = [&](uint32_t v) { memcpy(q, &v, 4); q += 4; };
            memcpy(q, "FTAP", 4); q += 4;
            put32(1); put32(bc.reg_count); put32(bc.mem_count); put32(info.n_vars); put32(info.output_count);
            put32(info.choice_count);
            put32(uint32_t(info.var_x)); put32(uint32_t(info.var_y)); put32(uint32_t(info.var_z));
            const uint64_t nw = bc.words.size();
            memcpy(q, &nw, 8); q += 8;
            memcpy(q, bc.words.data(), nw * 4);
        }
    });
}

size_t fh_tape_dump(const fh_tape* t, int32_t ssa, char* buf, size_t cap) {
    std::string s;
    if (ssa) {
        for (const auto& c : t->d.ssa.tape) s += to_string(c) + "\n";
    } else {
        for (auto it = t->d.asm_.tape.rbegin(); it != t->d.asm_.tape.rend(); ++it) s += to_string(*it) + "\n";
    }
    if (buf && cap) {
        size_t n = std::min(cap - 1, s.size());
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size() + 1;
}

}  // extern "C"
