// Host orchestration + C ABI of libfidget_cuda (include/fidget_cuda.h).
#include <algorithm>
#include <atomic>
#include <memory>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/fidget_cuda.h"
#include "kernels.cuh"
#include "effects.cuh"

using namespace fdev;

static thread_local std::string g_err;
static int32_t fail(int32_t code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(FC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() { return static_cast<T*>(p); }
};

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Pinned (page-locked, mapped) host memory can be written by kernels directly over PCIe.
// Measured on B200 (profiles/r01_prospero4096.md): SM stores over PCIe reach well under half the
// bandwidth of a DMA copy (3.07 ms vs 2.07 ms end to end for a 67 MB image), so this is opt-in
// (FIDGET_B200_ZEROCOPY=1); the default stages the image in HBM and copies it with the DMA engine.
// Returns the device alias of `p` or null.
void* pinned_device_alias(const void* p) {
    if (!p) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

}  // namespace

struct fc_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr;        // fills are painted here, concurrently with the next levels
    cudaEvent_t ev_fork[MAX_LEVELS] = {}, ev_join = nullptr;
    uint64_t arena_bytes = 1ull << 30;
    // render scratch
    DevBuf arena, jobs[MAX_LEVELS + 1], fills[MAX_LEVELS], choice_scratch, counters, stats, image, heightmap, leaf_tapes, zsort;
    DevBuf fx_in, fx_out, fx_tmp, fx_tables;  // effects: staged host images, intermediate maps, SSAO tables
    std::vector<cudaEvent_t> events;
    std::mutex mu;
    // tape uploads: released device buffers are reused (no cudaMalloc / cudaFree per tape) and the
    // clauses go through a pinned staging buffer with a stream-ordered copy (no host synchronisation)
    std::vector<std::pair<size_t, uint2*>> tape_pool;
    void* stage = nullptr;
    size_t stage_cap = 0;
    cudaEvent_t stage_ev = nullptr;
    struct { size_t smem; int per_sm, threads; } coop_memo[2] = {};   // level-0 launch shape per DIM (occupancy query cached)
    std::shared_ptr<struct Sched> sched_cache[4];
    unsigned sched_next = 0;
};

struct fc_tape {
    fc_ctx* ctx = nullptr;
    std::atomic<int> refs{1};
    uint2* dev = nullptr;
    size_t dev_cap = 0;       // bytes behind `dev` (a pooled buffer may be larger than the tape)
    bool pooled_ok = true;    // false for tapes whose buffer is not a plain cudaMalloc of their own
    std::vector<uint2> host;  // copy of the device clauses
    fc_tape_info info{};
    int ax[3] = {-1, -1, -1};  // input slots of X, Y, Z
    // cooperative level-0 schedule (null when the tape is unsuitable); shared between tapes
    // created from identical bytecode (re-uploading an unchanged shape every frame is the
    // common interactive pattern)
    std::shared_ptr<struct Sched> sched;
};

struct Sched {
    int device = 0;
    uint64_t hash = 0;
    size_t n_clauses = 0;
    CoopRec* d_recs = nullptr;
    CoopFwd* d_fwd = nullptr;
    uint32_t* d_wave_start = nullptr;
    uint32_t n_waves = 0, tail_begin = 0, tail_end = 0, n_slots = 0;
    std::vector<CoopSeg> segs;
    ~Sched() {
        cudaSetDevice(device);
        if (d_recs) cudaFree(d_recs);
        if (d_fwd) cudaFree(d_fwd);
        if (d_wave_start) cudaFree(d_wave_start);
    }
};

// Dependency-wave schedule of a register tape: value id = position of the
// defining clause; clauses are bucketed by dependency depth, sorted by opcode
// inside a bucket (keeps warps convergent); the trailing single-clause
// buckets form the serial tail.
static bool build_schedule(const std::vector<uint2>& cl, std::vector<CoopRec>& recs,
                           std::vector<uint32_t>& wave_start, uint32_t& tail_begin, std::vector<CoopSeg>& segs) {
    const size_t n = cl.size();
    if (n == 0 || n >= COOP_NONE) return false;
    std::vector<int> regdef(256, -1);
    std::vector<uint32_t> depth(n, 0);
    std::vector<CoopRec> byp(n);
    uint32_t ci = 0, max_depth = 0;
    for (size_t p = 0; p < n; ++p) {
        uint32_t x = cl[p].x, dop = x & 0xff, op = dop >> 2, form = dop & 3, out = (x >> 8) & 0xff,
                 lhs = (x >> 16) & 0xff, rhs = x >> 24;
        if (op == OP_MEM) return false;
        CoopRec r;
        r.x = x; r.y = cl[p].y; r.ia = COOP_NONE; r.ib = COOP_NONE; r.p = uint16_t(p); r.cidx = 0;
        bool use_l = false, use_r = false;
        if (op == OP_OUTPUT) use_l = true;
        else if (op == OP_INPUT) {}
        else if (op == OP_COPY) use_l = (form != F_RI);
        else if (op_is_unary(op)) use_l = true;
        else { use_l = (form != F_IR); use_r = (form != F_RI); }
        uint32_t d = 0;
        if (use_l) { if (regdef[lhs] < 0) return false; r.ia = uint16_t(regdef[lhs]); d = std::max(d, depth[r.ia] + 1); }
        if (use_r) { if (regdef[rhs] < 0) return false; r.ib = uint16_t(regdef[rhs]); d = std::max(d, depth[r.ib] + 1); }
        if (op_is_choice(op)) r.cidx = uint16_t(ci++);
        if (ci >= COOP_NONE) return false;
        depth[p] = d;
        max_depth = std::max(max_depth, d);
        if (op != OP_OUTPUT) regdef[out] = int(p);
        byp[p] = r;
    }
    std::vector<std::vector<uint32_t>> levels(max_depth + 1);
    for (size_t p = 0; p < n; ++p) levels[depth[p]].push_back(uint32_t(p));
    size_t first_tail = levels.size();
    while (first_tail > 0 && levels[first_tail - 1].size() == 1) --first_tail;
    recs.clear();
    wave_start.assign(1, 0);
    for (size_t l = 0; l < first_tail; ++l) {
        auto& v = levels[l];
        std::stable_sort(v.begin(), v.end(), [&](uint32_t a, uint32_t b) { return (cl[a].x & 0xff) < (cl[b].x & 0xff); });
        for (uint32_t p : v) recs.push_back(byp[p]);
        wave_start.push_back(uint32_t(recs.size()));
    }
    tail_begin = uint32_t(recs.size());
    for (size_t l = first_tail; l < levels.size(); ++l) recs.push_back(byp[levels[l][0]]);

    // Cut the tail into serial runs and min/max chains (see CoopSeg)
    std::vector<uint32_t> idx_of_pos(n, 0);
    for (size_t i = 0; i < recs.size(); ++i) idx_of_pos[recs[i].p] = uint32_t(i);
    segs.clear();
    const uint32_t tend = uint32_t(recs.size());
    auto chain_op = [&](uint32_t i) -> uint32_t {  // returns opcode if rec i can extend a chain, else 0
        if (i == 0 || i <= tail_begin) return 0;
        const CoopRec& r = recs[i];
        uint32_t dop = r.x & 0xff, op = dop >> 2, form = dop & 3;
        if ((op != OP_MIN && op != OP_MAX) || form != F_RR) return 0;
        uint16_t pp = recs[i - 1].p;
        if ((r.ia == pp) == (r.ib == pp)) return 0;
        return op;
    };
    uint32_t i = tail_begin, serial_start = tail_begin;
    auto flush_serial = [&](uint32_t upto) {
        if (upto > serial_start) segs.push_back(CoopSeg{serial_start, upto, 0, COOP_NONE});
        serial_start = upto;
    };
    while (i < tend) {
        uint32_t op = chain_op(i);
        if (!op) { ++i; continue; }
        uint32_t j = i;
        while (j < tend && chain_op(j) == op) {
            const CoopRec& r = recs[j];
            uint16_t pp = recs[j - 1].p;
            uint16_t side = (r.ia == pp) ? r.ib : r.ia;
            if (idx_of_pos[side] >= i) break;   // the side must be computed before the run starts
            ++j;
        }
        if (j - i >= 8 && segs.size() + 3 <= size_t(COOP_MAX_SEGS)) {
            flush_serial(i);
            segs.push_back(CoopSeg{i, j, 1, COOP_NONE});
            serial_start = j;
            i = j;
        } else {
            i = std::max(j, i + 1);
        }
    }
    flush_serial(tend);
    return segs.size() <= size_t(COOP_MAX_SEGS);
}

// Forward view of a schedule: colour the values with slots.  Execution steps are the waves,
// then every clause of a serial tail run, then each chain run as a whole; a slot is free again
// from the step after the last reader of its value.  A chain value read only by the next clause
// of the same chain needs no slot (the scan never loads it); in a chain clause the operand that
// is the previous chain value is marked COOP_NONE and the run's starting value goes to
// segs[].start_slot.  Returns the number of slots, or 0 if they do not fit 16 bits.
static uint32_t colour_slots(const std::vector<CoopRec>& recs, const std::vector<uint32_t>& wave_start,
                             std::vector<CoopSeg>& segs, std::vector<CoopFwd>& fwd) {
    const size_t m = recs.size();
    std::vector<uint32_t> step(m, 0), idx_of_pos(m, 0), chain_of(m, 0);   // chain_of: 1 + segment index for chain clauses
    uint32_t st = 0;
    for (size_t w = 0; w + 1 < wave_start.size(); ++w, ++st)
        for (uint32_t i = wave_start[w]; i < wave_start[w + 1]; ++i) step[i] = st;
    for (size_t k = 0; k < segs.size(); ++k) {
        if (segs[k].chain) {
            for (uint32_t i = segs[k].begin; i < segs[k].end; ++i) { step[i] = st; chain_of[i] = uint32_t(k) + 1; }
            ++st;
        } else {
            for (uint32_t i = segs[k].begin; i < segs[k].end; ++i) step[i] = st++;
        }
    }
    for (size_t i = 0; i < m; ++i) idx_of_pos[recs[i].p] = uint32_t(i);
    // last reading step of every value and whether anything but its chain successor reads it
    std::vector<uint32_t> last_read(m, 0), n_other(m, 0);
    std::vector<uint8_t> has_reader(m, 0);
    for (size_t i = 0; i < m; ++i) {
        for (uint16_t src : {recs[i].ia, recs[i].ib}) {
            if (src == COOP_NONE) continue;
            const uint32_t d = idx_of_pos[src];
            last_read[d] = std::max(last_read[d], step[i]);
            has_reader[d] = 1;
            const bool chain_succ = chain_of[i] && chain_of[d] == chain_of[i] && d + 1 == i;
            if (!chain_succ) ++n_other[d];
        }
    }
    std::vector<uint16_t> slot(m, uint16_t(COOP_NONE));
    std::vector<uint32_t> free_list;
    std::vector<std::vector<uint32_t>> release(st + 2);   // release[s]: record indices whose slot is free from step s on
    uint32_t n_slots = 0;
    // records sorted by step: waves and tail are already in step order
    uint32_t cur = 0;
    for (size_t i = 0; i < m; ++i) {
        while (cur <= step[i]) {
            for (uint32_t d : release[cur]) free_list.push_back(slot[d]);
            ++cur;
        }
        const bool is_output = ((recs[i].x & 0xff) >> 2) == OP_OUTPUT;
        const bool chain_internal = chain_of[i] && has_reader[i] && n_other[i] == 0 && i + 1 < m && chain_of[i + 1] == chain_of[i];
        if (is_output || chain_internal) continue;
        uint32_t sl;
        if (!free_list.empty()) { sl = free_list.back(); free_list.pop_back(); }
        else sl = n_slots++;
        if (sl >= COOP_NONE) return 0;
        slot[i] = uint16_t(sl);
        const uint32_t rel = (has_reader[i] ? last_read[i] : step[i]) + 1;
        release[std::min<uint32_t>(rel, st + 1)].push_back(uint32_t(i));
    }
    fwd.resize(m);
    for (size_t i = 0; i < m; ++i) {
        CoopFwd f;
        f.x = recs[i].x; f.y = recs[i].y; f.cidx = recs[i].cidx;
        f.sa = recs[i].ia == COOP_NONE ? uint16_t(COOP_NONE) : slot[idx_of_pos[recs[i].ia]];
        f.sb = recs[i].ib == COOP_NONE ? uint16_t(COOP_NONE) : slot[idx_of_pos[recs[i].ib]];
        f.so = slot[i];
        if (chain_of[i]) {
            // the previous chain value is the result of the record right before this one
            const uint16_t prev_pos = recs[i - 1].p;
            CoopSeg& sg = segs[chain_of[i] - 1];
            if (i == sg.begin) sg.start_slot = slot[i - 1];
            if (recs[i].ia == prev_pos) f.sa = uint16_t(COOP_NONE);
            else f.sb = uint16_t(COOP_NONE);
        }
        fwd[i] = f;
    }
    return std::max(n_slots, 1u);
}

// FNV-1a over 64-bit words (one device clause per step)
static uint64_t fnv1a(const uint2* cl, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) {
        h ^= uint64_t(cl[i].x) | uint64_t(cl[i].y) << 32;
        h *= 1099511628211ull;
    }
    return h;
}

static void upload_schedule(fc_tape* t) {
    fc_ctx* c = t->ctx;
    if (t->host.size() < 64) return;   // the cooperative kernel is never used for short tapes
    const uint64_t h = fnv1a(t->host.data(), t->host.size());
    {
        std::lock_guard<std::mutex> g(c->mu);
        for (auto& sp : c->sched_cache)
            if (sp && sp->hash == h && sp->n_clauses == t->host.size()) { t->sched = sp; return; }
    }
    std::vector<CoopRec> recs;
    std::vector<uint32_t> ws;
    uint32_t tb = 0;
    auto sc = std::make_shared<Sched>();
    sc->device = c->device;
    sc->hash = h;
    sc->n_clauses = t->host.size();
    if (!build_schedule(t->host, recs, ws, tb, sc->segs)) return;
    std::vector<CoopFwd> fwd;
    sc->n_slots = colour_slots(recs, ws, sc->segs, fwd);
    if (!sc->n_slots) return;
    if (cudaMalloc(&sc->d_fwd, fwd.size() * sizeof(CoopFwd)) != cudaSuccess) { sc->d_fwd = nullptr; cudaGetLastError(); return; }
    cudaMemcpy(sc->d_fwd, fwd.data(), fwd.size() * sizeof(CoopFwd), cudaMemcpyHostToDevice);
    if (cudaMalloc(&sc->d_recs, recs.size() * sizeof(CoopRec)) != cudaSuccess) { sc->d_recs = nullptr; cudaGetLastError(); return; }
    if (cudaMalloc(&sc->d_wave_start, ws.size() * 4) != cudaSuccess) { sc->d_wave_start = nullptr; cudaGetLastError(); return; }
    cudaMemcpy(sc->d_recs, recs.data(), recs.size() * sizeof(CoopRec), cudaMemcpyHostToDevice);
    cudaMemcpy(sc->d_wave_start, ws.data(), ws.size() * 4, cudaMemcpyHostToDevice);
    sc->n_waves = uint32_t(ws.size() - 1);
    sc->tail_begin = tb;
    sc->tail_end = uint32_t(recs.size());
    t->sched = sc;
    std::lock_guard<std::mutex> g(c->mu);
    c->sched_cache[c->sched_next++ % 4] = sc;
}

struct fc_eval {
    fc_ctx* ctx = nullptr;
    DevBuf in, out, choices, simplify, ptrs, tmp;
};

////////////////////////////////////////////////////////////////////////////
// bytecode <-> device clauses
static int32_t transcode(const uint32_t* words, size_t n_words, uint8_t reg_count, uint32_t mem_count,
                         uint32_t n_vars, uint32_t n_outputs, std::vector<uint2>& out, uint32_t& n_choices) {
    if (!words || n_words < 4 || (n_words & 1)) return fail(FC_ERR_INVALID, "bytecode: bad length");
    if (words[0] != 0xFFFFFFFFu || words[1] != 0u) return fail(FC_ERR_INVALID, "bytecode: missing start marker");
    if (words[n_words - 2] != 0xFFFFFFFFu || words[n_words - 1] != 0xFFFFFFFFu)
        return fail(FC_ERR_INVALID, "bytecode: missing end marker");
    out.clear();
    n_choices = 0;
    auto reg_ok = [&](uint32_t r) { return r < reg_count; };
    for (size_t i = 2; i + 2 < n_words; i += 2) {
        uint32_t w = words[i], imm = words[i + 1];
        uint32_t op = w & 0xff, b1 = (w >> 8) & 0xff, b2 = (w >> 16) & 0xff, b3 = w >> 24;
        if (op >= OP_COUNT) return fail(FC_ERR_INVALID, "bytecode: unknown opcode " + std::to_string(op));
        uint32_t x;
        bool ok = true;
        if (op == OP_OUTPUT) {
            ok = reg_ok(b1) && imm < n_outputs;
            x = enc(op, F_RR, 0xff, b1, 0xff);
        } else if (op == OP_INPUT) {
            ok = reg_ok(b1) && imm < n_vars;
            x = enc(op, F_RR, b1, 0xff, 0xff);
        } else if (op == OP_COPY) {
            if (b2 == 0xff) { ok = reg_ok(b1); x = enc(op, F_RI, b1, 0xff, 0xff); }
            else { ok = reg_ok(b1) && reg_ok(b2); x = enc(op, F_RR, b1, b2, 0xff); }
        } else if (op_is_unary(op)) {
            ok = reg_ok(b1) && reg_ok(b2);
            x = enc(op, F_RR, b1, b2, 0xff);
        } else if (op_is_binary(op)) {
            if (b2 == 0xff && b3 == 0xff) ok = false;
            else if (b2 == 0xff) { ok = reg_ok(b1) && reg_ok(b3); x = enc(op, F_IR, b1, 0xff, b3); }
            else if (b3 == 0xff) { ok = reg_ok(b1) && reg_ok(b2); x = enc(op, F_RI, b1, b2, 0xff); }
            else { ok = reg_ok(b1) && reg_ok(b2) && reg_ok(b3); x = enc(op, F_RR, b1, b2, b3); }
            if (op_is_choice(op)) {
                if (b2 == 0xff) ok = false;  // choice ops never have an immediate lhs (ssa_tape.rs:151-176)
                ++n_choices;
            }
        } else {  // OP_MEM
            if (imm >= mem_count) ok = false;
            else if (b2 == 0xff && b1 != 0xff) { ok = reg_ok(b1); x = enc(op, F_RI, b1, 0xff, 0xff); }
            else if (b1 == 0xff && b2 != 0xff) { ok = reg_ok(b2); x = enc(op, F_IR, 0xff, b2, 0xff); }
            else ok = false;
        }
        if (!ok) return fail(FC_ERR_INVALID, "bytecode: malformed clause at word " + std::to_string(i));
        out.push_back(make_uint2(x, imm));
    }
    return FC_OK;
}

static void to_bytecode(const std::vector<uint2>& cl, std::vector<uint32_t>& words) {
    words.assign({0xFFFFFFFFu, 0u});
    for (const uint2& c : cl) {
        uint32_t dop = c.x & 0xff, op = dop >> 2, form = dop & 3, out = (c.x >> 8) & 0xff, lhs = (c.x >> 16) & 0xff,
                 rhs = c.x >> 24;
        uint32_t b1 = 0xff, b2 = 0xff, b3 = 0xff;
        if (op == OP_OUTPUT) b1 = lhs;
        else if (op == OP_INPUT) b1 = out;
        else if (op == OP_COPY) { b1 = out; if (form != F_RI) b2 = lhs; }
        else if (op_is_unary(op)) { b1 = out; b2 = lhs; }
        else if (op_is_binary(op)) { b1 = out; if (form != F_IR) b2 = lhs; if (form != F_RI) b3 = rhs; }
        else { if (form == F_RI) b1 = out; else b2 = lhs; }
        words.push_back(op | b1 << 8 | b2 << 16 | b3 << 24);
        words.push_back(c.y);
    }
    words.push_back(0xFFFFFFFFu);
    words.push_back(0xFFFFFFFFu);
}

extern "C" {

const char* fc_last_error(void) { return g_err.c_str(); }
uint32_t fc_abi_version(void) { return 1; }

int32_t fc_ctx_create(int32_t device, fc_ctx** out) {
    if (!out) return fail(FC_ERR_INVALID, "null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(FC_ERR_NO_DEVICE, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                          "); libfidget_cuda has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(FC_ERR_INVALID, "bad device index");
    CU(cudaSetDevice(device));
    fc_ctx* c = new fc_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        delete c;
        return fail(FC_ERR_NO_DEVICE, "libfidget_cuda is built for sm_100a only; found sm_" +
                                          std::to_string(prop.major) + std::to_string(prop.minor));
    }
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    for (auto& e : c->ev_fork) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    c->stream = c->own_stream;
    *out = c;
    return FC_OK;
}

void fc_ctx_destroy(fc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->arena.release();
    for (auto& b : c->jobs) b.release();
    for (auto& b : c->fills) b.release();
    c->choice_scratch.release();
    c->counters.release();
    c->stats.release();
    c->image.release();
    c->heightmap.release();
    c->leaf_tapes.release();
    c->zsort.release();
    c->fx_in.release();
    c->fx_out.release();
    c->fx_tmp.release();
    c->fx_tables.release();
    for (auto& pb : c->tape_pool) cudaFree(pb.second);
    if (c->stage) cudaFreeHost(c->stage);
    if (c->stage_ev) cudaEventDestroy(c->stage_ev);
    for (auto ev : c->events) cudaEventDestroy(ev);
    for (auto e : c->ev_fork) if (e) cudaEventDestroy(e);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    cudaStreamDestroy(c->aux_stream);
    cudaStreamDestroy(c->own_stream);
    delete c;
}

int32_t fc_ctx_set_stream(fc_ctx* c, void* s, int32_t use_own) {
    if (!c) return fail(FC_ERR_INVALID, "null ctx");
    cudaStream_t ns = use_own ? c->own_stream : static_cast<cudaStream_t>(s);
    if (ns != c->stream) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);   // tape uploads and scratch reuse are ordered on the context's stream
    }
    c->stream = ns;
    return FC_OK;
}

static int32_t check_device_errors(fc_ctx* c) {
    if (!c->counters.p) return FC_OK;
    Counters h;
    CU(cudaMemcpy(&h, c->counters.p, sizeof h, cudaMemcpyDeviceToHost));
    if (h.error & 1u) return fail(FC_ERR_ARENA, "tape arena exhausted during on-device simplification; raise it with fc_ctx_set_arena_bytes");
    if (h.error & 2u) return fail(FC_ERR_CUDA, "internal work list overflow");
    return FC_OK;
}

int32_t fc_ctx_synchronize(fc_ctx* c) {
    if (!c) return fail(FC_ERR_INVALID, "null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return check_device_errors(c);
}

int32_t fc_ctx_set_arena_bytes(fc_ctx* c, uint64_t bytes) {
    if (!c || bytes < (1u << 20)) return fail(FC_ERR_INVALID, "arena must be at least 1 MiB");
    c->arena_bytes = bytes;
    return FC_OK;
}

////////////////////////////////////////////////////////////////////////////
int32_t fc_tape_create(fc_ctx* c, const uint32_t* words, size_t n_words, uint8_t reg_count, uint32_t mem_count,
                       uint32_t n_vars, uint32_t n_outputs, uint32_t choice_count, fc_tape** out) {
    if (!c || !out) return fail(FC_ERR_INVALID, "null argument");
    if (reg_count == 255) return fail(FC_ERR_INVALID, "register 255 is reserved");
    if (mem_count > 2048 - MEM_BASE) return fail(FC_ERR_UNSUPPORTED, "too many memory slots (max 1792)");
    std::vector<uint2> cl;
    uint32_t nch = 0;
    int32_t rc = transcode(words, n_words, reg_count, mem_count, n_vars, n_outputs, cl, nch);
    if (rc) return rc;
    if (nch != choice_count)
        return fail(FC_ERR_INVALID, "choice_count mismatch: bytecode has " + std::to_string(nch));
    CU(cudaSetDevice(c->device));
    fc_tape* t = new fc_tape();
    t->ctx = c;
    t->host = std::move(cl);
    t->info.n_ops = uint32_t(t->host.size());
    t->info.ref_len = uint32_t(t->host.size());
    t->info.choice_count = nch;
    t->info.reg_count = reg_count;
    t->info.mem_count = mem_count;
    t->info.n_vars = n_vars;
    t->info.n_outputs = n_outputs;
    for (int k = 0; k < 3; ++k) t->ax[k] = uint32_t(k) < n_vars ? k : -1;
    const size_t need = std::max<size_t>(t->host.size(), 1) * sizeof(uint2);
    cudaError_t e = cudaSuccess;
    {
        std::lock_guard<std::mutex> g(c->mu);
        for (size_t k = 0; k < c->tape_pool.size(); ++k)
            if (c->tape_pool[k].first >= need && c->tape_pool[k].first <= 2 * need + 4096) {
                t->dev = c->tape_pool[k].second;
                t->dev_cap = c->tape_pool[k].first;
                c->tape_pool.erase(c->tape_pool.begin() + k);
                break;
            }
        if (!t->dev) {
            e = cudaMalloc(&t->dev, need);
            t->dev_cap = need;
        }
        if (e == cudaSuccess && !t->host.empty()) {
            if (!c->stage_ev) e = cudaEventCreateWithFlags(&c->stage_ev, cudaEventDisableTiming);
            else e = cudaEventSynchronize(c->stage_ev);          // the previous upload has left the staging buffer
            if (e == cudaSuccess && c->stage_cap < need) {
                if (c->stage) cudaFreeHost(c->stage);
                c->stage = nullptr;
                c->stage_cap = 0;
                e = cudaHostAlloc(&c->stage, need * 2, cudaHostAllocDefault);
                if (e == cudaSuccess) c->stage_cap = need * 2;
            }
            if (e == cudaSuccess) {
                memcpy(c->stage, t->host.data(), t->host.size() * sizeof(uint2));
                e = cudaMemcpyAsync(t->dev, c->stage, t->host.size() * sizeof(uint2), cudaMemcpyHostToDevice, c->stream);
            }
            if (e == cudaSuccess) e = cudaEventRecord(c->stage_ev, c->stream);
        }
    }
    if (e != cudaSuccess) {
        if (t->dev) cudaFree(t->dev);
        delete t;
        return fail(FC_ERR_CUDA, cudaGetErrorString(e));
    }
    upload_schedule(t);
    *out = t;
    return FC_OK;
}

int32_t fc_tape_retain(fc_tape* t) {
    if (!t) return fail(FC_ERR_INVALID, "null tape");
    t->refs.fetch_add(1);
    return FC_OK;
}
int32_t fc_tape_release(fc_tape* t) {
    if (!t) return fail(FC_ERR_INVALID, "null tape");
    if (t->refs.fetch_sub(1) == 1) {
        fc_ctx* c = t->ctx;
        cudaSetDevice(c->device);
        bool pooled = false;
        if (t->pooled_ok) {
            // reuse is ordered on the context's stream, behind whatever still reads this tape
            std::lock_guard<std::mutex> g(c->mu);
            if (c->tape_pool.size() < 8) { c->tape_pool.push_back({t->dev_cap, t->dev}); pooled = true; }
        }
        if (!pooled) cudaFree(t->dev);
        delete t;
    }
    return FC_OK;
}
int32_t fc_tape_get_info(const fc_tape* t, fc_tape_info* info) {
    if (!t || !info) return fail(FC_ERR_INVALID, "null argument");
    *info = t->info;
    return FC_OK;
}
int32_t fc_tape_read(const fc_tape* t, uint32_t* words, size_t cap, size_t* n_words) {
    if (!t) return fail(FC_ERR_INVALID, "null tape");
    std::vector<uint32_t> w;
    to_bytecode(t->host, w);
    if (n_words) *n_words = w.size();
    if (words) {
        if (cap < w.size()) return fail(FC_ERR_INVALID, "buffer too small");
        memcpy(words, w.data(), w.size() * 4);
    }
    return FC_OK;
}

////////////////////////////////////////////////////////////////////////////
int32_t fc_eval_create(fc_ctx* c, fc_eval** out) {
    if (!c || !out) return fail(FC_ERR_INVALID, "null argument");
    fc_eval* e = new fc_eval();
    e->ctx = c;
    *out = e;
    return FC_OK;
}
void fc_eval_destroy(fc_eval* e) {
    if (!e) return;
    cudaSetDevice(e->ctx->device);
    cudaStreamSynchronize(e->ctx->stream);
    e->in.release(); e->out.release(); e->choices.release(); e->simplify.release(); e->ptrs.release(); e->tmp.release();
    delete e;
}

static int32_t tracing_eval(fc_eval* e, const fc_tape* t, const float* vars, uint64_t n, float* out, uint8_t* choices,
                            uint8_t* simplify, bool interval) {
    if (!e || !t || (!vars && t->info.n_vars) || !out) return fail(FC_ERR_INVALID, "null argument");
    fc_ctx* c = e->ctx;
    CU(cudaSetDevice(c->device));
    const size_t w = interval ? 2 : 1;
    const size_t in_bytes = n * t->info.n_vars * w * 4, out_bytes = n * t->info.n_outputs * w * 4;
    const size_t ch_bytes = n * t->info.choice_count;
    TracingParams p{};
    p.tape = t->dev;
    p.n_ops = t->info.n_ops;
    p.n_vars = t->info.n_vars;
    p.n_outputs = t->info.n_outputs;
    p.n_choices = t->info.choice_count;
    p.n_slots = MEM_BASE + t->info.mem_count;
    p.n = n;
    const bool dv = is_device_ptr(vars), dout = is_device_ptr(out), dch = is_device_ptr(choices),
               dsi = is_device_ptr(simplify);
    if (dv || !in_bytes) p.vars = vars;
    else {
        CU(e->in.ensure(in_bytes));
        CU(cudaMemcpyAsync(e->in.p, vars, in_bytes, cudaMemcpyHostToDevice, c->stream));
        p.vars = e->in.as<float>();
    }
    if (dout) p.out = out; else { CU(e->out.ensure(std::max<size_t>(out_bytes, 4))); p.out = e->out.as<float>(); }
    if (choices && ch_bytes) {
        if (dch) p.choices = choices; else { CU(e->choices.ensure(ch_bytes)); p.choices = e->choices.as<uint8_t>(); }
    }
    if (simplify) {
        if (dsi) p.simplify = simplify; else { CU(e->simplify.ensure(n)); p.simplify = e->simplify.as<uint8_t>(); }
    }
    if (interval) launch_interval_batch(p, c->stream); else launch_point_batch(p, c->stream);
    CU(cudaGetLastError());
    if (!dout && out_bytes) CU(cudaMemcpyAsync(out, p.out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (choices && ch_bytes && !dch) CU(cudaMemcpyAsync(choices, p.choices, ch_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (simplify && !dsi) CU(cudaMemcpyAsync(simplify, p.simplify, n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return FC_OK;
}

int32_t fc_interval_eval(fc_eval* e, const fc_tape* t, const float* vars, float* out, uint8_t* choices, uint8_t* simplify) {
    return tracing_eval(e, t, vars, 1, out, choices, simplify, true);
}
int32_t fc_point_eval(fc_eval* e, const fc_tape* t, const float* vars, float* out, uint8_t* choices, uint8_t* simplify) {
    return tracing_eval(e, t, vars, 1, out, choices, simplify, false);
}
int32_t fc_interval_eval_batch(fc_eval* e, const fc_tape* t, const float* vars, uint64_t n, float* out, uint8_t* choices,
                               uint8_t* simplify) {
    return tracing_eval(e, t, vars, n, out, choices, simplify, true);
}

static int32_t bulk_eval(fc_eval* e, const fc_tape* t, const void* const* vars, void* const* outs, uint64_t n,
                         size_t elem, bool grad) {
    if (!e || !t || (!vars && t->info.n_vars) || !outs) return fail(FC_ERR_INVALID, "null argument");
    fc_ctx* c = e->ctx;
    CU(cudaSetDevice(c->device));
    const uint32_t nv = t->info.n_vars, no = t->info.n_outputs;
    std::vector<const void*> dptr(nv + no);
    // stage host inputs / outputs in the evaluator's buffers
    size_t in_need = 0, out_need = 0;
    for (uint32_t i = 0; i < nv; ++i) if (!is_device_ptr(vars[i])) in_need += n * elem;
    for (uint32_t o = 0; o < no; ++o) if (!is_device_ptr(outs[o])) out_need += n * elem;
    CU(e->in.ensure(std::max<size_t>(in_need, 16)));
    CU(e->out.ensure(std::max<size_t>(out_need, 16)));
    size_t io = 0, oo = 0;
    std::vector<std::pair<void*, void*>> copy_back;
    for (uint32_t i = 0; i < nv; ++i) {
        if (is_device_ptr(vars[i])) dptr[i] = vars[i];
        else {
            char* d = e->in.as<char>() + io;
            if (n) CU(cudaMemcpyAsync(d, vars[i], n * elem, cudaMemcpyHostToDevice, c->stream));
            dptr[i] = d;
            io += n * elem;
        }
    }
    for (uint32_t o = 0; o < no; ++o) {
        if (is_device_ptr(outs[o])) dptr[nv + o] = outs[o];
        else {
            char* d = e->out.as<char>() + oo;
            dptr[nv + o] = d;
            copy_back.push_back({outs[o], d});
            oo += n * elem;
        }
    }
    CU(e->ptrs.ensure(std::max<size_t>(dptr.size(), 1) * sizeof(void*)));
    if (!dptr.empty())
        CU(cudaMemcpyAsync(e->ptrs.p, dptr.data(), dptr.size() * sizeof(void*), cudaMemcpyHostToDevice, c->stream));
    BulkParams p{};
    p.tape = t->dev;
    p.n_ops = t->info.n_ops;
    p.n_vars = nv;
    p.n_outputs = no;
    p.n_slots = MEM_BASE + t->info.mem_count;
    p.n = n;
    p.vars = e->ptrs.as<const void*>();
    p.outs = reinterpret_cast<void* const*>(e->ptrs.as<void*>() + nv);
    if (grad) launch_grad_slice(p, c->stream); else launch_float_slice(p, c->stream);
    CU(cudaGetLastError());
    for (auto& cb : copy_back)
        if (n) CU(cudaMemcpyAsync(cb.first, cb.second, n * elem, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return FC_OK;
}

int32_t fc_float_slice_eval(fc_eval* e, const fc_tape* t, const float* const* vars, float* const* out, uint64_t n) {
    return bulk_eval(e, t, reinterpret_cast<const void* const*>(vars), reinterpret_cast<void* const*>(out), n, 4, false);
}
int32_t fc_grad_slice_eval(fc_eval* e, const fc_tape* t, const fc_grad* const* vars, fc_grad* const* out, uint64_t n) {
    return bulk_eval(e, t, reinterpret_cast<const void* const*>(vars), reinterpret_cast<void* const*>(out), n, 16, true);
}

int32_t fc_simplify(fc_eval* e, const fc_tape* parent, const uint8_t* choices, size_t n_choices, fc_tape** child) {
    if (!e || !parent || !child || (!choices && n_choices)) return fail(FC_ERR_INVALID, "null argument");
    if (n_choices != parent->info.choice_count)
        return fail(FC_ERR_INVALID, "choice slice length (" + std::to_string(n_choices) + ") does not match choice count (" +
                                        std::to_string(parent->info.choice_count) + ")");
    if (parent->info.mem_count) return fail(FC_ERR_UNSUPPORTED, "fc_simplify: parent tape uses memory slots");
    for (size_t i = 0; i < n_choices; ++i)
        if (choices[i] < 1 || choices[i] > 3) return fail(FC_ERR_INVALID, "trace contains Choice::Unknown");
    fc_ctx* c = e->ctx;
    CU(cudaSetDevice(c->device));
    const uint32_t n = parent->info.n_ops;
    CU(e->tmp.ensure(std::max<size_t>(n, 1) * sizeof(uint2) + 16));
    CU(e->choices.ensure(std::max<size_t>(n_choices, 1)));
    if (n_choices) CU(cudaMemcpyAsync(e->choices.p, choices, n_choices, cudaMemcpyHostToDevice, c->stream));
    SimplifyParams p{};
    p.parent = parent->dev;
    p.n_ops = n;
    p.parent_ref_len = parent->info.ref_len;
    p.choices = e->choices.as<uint8_t>();
    p.n_choices = uint32_t(n_choices);
    p.out = e->tmp.as<uint2>();
    p.result = reinterpret_cast<uint32_t*>(e->tmp.as<char>() + size_t(n) * sizeof(uint2));
    launch_simplify_single(p, c->stream);
    CU(cudaGetLastError());
    uint32_t res[3];
    CU(cudaMemcpyAsync(res, p.result, sizeof res, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    fc_tape* t = new fc_tape();
    t->ctx = c;
    t->info = parent->info;
    memcpy(t->ax, parent->ax, sizeof t->ax);
    t->info.n_ops = res[0];
    t->info.ref_len = res[1];
    t->info.choice_count = res[2];
    t->host.resize(res[0]);
    t->dev_cap = std::max<size_t>(res[0], 1) * sizeof(uint2);
    cudaError_t err = cudaMalloc(&t->dev, t->dev_cap);
    if (err == cudaSuccess && res[0]) {
        err = cudaMemcpy(t->dev, p.out + (n - res[0]), res[0] * sizeof(uint2), cudaMemcpyDeviceToDevice);
        if (err == cudaSuccess) err = cudaMemcpy(t->host.data(), t->dev, res[0] * sizeof(uint2), cudaMemcpyDeviceToHost);
    }
    if (err != cudaSuccess) {
        if (t->dev) cudaFree(t->dev);
        delete t;
        return fail(FC_ERR_CUDA, cudaGetErrorString(err));
    }
    *child = t;
    return FC_OK;
}

////////////////////////////////////////////////////////////////////////////
// Renderers

// TileSizesRef::new (fidget-raster/src/lib.rs:59-66)
static int32_t pick_tile_sizes(const uint32_t* ts_in, uint32_t n_in, const uint32_t* dflt, uint32_t n_dflt,
                               uint32_t max_size, std::vector<uint32_t>& ts) {
    std::vector<uint32_t> all(n_in ? ts_in : dflt, n_in ? ts_in + n_in : dflt + n_dflt);
    if (all.empty() || all.size() > FC_MAX_TILE_LEVELS) return fail(FC_ERR_INVALID, "bad tile size count");
    for (size_t i = 0; i < all.size(); ++i) {
        if (all[i] == 0) return fail(FC_ERR_INVALID, "tile size 0");
        if (i && (all[i - 1] <= all[i] || all[i - 1] % all[i]))
            return fail(FC_ERR_INVALID, "tile sizes must decrease and divide each other");
    }
    size_t pos = all.size();
    for (size_t i = 0; i < all.size(); ++i) if (all[i] < max_size) { pos = i; break; }
    size_t start = pos ? pos - 1 : 0;
    ts.assign(all.begin() + start, all.end());
    return FC_OK;
}

struct AxisMap { int x, y, z; };

// Binds tape input slots to the X, Y, Z axes (ShapeTape::vars(),
// shape/mod.rs:355-376); -1 = axis unused.
int32_t fc_tape_set_axes(fc_tape* t, int32_t x, int32_t y, int32_t z) {
    if (!t) return fail(FC_ERR_INVALID, "null tape");
    int nv = int(t->info.n_vars);
    if (x >= nv || y >= nv || z >= nv) return fail(FC_ERR_INVALID, "axis slot out of range");
    t->ax[0] = x; t->ax[1] = y; t->ax[2] = z;
    return FC_OK;
}

static AxisMap axes_of(const fc_tape* t) { return AxisMap{t->ax[0], t->ax[1], t->ax[2]}; }

// ShapeVars: every non-axis input slot needs a value (MissingVar otherwise, shape/mod.rs:586-600)
static int32_t bind_vars(const fc_tape* t, const float* values, uint32_t n_values, VarBind& vb) {
    AxisMap ax = axes_of(t);
    vb.x = ax.x; vb.y = ax.y; vb.z = ax.z;
    if (t->info.n_vars > uint32_t(MAX_RENDER_VARS)) return fail(FC_ERR_UNSUPPORTED, "renderers support at most 16 input variables");
    for (int i = 0; i < MAX_RENDER_VARS; ++i) vb.values[i] = 0.0f;
    for (uint32_t i = 0; i < t->info.n_vars; ++i) {
        if (int(i) == ax.x || int(i) == ax.y || int(i) == ax.z) continue;
        if (i >= n_values) return fail(FC_ERR_INVALID, "missing value for bound variable in input slot " + std::to_string(i));
        vb.values[i] = values[i];
    }
    return FC_OK;
}

// Attach the tape's wave schedule to a level-0 launch when the cooperative kernel applies
// (long tape, few root tiles per SM); returns the grid size or 0.
static int coop_blocks(fc_ctx* c, const fc_tape* tape, uint64_t n_roots, LevelParams& p, int dim, int& threads) {
    const Sched* sc = tape->sched.get();
    if (!sc || !sc->d_recs || !sc->d_fwd || !sc->d_wave_start || env_int("FIDGET_B200_NO_COOP", 0)) return 0;
    size_t smem = coop_smem_bytes(tape->info.n_ops, tape->info.choice_count, sc->n_slots);
    if (smem > 220 * 1024) return 0;
    // with one lane per tile a warp walks the tape for 32 tiles at once; that only pays when
    // there are enough root tiles to fill the machine several times over
    if (n_roots > uint64_t(c->sm_count) * 32 * 24) return 0;
    p.sched.recs = sc->d_recs;
    p.sched.fwd = sc->d_fwd;
    p.sched.n_slots = sc->n_slots;
    p.sched.wave_start = sc->d_wave_start;
    p.sched.n_waves = sc->n_waves;
    p.sched.tail_begin = sc->tail_begin;
    p.sched.tail_end = sc->tail_end;
    p.sched.n_segs = uint32_t(sc->segs.size());
    for (size_t k = 0; k < sc->segs.size(); ++k) p.sched.segs[k] = sc->segs[k];
    // Tiles are latency chains of ~55 barrier steps: what matters is how many ROUNDS of tiles the
    // launch needs.  Take the fewest CTAs per SM that reach the minimal number of rounds (wider CTAs
    // shorten a tile), within shared memory (1 KB reserved + ~1.5 KB static per CTA), 2048 threads
    // and the register file.
    const int max_by_smem = int(std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (smem + 2560))));
    const int cap = std::min(max_by_smem, env_int("FIDGET_B200_COOP_PER_SM", 8));
    auto rounds = [&](int per_sm) { return (n_roots + uint64_t(c->sm_count) * per_sm - 1) / (uint64_t(c->sm_count) * per_sm); };
    int per_sm = 1;
    for (int k = 1; k <= cap; ++k) if (rounds(k) < rounds(per_sm)) per_sm = k;
    // widest CTA for which the runtime really keeps per_sm of them resident (register granularity
    // makes 7 x 224 threads x 40 registers NOT fit although 7 * 224 * 40 < 64 K)
    auto& mm = c->coop_memo[dim == 3];
    if (mm.threads == 0 || mm.smem != smem || mm.per_sm != per_sm) {
        int t = COOP_THREADS;
        while (t > 64 && coop_occupancy(dim, t, smem) < per_sm) t -= 32;
        mm = {smem, per_sm, t};
    }
    threads = mm.threads;
    threads = env_int("FIDGET_B200_COOP_THREADS", threads);
    if (env_int("FIDGET_B200_COOP_DEBUG", 0))
        fprintf(stderr, "coop: %u clauses, %u slots, %zu B smem, %d CTAs/SM x %d threads (%d regs), %llu roots, occupancy %d CTAs/SM\n",
                tape->info.n_ops, sc->n_slots, smem, per_sm, threads, coop_regs_per_thread(dim), (unsigned long long)n_roots,
                coop_occupancy(dim, threads, smem));
    return int(std::max<uint64_t>(1, std::min<uint64_t>(n_roots, uint64_t(c->sm_count) * per_sm)));
}

static cudaEvent_t get_event(fc_ctx* c, size_t i) {
    while (c->events.size() <= i) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        c->events.push_back(ev);
    }
    return c->events[i];
}

int32_t fc_render2d(fc_ctx* c, const fc_tape* tape, const fc_render2d_cfg* cfg, float* out, fc_render_stats* stats) {
    if (!c || !tape || !cfg || !out) return fail(FC_ERR_INVALID, "null argument");
    if (cfg->width == 0 || cfg->height == 0) return fail(FC_ERR_INVALID, "empty image");
    if (tape->info.mem_count) return fail(FC_ERR_UNSUPPORTED, "renderers need a tape without memory spills (<= 254 registers)");
    if (tape->info.n_outputs != 1) return fail(FC_ERR_INVALID, "ShapeTape has multiple outputs");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    static const uint32_t DFLT[3] = {128, 32, 8};
    std::vector<uint32_t> ts;
    int32_t rc = pick_tile_sizes(cfg->tile_sizes, cfg->n_tile_sizes, DFLT, 3, std::max(cfg->width, cfg->height), ts);
    if (rc) return rc;
    const int L = int(ts.size());
    const uint32_t T0 = ts[0];
    const uint32_t roots_x = (cfg->width + T0 - 1) / T0;
    uint32_t roots_y_all = (cfg->height + T0 - 1) / T0;
    uint32_t row0 = cfg->root_row_begin, row1 = cfg->root_row_end ? cfg->root_row_end : roots_y_all;
    if (row0 > row1 || row1 > roots_y_all) return fail(FC_ERR_INVALID, "bad root row band");
    const uint32_t roots_y = row1 - row0;
    const uint64_t n_roots = uint64_t(roots_x) * roots_y;
    const bool timing = (cfg->flags & FC_FLAG_TIMING) != 0;
    const bool async = (cfg->flags & FC_FLAG_ASYNC) != 0;
    const bool want_stats = stats != nullptr;
    cudaStream_t s = c->stream;
    const bool serial_fill = env_int("FIDGET_B200_SERIAL_FILL", 0) != 0;

    // ---- scratch ----
    const int bps = env_int("FIDGET_B200_BLOCKS_PER_SM", 6);
    const int grid_blocks = c->sm_count * bps;
    const uint32_t choice_words = (tape->info.choice_count + 15) / 16 + 1;
    CU(c->choice_scratch.ensure(size_t(grid_blocks) * WARPS_PER_BLOCK * choice_words * 32 * 4));
    CU(c->arena.ensure(c->arena_bytes));
    CU(c->counters.ensure(sizeof(Counters)));
    CU(c->stats.ensure(sizeof(Stats)));
    std::vector<uint64_t> level_tiles(L + 1);
    for (int l = 1; l <= L; ++l) {
        // jobs queued for level l are tiles of size ts[l-1]
        uint64_t per_root = uint64_t(T0 / ts[l - 1]) * (T0 / ts[l - 1]);
        level_tiles[l] = n_roots * per_root;
        if (level_tiles[l] > 0xfffffff0ull) return fail(FC_ERR_UNSUPPORTED, "image too large for 32-bit tile lists");
        CU(c->jobs[l].ensure(level_tiles[l] * sizeof(TileJob)));
        CU(c->fills[l - 1].ensure(level_tiles[l] * sizeof(FillRec)));
    }
    bool out_dev = is_device_ptr(out);
    float* dimg = out;
    const size_t img_bytes = size_t(cfg->width) * cfg->height * 4;
    if (!out_dev) {
        void* alias = env_int("FIDGET_B200_ZEROCOPY", 0) ? pinned_device_alias(out) : nullptr;
        if (alias) {
            dimg = static_cast<float*>(alias);   // zero-copy: kernels store straight into the host image
            out_dev = true;
        } else {
            CU(c->image.ensure(img_bytes));
            dimg = c->image.as<float>();
        }
    }
    CU(cudaMemsetAsync(c->counters.p, 0, sizeof(Counters), s));
    if (want_stats) CU(cudaMemsetAsync(c->stats.p, 0, sizeof(Stats), s));
    // (pixels outside the requested band of root rows are left untouched)

    VarBind vb;
    if (int32_t vrc = bind_vars(tape, cfg->var_values, cfg->n_var_values, vb)) return vrc;
    size_t ev = 0;
    if (timing) CU(cudaEventRecord(get_event(c, ev++), s));
    uint32_t launches = 0;
    for (int l = 0; l < L; ++l) {
        LevelParams p{};
        p.level = l;
        p.tile = ts[l];
        p.n_axis = l ? ts[l - 1] / ts[l] : 0;
        p.is_last = (l == L - 1);
        p.pixel_perfect = cfg->pixel_perfect;
        p.root_mode = (l == 0);
        p.roots_x = roots_x; p.roots_y = roots_y; p.roots_z = 1;
        p.root_x0 = 0; p.root_y0 = row0 * T0; p.root_z0 = 0;
        p.root_tape.ptr = tape->dev;
        p.root_tape.n_ops = tape->info.n_ops;
        p.root_tape.ref_len = tape->info.ref_len;
        p.root_tape.n_choices = tape->info.choice_count;
        p.width = cfg->width; p.height = cfg->height; p.depth = 1;
        p.z2d = cfg->z;
        memcpy(p.mat.m, cfg->mat, sizeof p.mat.m);
        p.jobs_in = l ? c->jobs[l].as<TileJob>() : nullptr;
        p.cap_in = l ? uint32_t(level_tiles[l]) : 0;
        p.jobs_out = c->jobs[l + 1].as<TileJob>();
        p.cap_out = uint32_t(level_tiles[l + 1]);
        p.fills = c->fills[l].as<FillRec>();
        p.cap_fills = uint32_t(level_tiles[l + 1]);
        p.arena = c->arena.as<uint2>();
        p.arena_cap = std::min<uint64_t>(c->arena.cap, c->arena_bytes) / sizeof(uint2);
        p.choice_scratch = c->choice_scratch.as<uint32_t>();
        p.choice_words = choice_words;
        p.ctr = c->counters.as<Counters>();
        p.stats = want_stats ? c->stats.as<Stats>() : nullptr;
        p.vb = vb;
        int blocks = grid_blocks;
        if (l == 0) {
            uint64_t warps = (n_roots + 31) / 32;
            blocks = int(std::min<uint64_t>((warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, uint64_t(grid_blocks)));
        }
        bool coop = false;
        if (l == 0) {
            int ct = COOP_THREADS;
            int cb = coop_blocks(c, tape, n_roots, p, 2, ct);
            if (cb > 0) {
                CU(launch_interval_root_coop_2d(p, cb, ct, s));
                coop = true;
            }
        }
        if (!coop) launch_interval_level_2d(p, std::max(blocks, 1), s);
        ++launches;
        if (timing) CU(cudaEventRecord(get_event(c, ev++), s));
        {
            // the tiles this level proved inside/outside are painted on a second stream while the
            // next (latency-bound) levels run: fills and leaf pixels never touch the same pixel
            FillParams f{};
            f.tile = ts[l];
            f.width = cfg->width; f.height = cfg->height;
            f.fills = c->fills[l].as<FillRec>();
            f.n_fills = &c->counters.as<Counters>()->n_fills[l];
            f.out = dimg;
            cudaStream_t fs = serial_fill ? s : c->aux_stream;
            if (!serial_fill) {
                CU(cudaEventRecord(c->ev_fork[l], s));
                CU(cudaStreamWaitEvent(c->aux_stream, c->ev_fork[l], 0));
            }
            launch_fill_2d(f, c->sm_count * 2, fs);
            ++launches;
        }
    }
    if (timing) CU(cudaEventRecord(get_event(c, ev++), s));
    {
        PixelParams q{};
        q.tile = ts[L - 1];
        q.width = cfg->width; q.height = cfg->height;
        q.z2d = cfg->z;
        memcpy(q.mat.m, cfg->mat, sizeof q.mat.m);
        q.jobs = c->jobs[L].as<TileJob>();
        q.out = dimg;
        q.ctr = c->counters.as<Counters>();
        q.list = L;
        q.cursor = L;
        q.stats = want_stats ? c->stats.as<Stats>() : nullptr;
        q.vb = vb;
        launch_pixels_2d(q, c->sm_count * env_int("FIDGET_B200_PIXEL_BLOCKS_PER_SM", 8), s);
        ++launches;
    }
    if (!serial_fill) {
        CU(cudaEventRecord(c->ev_join, c->aux_stream));
        CU(cudaStreamWaitEvent(s, c->ev_join, 0));
    }
    if (timing) CU(cudaEventRecord(get_event(c, ev++), s));
    CU(cudaGetLastError());
    if (!out_dev) CU(cudaMemcpyAsync(out, dimg, img_bytes, cudaMemcpyDeviceToHost, s));
    if (async && out_dev && !want_stats) return FC_OK;
    CU(cudaStreamSynchronize(s));
    rc = check_device_errors(c);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        Stats h;
        Counters hc;
        CU(cudaMemcpy(&h, c->stats.p, sizeof h, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&hc, c->counters.p, sizeof hc, cudaMemcpyDeviceToHost));
        for (int l = 0; l < FC_MAX_TILE_LEVELS; ++l) {
            stats->evaluated[l] = h.evaluated[l];
            stats->filled_inside[l] = h.filled_inside[l];
            stats->filled_outside[l] = h.filled_outside[l];
            stats->ambiguous[l] = h.ambiguous[l];
            stats->simplified[l] = h.simplified[l];
        }
        stats->pixels = h.pixels;
        stats->arena_bytes_used = hc.arena_top * sizeof(uint2);
        stats->kernel_launches = launches;
        if (timing) {
            float ms = 0;
            for (int l = 0; l < L; ++l) {
                cudaEventElapsedTime(&ms, c->events[l], c->events[l + 1]);
                stats->stage_ms[l] = ms;
            }
            cudaEventElapsedTime(&ms, c->events[L], c->events[L + 1]);
            stats->stage_ms[8] = ms;
            cudaEventElapsedTime(&ms, c->events[L + 1], c->events[L + 2]);
            stats->stage_ms[9] = ms;
            cudaEventElapsedTime(&ms, c->events[0], c->events[L + 2]);
            stats->stage_ms[15] = ms;
        }
    }
    return rc;
}

int32_t fc_render3d(fc_ctx* c, const fc_tape* tape, const fc_render3d_cfg* cfg, fc_geometry_pixel* out,
                    fc_render_stats* stats) {
    if (!c || !tape || !cfg || !out) return fail(FC_ERR_INVALID, "null argument");
    if (cfg->width == 0 || cfg->height == 0 || cfg->depth == 0) return fail(FC_ERR_INVALID, "empty volume");
    if (tape->info.mem_count) return fail(FC_ERR_UNSUPPORTED, "renderers need a tape without memory spills (<= 254 registers)");
    if (tape->info.n_outputs != 1) return fail(FC_ERR_INVALID, "ShapeTape has multiple outputs");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    static const uint32_t DFLT[5] = {128, 64, 32, 16, 8};
    std::vector<uint32_t> ts;
    int32_t rc = pick_tile_sizes(cfg->tile_sizes, cfg->n_tile_sizes, DFLT, 5, std::max(cfg->width, cfg->height), ts);
    if (rc) return rc;
    const int L = int(ts.size());
    const uint32_t T0 = ts[0];
    const uint32_t roots_x = (cfg->width + T0 - 1) / T0, roots_y = (cfg->height + T0 - 1) / T0;
    const uint32_t z_begin = cfg->z_begin, z_end = cfg->z_end ? cfg->z_end : cfg->depth;
    if (z_begin % T0 || z_begin >= z_end || z_end > ((cfg->depth + T0 - 1) / T0) * T0)
        return fail(FC_ERR_INVALID, "z slab must start on a root-tile boundary inside the volume");
    const uint32_t roots_z = (std::min(z_end, ((cfg->depth + T0 - 1) / T0) * T0) - z_begin + T0 - 1) / T0;
    const uint64_t n_roots = uint64_t(roots_x) * roots_y * roots_z;
    if (n_roots > 0xfffffff0ull) return fail(FC_ERR_UNSUPPORTED, "volume too large");
    const bool timing = (cfg->flags & FC_FLAG_TIMING) != 0;
    const bool async = (cfg->flags & FC_FLAG_ASYNC) != 0;
    const bool want_stats = stats != nullptr;
    cudaStream_t s = c->stream;

    const int bps = env_int("FIDGET_B200_BLOCKS_PER_SM", 6);
    const int grid_blocks = c->sm_count * bps;
    const uint32_t choice_words = (tape->info.choice_count + 15) / 16 + 1;
    CU(c->choice_scratch.ensure(size_t(grid_blocks) * WARPS_PER_BLOCK * choice_words * 32 * 4));
    CU(c->arena.ensure(c->arena_bytes));
    CU(c->counters.ensure(sizeof(Counters)));
    CU(c->stats.ensure(sizeof(Stats)));
    // Work lists hold only ambiguous tiles (a surface-like set), so they are
    // capped well below the N^3 tile count; overflow is reported, not ignored.
    const uint64_t cap_limit = uint64_t(env_int("FIDGET_B200_MAX_TILES_M", 16)) << 20;
    std::vector<uint64_t> level_cap(L + 1);
    for (int l = 1; l <= L; ++l) {
        uint64_t r = T0 / ts[l - 1];
        level_cap[l] = std::min<uint64_t>(n_roots * r * r * r, cap_limit);
        CU(c->jobs[l].ensure(level_cap[l] * sizeof(TileJob)));
    }
    const size_t npix = size_t(cfg->width) * cfg->height;
    CU(c->heightmap.ensure(npix * 8));
    bool out_dev = is_device_ptr(out);
    void* dimg = out;
    if (!out_dev) {
        void* alias = env_int("FIDGET_B200_ZEROCOPY", 0) ? pinned_device_alias(out) : nullptr;
        if (alias) {
            dimg = alias;
            out_dev = true;
        } else {
            CU(c->image.ensure(npix * 16));
            dimg = c->image.p;
        }
    }
    CU(cudaMemsetAsync(c->counters.p, 0, sizeof(Counters), s));
    CU(cudaMemsetAsync(c->heightmap.p, 0, npix * 8, s));
    if (want_stats) CU(cudaMemsetAsync(c->stats.p, 0, sizeof(Stats), s));

    VarBind vb;
    if (int32_t vrc = bind_vars(tape, cfg->var_values, cfg->n_var_values, vb)) return vrc;
    size_t ev = 0;
    if (timing) CU(cudaEventRecord(get_event(c, ev++), s));
    uint32_t launches = 0;
    for (int l = 0; l < L; ++l) {
        LevelParams p{};
        p.level = l;
        p.tile = ts[l];
        p.n_axis = l ? ts[l - 1] / ts[l] : 0;
        p.is_last = (l == L - 1);
        p.pixel_perfect = 0;
        p.root_mode = (l == 0);
        p.roots_x = roots_x; p.roots_y = roots_y; p.roots_z = roots_z;
        p.root_x0 = 0; p.root_y0 = 0; p.root_z0 = z_begin;
        p.root_tape.ptr = tape->dev;
        p.root_tape.n_ops = tape->info.n_ops;
        p.root_tape.ref_len = tape->info.ref_len;
        p.root_tape.n_choices = tape->info.choice_count;
        p.width = cfg->width; p.height = cfg->height; p.depth = cfg->depth;
        memcpy(p.mat.m, cfg->mat, sizeof p.mat.m);
        p.jobs_in = l ? c->jobs[l].as<TileJob>() : nullptr;
        p.cap_in = l ? uint32_t(level_cap[l]) : 0;
        p.jobs_out = c->jobs[l + 1].as<TileJob>();
        p.cap_out = uint32_t(level_cap[l + 1]);
        p.arena = c->arena.as<uint2>();
        p.arena_cap = std::min<uint64_t>(c->arena.cap, c->arena_bytes) / sizeof(uint2);
        p.choice_scratch = c->choice_scratch.as<uint32_t>();
        p.choice_words = choice_words;
        p.ctr = c->counters.as<Counters>();
        p.stats = want_stats ? c->stats.as<Stats>() : nullptr;
        p.heightmap = c->heightmap.as<unsigned long long>();
        p.vb = vb;
        int blocks = grid_blocks;
        if (l == 0) {
            uint64_t warps = (n_roots + 31) / 32;
            blocks = int(std::min<uint64_t>((warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, uint64_t(grid_blocks)));
        }
        bool coop = false;
        if (l == 0) {
            int ct = COOP_THREADS;
            int cb = coop_blocks(c, tape, n_roots, p, 3, ct);
            if (cb > 0) {
                CU(launch_interval_root_coop_3d(p, cb, ct, s));
                coop = true;
            }
        }
        if (!coop) launch_interval_level_3d(p, std::max(blocks, 1), s);
        ++launches;
        if (timing) CU(cudaEventRecord(get_event(c, ev++), s));
    }
    {
        VoxelParams q{};
        q.tile = ts[L - 1];
        if (!env_int("FIDGET_B200_NO_ZSORT", 0)) {
            const uint32_t n_layers = (roots_z * T0) / ts[L - 1];
            CU(c->zsort.ensure(size_t(n_layers + 1) * 4 + level_cap[L] * 4));
            uint32_t* hist = c->zsort.as<uint32_t>();
            uint32_t* order = hist + n_layers + 1;
            launch_leaf_zsort(c->jobs[L].as<TileJob>(), &c->counters.as<Counters>()->n_jobs[L], uint32_t(level_cap[L]),
                              z_begin, ts[L - 1], n_layers, hist, order, s);
            launches += 3;
            q.order = order;
        }
        q.width = cfg->width; q.height = cfg->height;
        memcpy(q.mat.m, cfg->mat, sizeof q.mat.m);
        q.jobs = c->jobs[L].as<TileJob>();
        q.cap_jobs = uint32_t(level_cap[L]);
        q.heightmap = c->heightmap.as<unsigned long long>();
        q.ctr = c->counters.as<Counters>();
        q.list = L; q.cursor = L;
        q.stats = want_stats ? c->stats.as<Stats>() : nullptr;
        q.vb = vb;
        launch_voxels_3d(q, c->sm_count * env_int("FIDGET_B200_PIXEL_BLOCKS_PER_SM", 8), s);
        ++launches;
    }
    if (timing) CU(cudaEventRecord(get_event(c, ev++), s));
    {
        NormalParams q{};
        q.width = cfg->width; q.height = cfg->height; q.depth = cfg->depth;
        q.clamp = (cfg->flags & FC_FLAG_NO_CLAMP) ? 0 : 1;
        memcpy(q.mat.m, cfg->mat, sizeof q.mat.m);
        q.jobs = c->jobs[L].as<TileJob>();
        q.heightmap = c->heightmap.as<unsigned long long>();
        q.out = dimg;
        q.stats = want_stats ? c->stats.as<Stats>() : nullptr;
        q.vb = vb;
        launch_normals_3d(q, s);
        ++launches;
    }
    if (timing) CU(cudaEventRecord(get_event(c, ev++), s));
    CU(cudaGetLastError());
    if (!out_dev) CU(cudaMemcpyAsync(out, dimg, npix * 16, cudaMemcpyDeviceToHost, s));
    if (async && out_dev && !want_stats) return FC_OK;
    CU(cudaStreamSynchronize(s));
    rc = check_device_errors(c);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        Stats h;
        Counters hc;
        CU(cudaMemcpy(&h, c->stats.p, sizeof h, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&hc, c->counters.p, sizeof hc, cudaMemcpyDeviceToHost));
        for (int l = 0; l < FC_MAX_TILE_LEVELS; ++l) {
            stats->evaluated[l] = h.evaluated[l];
            stats->filled_inside[l] = h.filled_inside[l];
            stats->filled_outside[l] = h.filled_outside[l];
            stats->ambiguous[l] = h.ambiguous[l];
            stats->simplified[l] = h.simplified[l];
        }
        stats->pixels = h.pixels;
        stats->grads = h.grads;
        stats->arena_bytes_used = hc.arena_top * sizeof(uint2);
        stats->kernel_launches = launches;
        if (timing) {
            float ms = 0;
            for (int l = 0; l < L; ++l) {
                cudaEventElapsedTime(&ms, c->events[l], c->events[l + 1]);
                stats->stage_ms[l] = ms;
            }
            cudaEventElapsedTime(&ms, c->events[L], c->events[L + 1]);
            stats->stage_ms[9] = ms;
            cudaEventElapsedTime(&ms, c->events[L + 1], c->events[L + 2]);
            stats->stage_ms[10] = ms;
            cudaEventElapsedTime(&ms, c->events[0], c->events[L + 2]);
            stats->stage_ms[15] = ms;
        }
    }
    return rc;
}

int32_t fc_merge_slabs(fc_ctx* c, const fc_geometry_pixel* const* slabs, uint32_t n_slabs, uint32_t width,
                       uint32_t height, uint32_t depth, fc_geometry_pixel* out) {
    if (!c || !slabs || !n_slabs || !out) return fail(FC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    for (uint32_t i = 0; i < n_slabs; ++i)
        if (!is_device_ptr(slabs[i])) return fail(FC_ERR_INVALID, "fc_merge_slabs takes device pointers");
    if (!is_device_ptr(out)) return fail(FC_ERR_INVALID, "fc_merge_slabs takes device pointers");
    CU(c->image.ensure(std::max<size_t>(n_slabs * sizeof(void*), 16)));
    CU(cudaMemcpyAsync(c->image.p, slabs, n_slabs * sizeof(void*), cudaMemcpyHostToDevice, c->stream));
    launch_merge_slabs(c->image.as<const void*>(), n_slabs, width * height, depth, out, c->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return FC_OK;
}

int32_t fc_octree_sample(fc_ctx* c, const fc_tape* tape, const fc_octree_cfg* cfg, fc_octree_leaf* out, uint64_t cap,
                         uint64_t* n_leaves, fc_octree_stats* stats) {
    static_assert(sizeof(fc_octree_leaf) == sizeof(OctreeLeaf) && sizeof(OctreeLeaf) == 348, "leaf layout");
    if (!c || !tape || !cfg || !n_leaves || (!out && cap)) return fail(FC_ERR_INVALID, "null argument");
    if (cfg->depth > FC_MAX_OCTREE_DEPTH) return fail(FC_ERR_INVALID, "octree depth too large");
    if (tape->info.mem_count) return fail(FC_ERR_UNSUPPORTED, "the octree sampler needs a tape without memory spills");
    if (tape->info.n_outputs != 1) return fail(FC_ERR_INVALID, "ShapeTape has multiple outputs");
    if (cap > 0xfffffff0ull) return fail(FC_ERR_INVALID, "leaf capacity too large");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    VarBind vb;
    if (int32_t vrc = bind_vars(tape, cfg->var_values, cfg->n_var_values, vb)) return vrc;
    const uint32_t D = cfg->depth;
    const int L = int(D) + 1;   // interval levels: depth 0 (the root cell) .. D
    cudaStream_t s = c->stream;
    const bool timing = (cfg->flags & FC_FLAG_TIMING) != 0;
    const int bps = env_int("FIDGET_B200_BLOCKS_PER_SM", 6);
    const int grid_blocks = c->sm_count * bps;
    const uint32_t choice_words = (tape->info.choice_count + 15) / 16 + 1;
    CU(c->choice_scratch.ensure(size_t(grid_blocks) * WARPS_PER_BLOCK * choice_words * 32 * 4));
    CU(c->arena.ensure(c->arena_bytes));
    CU(c->counters.ensure(sizeof(Counters) + 64));
    CU(c->stats.ensure(sizeof(Stats)));
    const uint64_t cap_limit = uint64_t(env_int("FIDGET_B200_MAX_TILES_M", 16)) << 20;
    std::vector<uint64_t> level_cap(L + 1);
    for (int l = 1; l <= L; ++l) {
        uint64_t cells = 1ull << (3 * std::min(l, int(D)));   // cells at depth l (the leaf list holds depth-D cells)
        level_cap[l] = std::min<uint64_t>(cells, cap_limit);
        CU(c->jobs[l].ensure(level_cap[l] * sizeof(TileJob)));
    }
    const bool out_dev = is_device_ptr(out);
    OctreeLeaf* dout = reinterpret_cast<OctreeLeaf*>(out);
    if (!out_dev) {
        CU(c->image.ensure(std::max<uint64_t>(cap, 1) * sizeof(OctreeLeaf)));
        dout = c->image.as<OctreeLeaf>();
    }
    CU(c->leaf_tapes.ensure(std::max<uint64_t>(cap, 1) * sizeof(TapeRef)));
    CU(cudaMemsetAsync(c->counters.p, 0, sizeof(Counters) + 64, s));
    CU(cudaMemsetAsync(c->stats.p, 0, sizeof(Stats), s));
    // extra device words after Counters: [0] n_out, then 5 u64 leaf statistics (8-byte aligned)
    uint32_t* d_n_out = reinterpret_cast<uint32_t*>(c->counters.as<char>() + sizeof(Counters));
    unsigned long long* d_leaf_stats = reinterpret_cast<unsigned long long*>(c->counters.as<char>() + sizeof(Counters) + 8);
    if (timing) CU(cudaEventRecord(get_event(c, 0), s));
    uint32_t launches = 0;
    for (int l = 0; l < L; ++l) {
        LevelParams p{};
        p.level = l;
        p.tile = 1u << (D - uint32_t(l));
        p.n_axis = l ? 2 : 0;
        p.is_last = (l == L - 1);
        p.root_mode = (l == 0);
        p.roots_x = p.roots_y = p.roots_z = 1;
        p.root_tape.ptr = tape->dev;
        p.root_tape.n_ops = tape->info.n_ops;
        p.root_tape.ref_len = tape->info.ref_len;
        p.root_tape.n_choices = tape->info.choice_count;
        p.width = p.height = p.depth = 1u << D;
        memcpy(p.mat.m, cfg->world_to_model, sizeof p.mat.m);
        p.jobs_in = l ? c->jobs[l].as<TileJob>() : nullptr;
        p.cap_in = l ? uint32_t(level_cap[l]) : 0;
        p.jobs_out = c->jobs[l + 1].as<TileJob>();
        p.cap_out = uint32_t(level_cap[l + 1]);
        p.arena = c->arena.as<uint2>();
        p.arena_cap = std::min<uint64_t>(c->arena.cap, c->arena_bytes) / sizeof(uint2);
        p.choice_scratch = c->choice_scratch.as<uint32_t>();
        p.choice_words = choice_words;
        p.ctr = c->counters.as<Counters>();
        p.stats = c->stats.as<Stats>();
        p.mode = 1;
        p.has_transform = cfg->has_transform;
        p.cell_h = 2.0f / float(1u << D);
        p.vb = vb;
        uint64_t cells = 1ull << (3 * l);
        uint64_t warps = l ? std::max<uint64_t>(1, cells / 8) : 1;
        int blocks = int(std::min<uint64_t>((warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, uint64_t(grid_blocks)));
        launch_interval_level_3d(p, std::max(blocks, 1), s);
        ++launches;
    }
    OctreeLeafParams q{};
    q.jobs = c->jobs[L].as<TileJob>();
    q.cap_jobs = uint32_t(level_cap[L]);
    q.ctr = c->counters.as<Counters>();
    q.list = L; q.cursor = L;
    q.cell_h = 2.0f / float(1u << D);
    q.has_transform = cfg->has_transform;
    memcpy(q.mat.m, cfg->world_to_model, sizeof q.mat.m);
    q.vb = vb;
    q.out = dout;
    q.out_tapes = c->leaf_tapes.as<TapeRef>();
    q.cap_out = uint32_t(cap);
    q.n_out = d_n_out;
    q.stats = d_leaf_stats;
    launch_octree_leaf(q, c->sm_count * 8, s);
    launch_octree_grads(q, c->sm_count * 8, s);
    launches += 2;
    if (timing) CU(cudaEventRecord(get_event(c, 1), s));
    CU(cudaGetLastError());
    uint32_t n_out = 0;
    CU(cudaMemcpyAsync(&n_out, d_n_out, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    *n_leaves = n_out;
    int32_t rc = check_device_errors(c);
    if (n_out > cap) rc = fail(FC_ERR_INVALID, "leaf buffer too small: " + std::to_string(n_out) + " surface leaves");
    if (!rc && !out_dev && n_out) CU(cudaMemcpy(out, dout, size_t(n_out) * sizeof(OctreeLeaf), cudaMemcpyDeviceToHost));
    if (stats) {
        memset(stats, 0, sizeof *stats);
        Stats h;
        Counters hc;
        unsigned long long ls[5];
        CU(cudaMemcpy(&h, c->stats.p, sizeof h, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&hc, c->counters.p, sizeof hc, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(ls, d_leaf_stats, sizeof ls, cudaMemcpyDeviceToHost));
        for (int l = 0; l < 16 && l < MAX_LEVELS; ++l) {
            stats->evaluated[l] = h.evaluated[l];
            stats->full[l] = h.filled_inside[l];
            stats->empty[l] = h.filled_outside[l];
            stats->ambiguous[l] = h.ambiguous[l];
        }
        stats->leaf_empty = ls[0]; stats->leaf_full = ls[1]; stats->leaf_surface = ls[2];
        stats->float_points = ls[3]; stats->grad_points = ls[4];
        stats->arena_bytes_used = hc.arena_top * sizeof(uint2);
        stats->kernel_launches = launches;
        if (timing) cudaEventElapsedTime(&stats->total_ms, c->events[0], c->events[1]);
    }
    return rc;
}

}  // extern "C"

// ---- post-processing effects (fidget-raster/src/effects.rs) ---------------------------------------
namespace {

// Device view of an input image: the pointer itself, or a staged copy of a host image.
int32_t fx_input(fc_ctx* c, DevBuf& buf, const void* p, size_t bytes, const void** dev) {
    if (is_device_ptr(p)) { *dev = p; return FC_OK; }
    CU(buf.ensure(std::max<size_t>(bytes, 16)));
    CU(cudaMemcpyAsync(buf.p, p, bytes, cudaMemcpyHostToDevice, c->stream));
    *dev = buf.p;
    return FC_OK;
}
int32_t fx_output(fc_ctx* c, void* p, size_t bytes, void** dev) {
    if (is_device_ptr(p)) { *dev = p; return FC_OK; }
    CU(c->fx_out.ensure(std::max<size_t>(bytes, 16)));
    *dev = c->fx_out.p;
    return FC_OK;
}
int32_t fx_finish(fc_ctx* c, void* p, const void* dev, size_t bytes) {
    CU(cudaGetLastError());
    if (p != dev) CU(cudaMemcpyAsync(p, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return FC_OK;
}
int32_t fx_tables(fc_ctx* c, const float* kernel, uint32_t nk, const float* noise, uint32_t nn, const float** dk,
                  const float** dn) {
    if (!kernel || !noise || !nk || !nn) return fail(FC_ERR_INVALID, "SSAO needs a kernel and a noise table");
    if (nk > 4096) return fail(FC_ERR_INVALID, "SSAO kernel table too large (max 4096 samples)");
    CU(c->fx_tables.ensure((size_t(nk) * 3 + size_t(nn) * 2) * 4));
    float* t = c->fx_tables.as<float>();
    const cudaMemcpyKind any = cudaMemcpyDefault;
    CU(cudaMemcpyAsync(t, kernel, size_t(nk) * 12, any, c->stream));
    CU(cudaMemcpyAsync(t + size_t(nk) * 3, noise, size_t(nn) * 8, any, c->stream));
    *dk = t;
    *dn = t + size_t(nk) * 3;
    return FC_OK;
}
#define FX(call) do { if (int32_t rc_ = (call)) return rc_; } while (0)

int32_t fx_to_rgba(fc_ctx* c, int mode, const float* image, uint32_t w, uint32_t h, uint8_t* out) {
    if (!c || !image || !out) return fail(FC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 4, &din));
    FX(fx_output(c, out, n * 4, &dout));
    launch_to_rgba(mode, static_cast<const float*>(din), n, static_cast<uint8_t*>(dout), c->stream);
    return fx_finish(c, out, dout, n * 4);
}

}  // namespace

extern "C" {

int32_t fc_denoise_normals(fc_ctx* c, const fc_geometry_pixel* image, uint32_t w, uint32_t h, fc_geometry_pixel* out) {
    if (!c || !image || !out) return fail(FC_ERR_INVALID, "null argument");
    if (image == out) return fail(FC_ERR_INVALID, "fc_denoise_normals cannot run in place");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 16, &din));
    FX(fx_output(c, out, n * 16, &dout));
    launch_denoise_normals(static_cast<const GeoPixel*>(din), w, h, static_cast<GeoPixel*>(dout), c->stream);
    return fx_finish(c, out, dout, n * 16);
}

int32_t fc_compute_ssao(fc_ctx* c, const fc_geometry_pixel* image, uint32_t w, uint32_t h, uint32_t d,
                        const float* kernel, uint32_t n_kernel, const float* noise, uint32_t n_noise, float* out) {
    if (!c || !image || !out) return fail(FC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const float *dk, *dn;
    FX(fx_tables(c, kernel, n_kernel, noise, n_noise, &dk, &dn));
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 16, &din));
    FX(fx_output(c, out, n * 4, &dout));
    launch_compute_ssao(static_cast<const GeoPixel*>(din), w, h, d, dk, n_kernel, dn, n_noise,
                        static_cast<float*>(dout), c->stream);
    return fx_finish(c, out, dout, n * 4);
}

int32_t fc_blur_ssao(fc_ctx* c, const float* ssao, uint32_t w, uint32_t h, float* out) {
    if (!c || !ssao || !out) return fail(FC_ERR_INVALID, "null argument");
    if (ssao == out) return fail(FC_ERR_INVALID, "fc_blur_ssao cannot run in place");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, ssao, n * 4, &din));
    FX(fx_output(c, out, n * 4, &dout));
    launch_blur_ssao(static_cast<const float*>(din), w, h, static_cast<float*>(dout), c->stream);
    return fx_finish(c, out, dout, n * 4);
}

int32_t fc_apply_shading(fc_ctx* c, const fc_geometry_pixel* image, uint32_t w, uint32_t h, uint32_t d, int32_t ssao,
                         const float* kernel, uint32_t n_kernel, const float* noise, uint32_t n_noise,
                         uint8_t* out_rgb) {
    if (!c || !image || !out_rgb) return fail(FC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 16, &din));
    FX(fx_output(c, out_rgb, n * 3, &dout));
    const float* occl = nullptr;
    if (ssao) {
        const float *dk, *dn;
        FX(fx_tables(c, kernel, n_kernel, noise, n_noise, &dk, &dn));
        CU(c->fx_tmp.ensure(n * 4));
        launch_compute_ssao(static_cast<const GeoPixel*>(din), w, h, d, dk, n_kernel, dn, n_noise,
                            c->fx_tmp.as<float>(), c->stream);
        occl = c->fx_tmp.as<float>();
    }
    // the 3x3-window blur of the occlusion map is applied inside the shading kernel
    launch_apply_shading(static_cast<const GeoPixel*>(din), w, h, d, occl, 1, static_cast<uint8_t*>(dout), c->stream);
    return fx_finish(c, out_rgb, dout, n * 3);
}

int32_t fc_shade_with_occlusion(fc_ctx* c, const fc_geometry_pixel* image, uint32_t w, uint32_t h, uint32_t d,
                                const float* blurred_ssao, uint8_t* out_rgb) {
    if (!c || !image || !out_rgb) return fail(FC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void *din, *docc = nullptr; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 16, &din));
    if (blurred_ssao) FX(fx_input(c, c->fx_tmp, blurred_ssao, n * 4, &docc));
    FX(fx_output(c, out_rgb, n * 3, &dout));
    launch_apply_shading(static_cast<const GeoPixel*>(din), w, h, d, static_cast<const float*>(docc), 0,
                         static_cast<uint8_t*>(dout), c->stream);
    return fx_finish(c, out_rgb, dout, n * 3);
}

int32_t fc_normals_to_color(fc_ctx* c, const fc_geometry_pixel* image, uint32_t w, uint32_t h, uint8_t* out_rgb) {
    if (!c || !image || !out_rgb) return fail(FC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 16, &din));
    FX(fx_output(c, out_rgb, n * 3, &dout));
    launch_normals_to_color(static_cast<const GeoPixel*>(din), n, static_cast<uint8_t*>(dout), c->stream);
    return fx_finish(c, out_rgb, dout, n * 3);
}

int32_t fc_to_rgba_bitmap(fc_ctx* c, const float* image, uint32_t w, uint32_t h, int32_t transparent, uint8_t* out) {
    return fx_to_rgba(c, transparent ? 1 : 0, image, w, h, out);
}
int32_t fc_to_debug_bitmap(fc_ctx* c, const float* image, uint32_t w, uint32_t h, uint8_t* out) {
    return fx_to_rgba(c, 2, image, w, h, out);
}
int32_t fc_to_rgba_distance(fc_ctx* c, const float* image, uint32_t w, uint32_t h, uint8_t* out) {
    return fx_to_rgba(c, 3, image, w, h, out);
}

}  // extern "C"
