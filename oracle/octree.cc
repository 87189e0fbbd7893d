// ORACLE -- test infrastructure, not product code.  See octree.h.
#include "octree.h"

#include <algorithm>
#include <atomic>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace oracle {
using namespace fhost;

namespace {

struct Bounds { Interval b[3]; };

Bounds child_bounds(const Bounds& p, int corner) {   // CellBounds::child (cell.rs:155-166)
    Bounds c;
    for (int i = 0; i < 3; ++i) {
        float mid = (p.b[i].lo + p.b[i].hi) / 2.0f;
        c.b[i] = (corner >> i) & 1 ? Interval(mid, p.b[i].hi) : Interval(p.b[i].lo, mid);
    }
    return c;
}
float lerp(Interval iv, float frac) { return iv.lo * (1.0f - frac) + iv.hi * frac; }   // interval.rs:467-470
void cell_pos(const Bounds& b, const uint32_t p[3], float out[3]) {   // CellBounds::pos (cell.rs:183-192)
    for (int i = 0; i < 3; ++i) out[i] = lerp(b.b[i], float(uint16_t(p[i])) / 65535.0f);
}

struct Builder {
    const OctreeConfig& cfg;
    IntervalEval ieval;
    FloatSliceEval feval;
    GradSliceEval geval;
    std::vector<OctreeLeaf> leaves;
    OctreeStats stats;
    std::vector<float> sx, sy, sz, sout;
    std::vector<Grad> gx, gy, gz, gout;
    bool count_only = false;   // stop at cfg.depth without sampling leaves (census of the top levels)

    explicit Builder(const OctreeConfig& c) : cfg(c), sx(192), sy(192), sz(192), sout(192), gx(12), gy(12), gz(12), gout(12) {}

    struct Slots { int x = -1, y = -1, z = -1; };
    static Slots slots(const Tape& t) {
        Slots s;
        for (size_t i = 0; i < t.d.vars.order.size(); ++i) {
            auto k = t.d.vars.order[i].kind;
            if (k == Var::X) s.x = int(i); else if (k == Var::Y) s.y = int(i); else if (k == Var::Z) s.z = int(i);
            else throw std::runtime_error("octree oracle supports X/Y/Z only");
        }
        return s;
    }

    void eval_points(const Tape& t, size_t n) {
        for (size_t i = 0; i < n; ++i)
            if (cfg.has_transform) {
                float p[3];
                transform_f32(sx[i], sy[i], sz[i], cfg.world_to_model, p);
                sx[i] = p[0]; sy[i] = p[1]; sz[i] = p[2];
            }
        Slots s = slots(t);
        const float* vars[3] = {sx.data(), sx.data(), sx.data()};
        if (s.x >= 0) vars[s.x] = sx.data();
        if (s.y >= 0) vars[s.y] = sy.data();
        if (s.z >= 0) vars[s.z] = sz.data();
        float* outs[1] = {sout.data()};
        feval.eval(t, vars, n, outs);
        stats.float_points += n;
    }

    void recurse(RenderHandle* h, const Bounds& b, uint32_t depth, uint32_t ix, uint32_t iy, uint32_t iz) {
        const Tape& tape = *h->shape;
        Interval xyz[3] = {b.b[0], b.b[1], b.b[2]};
        if (cfg.has_transform) transform_interval(b.b[0], b.b[1], b.b[2], cfg.world_to_model, xyz);
        Slots s = slots(tape);
        Interval vars[3];
        if (s.x >= 0) vars[s.x] = xyz[0];
        if (s.y >= 0) vars[s.y] = xyz[1];
        if (s.z >= 0) vars[s.z] = xyz[2];
        Interval r;
        bool has_trace = ieval.eval(tape, vars, &r);
        stats.evaluated[depth]++;
        if (r.hi < 0.0f) { stats.full[depth]++; return; }
        if (r.lo > 0.0f) { stats.empty[depth]++; return; }
        stats.ambiguous[depth]++;
        RenderHandle* sub = h;
        if (has_trace) {   // VM: simplify_tree_during_meshing is always true (render/mod.rs:269-273)
            std::vector<uint8_t> trace = ieval.choices;
            sub = h->simplify(trace);
        }
        if (depth == cfg.depth) {
            if (!count_only) leaf(sub, b, ix, iy, iz);
        } else {
            for (int c = 0; c < 8; ++c)
                recurse(sub, child_bounds(b, c), depth + 1, 2 * ix + (c & 1), 2 * iy + ((c >> 1) & 1), 2 * iz + ((c >> 2) & 1));
        }
    }

    // One step of recurse() without statistics: the handle the children of this cell are evaluated with
    // (alive = false when the cell is full or empty, i.e. has no children)
    RenderHandle* descend(RenderHandle* h, const Bounds& b, bool& alive) {
        const Tape& tape = *h->shape;
        Interval xyz[3] = {b.b[0], b.b[1], b.b[2]};
        if (cfg.has_transform) transform_interval(b.b[0], b.b[1], b.b[2], cfg.world_to_model, xyz);
        Slots s = slots(tape);
        Interval vars[3];
        if (s.x >= 0) vars[s.x] = xyz[0];
        if (s.y >= 0) vars[s.y] = xyz[1];
        if (s.z >= 0) vars[s.z] = xyz[2];
        Interval r;
        bool has_trace = ieval.eval(tape, vars, &r);
        if (r.hi < 0.0f || r.lo > 0.0f) { alive = false; return h; }
        if (has_trace) {
            std::vector<uint8_t> trace = ieval.choices;
            return h->simplify(trace);
        }
        return h;
    }

    void leaf(RenderHandle* h, const Bounds& b, uint32_t ix, uint32_t iy, uint32_t iz) {
        const Tape& tape = *h->shape;
        for (int c = 0; c < 8; ++c) {   // CellBounds::corner (cell.rs:170-180)
            sx[c] = (c & 1) ? b.b[0].hi : b.b[0].lo;
            sy[c] = (c & 2) ? b.b[1].hi : b.b[1].lo;
            sz[c] = (c & 4) ? b.b[2].hi : b.b[2].lo;
        }
        eval_points(tape, 8);
        uint32_t mask = 0;
        for (int c = 0; c < 8; ++c) if (sout[c] < 0.0f) mask |= 1u << c;
        if (mask == 0) { stats.leaf_empty++; return; }
        if (mask == 255) { stats.leaf_full++; return; }
        stats.leaf_surface++;

        // active edges in undirected-edge order (types.rs:208-219): e = 4 t + 2 [start & v] + [start & u]
        struct E { uint32_t start[3], end[3]; int index; };
        E edges[12];
        int ne = 0;
        for (int t = 0; t < 3; ++t) {
            int u = (t + 1) % 3, v = (t + 2) % 3;
            for (int sv = 0; sv < 2; ++sv)
                for (int su = 0; su < 2; ++su) {
                    int c0 = (su << u) | (sv << v), c1 = c0 | (1 << t);
                    bool in0 = (mask >> c0) & 1, in1 = (mask >> c1) & 1;
                    if (in0 == in1) continue;
                    E e;
                    e.index = 4 * t + 2 * sv + su;
                    for (int a = 0; a < 3; ++a) e.start[a] = e.end[a] = 0;
                    e.start[u] = e.end[u] = su ? 65535u : 0u;
                    e.start[v] = e.end[v] = sv ? 65535u : 0u;
                    e.start[t] = in0 ? 0u : 65535u;   // start is always inside (octree.rs:650-676)
                    e.end[t] = in0 ? 65535u : 0u;
                    edges[ne++] = e;
                }
        }
        const int SIZE = 16, ROUNDS = 4;
        for (int round = 0; round < ROUNDS; ++round) {
            int i = 0;
            for (int k = 0; k < ne; ++k)
                for (int j = 0; j < SIZE; ++j) {
                    uint32_t p[3];
                    for (int a = 0; a < 3; ++a)
                        p[a] = (edges[k].start[a] * uint32_t(SIZE - j - 1) + edges[k].end[a] * uint32_t(j)) / uint32_t(SIZE - 1);
                    float f[3];
                    cell_pos(b, p, f);
                    sx[i] = f[0]; sy[i] = f[1]; sz[i] = f[2];
                    ++i;
                }
            eval_points(tape, size_t(i));
            for (int k = 0; k < ne; ++k) {
                const float* search = &sout[k * SIZE];
                int frac = SIZE - 1;   // the reference unwraps `find(v >= 0)`; the last sample is outside
                for (int j = 0; j < SIZE; ++j) if (search[j] >= 0.0f) { frac = j; break; }
                if (frac == 0) frac = 1;   // debug_assert!(frac > 0) in the reference
                uint32_t a[3], c[3];
                for (int q = 0; q < 3; ++q) {
                    a[q] = (edges[k].start[q] * uint32_t(SIZE - (frac - 1) - 1) + edges[k].end[q] * uint32_t(frac - 1)) / uint32_t(SIZE - 1);
                    c[q] = (edges[k].start[q] * uint32_t(SIZE - frac - 1) + edges[k].end[q] * uint32_t(frac)) / uint32_t(SIZE - 1);
                }
                for (int q = 0; q < 3; ++q) { edges[k].start[q] = uint16_t(a[q]); edges[k].end[q] = uint16_t(c[q]); }
            }
        }
        OctreeLeaf L{};
        L.ix = uint16_t(ix); L.iy = uint16_t(iy); L.iz = uint16_t(iz);
        L.mask = uint8_t(mask);
        L.n_edges = uint8_t(ne);
        for (int k = 0; k < ne; ++k) {
            uint32_t p[3];
            for (int q = 0; q < 3; ++q) p[q] = uint16_t((edges[k].start[q] + edges[k].end[q]) / 2);
            float f[3];
            cell_pos(b, p, f);
            gx[k] = Grad(f[0], 1, 0, 0); gy[k] = Grad(f[1], 0, 1, 0); gz[k] = Grad(f[2], 0, 0, 1);
            L.pos[edges[k].index][0] = f[0]; L.pos[edges[k].index][1] = f[1]; L.pos[edges[k].index][2] = f[2];
            L.present |= uint16_t(1u << edges[k].index);
        }
        if (cfg.has_transform)
            for (int k = 0; k < ne; ++k) {
                Grad t[3];
                transform_grad(gx[k], gy[k], gz[k], cfg.world_to_model, t);
                gx[k] = t[0]; gy[k] = t[1]; gz[k] = t[2];
            }
        Slots s = slots(tape);
        const Grad* gv[3] = {gx.data(), gx.data(), gx.data()};
        if (s.x >= 0) gv[s.x] = gx.data();
        if (s.y >= 0) gv[s.y] = gy.data();
        if (s.z >= 0) gv[s.z] = gz.data();
        Grad* go[1] = {gout.data()};
        geval.eval(tape, gv, size_t(ne), go);
        stats.grad_points += uint64_t(ne);
        for (int k = 0; k < ne; ++k) {
            float* g = L.grad[edges[k].index];
            g[0] = gout[k].dx; g[1] = gout[k].dy; g[2] = gout[k].dz; g[3] = gout[k].v;
        }
        leaves.push_back(L);
    }
};

}  // namespace

namespace {
void sort_leaves(std::vector<OctreeLeaf>& leaves) {
    std::sort(leaves.begin(), leaves.end(), [](const OctreeLeaf& a, const OctreeLeaf& c) {
        if (a.iz != c.iz) return a.iz < c.iz;
        if (a.iy != c.iy) return a.iy < c.iy;
        return a.ix < c.ix;
    });
}
void add_stats(OctreeStats& a, const OctreeStats& b) {
    for (int i = 0; i < 16; ++i) {
        a.evaluated[i] += b.evaluated[i]; a.full[i] += b.full[i]; a.empty[i] += b.empty[i]; a.ambiguous[i] += b.ambiguous[i];
    }
    a.leaf_empty += b.leaf_empty; a.leaf_full += b.leaf_full; a.leaf_surface += b.leaf_surface;
    a.float_points += b.float_points; a.grad_points += b.grad_points;
}
}  // namespace

// cfg.threads > 1: the counterpart of Octree::build_inner_mt (fidget-mesh/src/octree.rs:162-208), which hands
// subtrees to worker threads.  Here every cell at depth SPLIT is a task; a worker walks the path from the
// root to its cell first (same interval evaluations and simplifications as the serial descent, not
// counted), then recurses below it with its own evaluators.  Leaves and statistics are the serial ones.
void octree_sample(const TapeP& tape, const OctreeConfig& cfg, std::vector<OctreeLeaf>& leaves, OctreeStats* stats) {
    leaves.clear();
    Bounds root;
    for (int i = 0; i < 3; ++i) root.b[i] = Interval(-1.0f, 1.0f);   // CellBounds::new (cell.rs:146-150)
    const uint32_t SPLIT = 3;
    if (cfg.threads <= 1 || cfg.depth <= SPLIT) {
        Builder b(cfg);
        RenderHandle h(tape);
        b.recurse(&h, root, 0, 0, 0, 0);
        leaves = std::move(b.leaves);
        sort_leaves(leaves);
        if (stats) *stats = b.stats;
        return;
    }
    // serial census of depths < SPLIT (a shallow copy of the configuration that stops there)
    OctreeStats total;
    {
        OctreeConfig top = cfg;
        top.depth = SPLIT - 1;
        Builder b(top);
        b.count_only = true;
        RenderHandle h(tape);
        b.recurse(&h, root, 0, 0, 0, 0);
        total = b.stats;
    }
    const uint32_t n_tasks = 1u << (3 * SPLIT);
    std::atomic<uint32_t> next{0};
    std::mutex mu;
    std::vector<std::thread> pool;
    for (int t = 0; t < cfg.threads; ++t)
        pool.emplace_back([&]() {
            Builder b(cfg);
            for (;;) {
                const uint32_t task = next.fetch_add(1);
                if (task >= n_tasks) break;
                // path: 3 bits per level, most significant level first
                RenderHandle h(tape);
                RenderHandle* cur = &h;
                Bounds bb = root;
                uint32_t ix = 0, iy = 0, iz = 0;
                bool alive = true;
                for (uint32_t d = 0; d < SPLIT && alive; ++d) {
                    cur = b.descend(cur, bb, alive);
                    const int c = int((task >> (3 * (SPLIT - 1 - d))) & 7u);
                    bb = child_bounds(bb, c);
                    ix = 2 * ix + (c & 1); iy = 2 * iy + ((c >> 1) & 1); iz = 2 * iz + ((c >> 2) & 1);
                }
                if (alive) b.recurse(cur, bb, SPLIT, ix, iy, iz);
            }
            std::lock_guard<std::mutex> g(mu);
            leaves.insert(leaves.end(), b.leaves.begin(), b.leaves.end());
            add_stats(total, b.stats);
        });
    for (auto& th : pool) th.join();
    sort_leaves(leaves);
    if (stats) *stats = total;
}

}  // namespace oracle
