// Meshing back half on the device: the Hermite data fc_octree_sample leaves in HBM is
// turned into a triangle mesh without going back to the host.
//
//   vertices   one per connected group of inside corners of every surface leaf (the rule behind
//              CELL_TO_VERT_TO_EDGES, fidget-mesh/build.rs:25-130), positioned by the quadratic error
//              function of the group's edge intersections: QuadraticErrorSolver::add_intersection / solve
//              (fidget-mesh/src/qef.rs:44-168) -- mass point, A^T A, truncated pseudo-inverse with the
//              relative eigenvalue cut-off 1e-3; a NaN gradient snaps to the intersection (octree.rs:793-801);
//   triangles  dc_edge (fidget-mesh/src/dc.rs:104-213) for leaves of equal depth: every sign-changing cell edge is
//              shared by four leaves [a, b, c, d] = [0, U, U|V, V] around +T; the four cell vertices form a fan
//              around the edge's intersection vertex (taken from cell d), winding 3 or 1 by the sign at the
//              edge's start;
//   STL        Mesh::write_stl (fidget-mesh/src/output.rs:7-38).
// Out of scope here: cell collapse (octree.rs:252-440).  The reference merges the eight children of a branch into
// one leaf when the merged QEF error is small and the result stays manifold; without it the mesh is the
// uniform-depth Manifold Dual Contouring mesh -- same surface, more triangles in flat regions.
//
// The 3x3 symmetric eigen-problem is solved with cyclic Jacobi rotations in f32 (the reference calls nalgebra's
// SVD, a third-party algorithm not under /root/reference); positions agree to ~1e-5 of a cell, not bit for bit.
#include "capi_internal.h"

namespace fdev {

__host__ __device__ inline uint32_t next_axis(uint32_t a) { return a == 1u ? 2u : (a == 2u ? 4u : 1u); }   // X -> Y -> Z -> X

struct MeshScratch {
    const OctreeLeaf* leaves;
    uint32_t n_leaves;
    unsigned long long* hkeys;   // open-addressing table: packed cell coordinates -> leaf index
    uint32_t* hvals;
    uint32_t hmask;
    float3* cell_verts;          // [n_leaves][4]
    uint32_t* corner_vert;       // [n_leaves]: 2 bits per corner = vertex (group) of an inside corner
    uint32_t* remap;             // [n_leaves][16]: slot (4 cell vertices, 12 edge vertices) -> output vertex, or ~0
    uint32_t* counts;            // [0] vertices, [1] triangles, [2] triangle cursor, [3] edges without four leaves
    float3* out_verts;
    uint32_t cap_verts;
    uint3* out_tris;
    uint32_t cap_tris;
};

__device__ __forceinline__ unsigned long long cell_key(uint32_t x, uint32_t y, uint32_t z) {
    return (unsigned long long)x | ((unsigned long long)y << 16) | ((unsigned long long)z << 32);
}
__device__ __forceinline__ uint32_t hash_key(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return uint32_t(k);
}
__global__ void k_mesh_hash(MeshScratch m) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.n_leaves) return;
    const OctreeLeaf& L = m.leaves[i];
    const unsigned long long key = cell_key(L.ix, L.iy, L.iz);
    uint32_t h = hash_key(key) & m.hmask;
    for (;;) {
        const unsigned long long old = atomicCAS(&m.hkeys[h], ~0ull, key);
        if (old == ~0ull || old == key) { m.hvals[h] = i; return; }
        h = (h + 1u) & m.hmask;
    }
}
__device__ __forceinline__ uint32_t find_leaf(const MeshScratch& m, uint32_t x, uint32_t y, uint32_t z) {
    const unsigned long long key = cell_key(x, y, z);
    uint32_t h = hash_key(key) & m.hmask;
    for (;;) {
        const unsigned long long k = m.hkeys[h];
        if (k == key) return m.hvals[h];
        if (k == ~0ull) return ~0u;
        h = (h + 1u) & m.hmask;
    }
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations: a = V diag(w) V^T
__device__ inline void jacobi3(float a[3][3], float w[3], float v[3][3]) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0f : 0.0f;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const float off = fabsf(a[0][1]) + fabsf(a[0][2]) + fabsf(a[1][2]);
        if (off < 1e-30f) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabsf(a[p][q]) < 1e-37f) continue;
                const float theta = (a[q][q] - a[p][p]) / (2.0f * a[p][q]);
                const float t = (theta >= 0.0f ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
                const float c = 1.0f / sqrtf(t * t + 1.0f), s = t * c;
                for (int k = 0; k < 3; ++k) {   // A <- A J
                    const float akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {   // A <- J^T A
                    const float apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const float vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) w[i] = a[i][i];
}

// One thread per leaf: groups of inside corners, one QEF vertex per group
__global__ void __launch_bounds__(128) k_mesh_vertices(MeshScratch m) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.n_leaves) return;
    const OctreeLeaf& L = m.leaves[i];
    const uint32_t mask = L.mask;
    // connected groups of inside corners along cube edges (label = lowest corner of the group)
    uint32_t label[8];
    for (uint32_t c = 0; c < 8; ++c) label[c] = c;
    for (int it = 0; it < 8; ++it) {
        bool changed = false;
        for (uint32_t c = 0; c < 8; ++c) {
            if (!((mask >> c) & 1u)) continue;
            for (uint32_t ax = 1; ax < 8; ax <<= 1) {
                const uint32_t g = c ^ ax;
                if (!((mask >> g) & 1u)) continue;
                const uint32_t lo = min(label[c], label[g]);
                changed |= (label[c] != lo) | (label[g] != lo);
                label[c] = lo;
                label[g] = lo;
            }
        }
        if (!changed) break;
    }
    uint32_t n_groups = 0, group_of[8], packed = 0;
    for (uint32_t c = 0; c < 8; ++c) {
        group_of[c] = 0;
        if (!((mask >> c) & 1u)) continue;
        if (label[c] == c) group_of[c] = n_groups++;
        else group_of[c] = group_of[label[c]];   // label[c] < c: already numbered
        packed |= (group_of[c] & 3u) << (2u * c);
    }
    m.corner_vert[i] = packed | (n_groups << 16);
    for (uint32_t g = 0; g < n_groups && g < 4u; ++g) {
        // QuadraticErrorSolver (qef.rs:44-61)
        float ata[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, atb[3] = {0, 0, 0}, mp[4] = {0, 0, 0, 0};
        bool forced = false;
        float3 force_pos = make_float3(0, 0, 0);
        for (uint32_t s = 0; s < 8 && !forced; ++s) {
            if (!((mask >> s) & 1u) || group_of[s] != g) continue;
            for (uint32_t t = 1; t < 8 && !forced; t <<= 1) {
                const uint32_t e_end = s ^ t;
                if ((mask >> e_end) & 1u) continue;   // not a transition
                const uint32_t u = next_axis(t), v = next_axis(u);
                const uint32_t ti = t == 1u ? 0u : (t == 2u ? 1u : 2u);
                const uint32_t e = ti * 4u + ((s & u) ? 1u : 0u) + ((s & v) ? 2u : 0u);
                const float px = L.pos[e][0], py = L.pos[e][1], pz = L.pos[e][2];
                const float gx = L.grad[e][0], gy = L.grad[e][1], gz = L.grad[e][2], gw = L.grad[e][3];
                if (gx != gx || gy != gy || gz != gz || gw != gw) {   // octree.rs:793-801
                    forced = true;
                    force_pos = make_float3(px, py, pz);
                    break;
                }
                mp[0] += px; mp[1] += py; mp[2] += pz; mp[3] += 1.0f;
                const float nl = sqrtf(gx * gx + gy * gy + gz * gz);
                const float n[3] = {gx / nl, gy / nl, gz / nl};
                const float d = n[0] * px + n[1] * py + n[2] * pz;
                for (int r = 0; r < 3; ++r) {
                    for (int c2 = 0; c2 < 3; ++c2) ata[r][c2] += n[r] * n[c2];
                    atb[r] += n[r] * d;
                }
            }
        }
        float3 pos;
        if (forced) {
            pos = force_pos;
        } else {
            // QuadraticErrorSolver::solve (qef.rs:67-168)
            const float center[3] = {mp[0] / mp[3], mp[1] / mp[3], mp[2] / mp[3]};
            float b[3];
            for (int r = 0; r < 3; ++r) b[r] = atb[r] - (ata[r][0] * center[0] + ata[r][1] * center[1] + ata[r][2] * center[2]);
            float w[3], V[3][3], a2[3][3];
            for (int r = 0; r < 3; ++r) for (int c2 = 0; c2 < 3; ++c2) a2[r][c2] = ata[r][c2];
            jacobi3(a2, w, V);
            // singular values of a symmetric matrix = |eigenvalues|, sorted descending
            int order[3] = {0, 1, 2};
            for (int x = 0; x < 2; ++x) for (int y = x + 1; y < 3; ++y)
                if (fabsf(w[order[y]]) > fabsf(w[order[x]])) { const int tmp = order[x]; order[x] = order[y]; order[y] = tmp; }
            const float cutoff = fabsf(w[order[0]]) * 1e-3f;
            int rank = 3;
            for (int k = 0; k < 3; ++k) if (fabsf(w[order[k]]) < cutoff) { rank = k; break; }
            const float eps = rank < 3 ? fabsf(w[order[rank]]) : 0.0f;
            float sol[3] = {0, 0, 0};
            for (int k = 0; k < 3; ++k) {
                const int j = order[k];
                if (!(fabsf(w[j]) > eps)) continue;   // svd.solve: singular values <= eps are dropped
                const float coef = (V[0][j] * b[0] + V[1][j] * b[1] + V[2][j] * b[2]) / w[j];
                sol[0] += coef * V[0][j]; sol[1] += coef * V[1][j]; sol[2] += coef * V[2][j];
            }
            pos = make_float3(sol[0] + center[0], sol[1] + center[1], sol[2] + center[2]);
            if (!(pos.x == pos.x && pos.y == pos.y && pos.z == pos.z)) pos = make_float3(center[0], center[1], center[2]);
        }
        m.cell_verts[size_t(i) * 4 + g] = pos;
    }
}

// The four leaves around the +T edge at corner 0 of leaf `c` (dc.rs:104-119), or false at the domain boundary
struct EdgeCells { uint32_t leaf[4]; };
__device__ __forceinline__ bool edge_cells(const MeshScratch& m, uint32_t ci, uint32_t t, EdgeCells& ec) {
    const OctreeLeaf& C = m.leaves[ci];
    const uint32_t u = next_axis(t), v = next_axis(u);
    const uint32_t x = C.ix, y = C.iy, z = C.iz;
    const uint32_t du[3] = {(u & 1u) ? 1u : 0u, (u & 2u) ? 1u : 0u, (u & 4u) ? 1u : 0u};
    const uint32_t dv[3] = {(v & 1u) ? 1u : 0u, (v & 2u) ? 1u : 0u, (v & 4u) ? 1u : 0u};
    if ((du[0] + dv[0]) > x || (du[1] + dv[1]) > y || (du[2] + dv[2]) > z) return false;
    ec.leaf[2] = ci;                                                                   // c = a + U + V
    ec.leaf[0] = find_leaf(m, x - du[0] - dv[0], y - du[1] - dv[1], z - du[2] - dv[2]);  // a
    ec.leaf[1] = find_leaf(m, x - dv[0], y - dv[1], z - dv[2]);                          // b = a + U
    ec.leaf[3] = find_leaf(m, x - du[0], y - du[1], z - du[2]);                          // d = a + V
    return ec.leaf[0] != ~0u && ec.leaf[1] != ~0u && ec.leaf[3] != ~0u;
}

// pass 0: mark the vertex slots in use and count triangles; pass 1: emit
template <int PASS>
__global__ void __launch_bounds__(128) k_mesh_faces(MeshScratch m) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m.n_leaves * 3u) return;
    const uint32_t ci = gid / 3u, ti = gid % 3u, t = 1u << ti;
    const OctreeLeaf& C = m.leaves[ci];
    const uint32_t in0 = C.mask & 1u, in1 = (C.mask >> t) & 1u;
    if (in0 == in1) return;
    EdgeCells ec;
    if (!edge_cells(m, ci, t, ec)) {
        if (PASS == 0) atomicAdd(&m.counts[3], 1u);
        return;
    }
    const uint32_t u = next_axis(t), v = next_axis(u);
    const uint32_t edge_of[4] = {ti * 4u + 3u, ti * 4u + 2u, ti * 4u + 0u, ti * 4u + 1u};   // a, b, c, d
    uint32_t slot[4];
    for (int k = 0; k < 4; ++k) {
        const uint32_t e = edge_of[k];
        const uint32_t start = ((e & 1u) ? u : 0u) | ((e & 2u) ? v : 0u), end = start | t;
        const uint32_t mk = m.leaves[ec.leaf[k]].mask;
        const uint32_t inside_corner = ((mk >> start) & 1u) ? start : end;
        if ((((mk >> start) & 1u) == ((mk >> end) & 1u))) {   // the neighbour does not see the sign change: skip
            if (PASS == 0) atomicAdd(&m.counts[3], 1u);
            return;
        }
        slot[k] = ec.leaf[k] * 16u + ((m.corner_vert[ec.leaf[k]] >> (2u * inside_corner)) & 3u);
    }
    const uint32_t islot = ec.leaf[3] * 16u + 4u + edge_of[3];   // intersection vertex: cell d's copy
    if (PASS == 0) {
        for (int k = 0; k < 4; ++k) m.remap[slot[k]] = 1u;
        m.remap[islot] = 1u;
        atomicAdd(&m.counts[1], 4u);
        return;
    }
    // winding (dc.rs:188-196): 3 when the edge's start corner is outside, else 1
    const uint32_t md = m.leaves[ec.leaf[3]].mask;
    const uint32_t start_d = ((edge_of[3] & 1u) ? u : 0u) | ((edge_of[3] & 2u) ? v : 0u);
    const uint32_t winding = ((md >> start_d) & 1u) ? 1u : 3u;
    const uint32_t base = atomicAdd(&m.counts[2], 4u);
    const uint32_t iv = m.remap[islot];
    for (uint32_t j = 0; j < 4u; ++j)
        if (base + j < m.cap_tris) m.out_tris[base + j] = make_uint3(m.remap[slot[j]], m.remap[slot[(j + winding) & 3u]], iv);
}

// compaction of the used vertex slots
__global__ void k_mesh_assign(MeshScratch m) {
    const uint64_t s = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= uint64_t(m.n_leaves) * 16u) return;
    if (m.remap[s] != 1u) { m.remap[s] = ~0u; return; }
    const uint32_t id = atomicAdd(&m.counts[0], 1u);
    m.remap[s] = id;
    if (id >= m.cap_verts) return;
    const uint32_t leaf = uint32_t(s / 16u), k = uint32_t(s % 16u);
    if (k < 4u) m.out_verts[id] = m.cell_verts[size_t(leaf) * 4 + k];
    else {
        const OctreeLeaf& L = m.leaves[leaf];
        m.out_verts[id] = make_float3(L.pos[k - 4u][0], L.pos[k - 4u][1], L.pos[k - 4u][2]);
    }
}

// Mesh::write_stl (output.rs:7-38): 80-byte header, u32 count, 50 bytes per triangle
__global__ void k_mesh_stl(const float3* verts, const uint3* tris, uint32_t n_tris, uint8_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        const char hdr[] = "This is a binary STL file exported by Fidget";
        for (int k = 0; k < 80; ++k) out[k] = k < int(sizeof(hdr) - 1) ? uint8_t(hdr[k]) : 0;
        for (int k = 0; k < 4; ++k) out[80 + k] = uint8_t(n_tris >> (8 * k));
    }
    if (i >= n_tris) return;
    const uint3 t = tris[i];
    const float3 a = verts[t.x], b = verts[t.y], c = verts[t.z];
    const float3 ab = make_float3(b.x - a.x, b.y - a.y, b.z - a.z), ac = make_float3(c.x - a.x, c.y - a.y, c.z - a.z);
    const float rec[12] = {ab.y * ac.z - ab.z * ac.y, ab.z * ac.x - ab.x * ac.z, ab.x * ac.y - ab.y * ac.x,
                           a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z};
    uint16_t* dst = reinterpret_cast<uint16_t*>(out + 84 + size_t(i) * 50);   // 84 + 50 i is even
    for (int k = 0; k < 12; ++k) {
        const uint32_t bits = __float_as_uint(rec[k]);
        dst[2 * k] = uint16_t(bits & 0xffffu);
        dst[2 * k + 1] = uint16_t(bits >> 16);
    }
    dst[24] = 0;
}

}  // namespace fdev

// fc_octree_sample's device half (octree_capi.cu)
int32_t octree_sample_device(fc_ctx* c, const fc_tape* tape, const fc_octree_cfg* cfg, OctreeLeaf* dout, uint64_t cap,
                             uint32_t* n_out, fc_octree_stats* stats);

extern "C" {

int32_t fc_mesh_build(fc_ctx* c, const fc_tape* tape, const fc_octree_cfg* cfg, fc_mesh_info* info) {
    if (!c || !tape || !cfg || !info) return fail(FC_ERR_INVALID, "null argument");
    memset(info, 0, sizeof *info);
    // ---- sampler: leaves stay in HBM ----
    uint64_t cap = c->mesh_leaves.cap / sizeof(OctreeLeaf);
    if (cap < 1024) cap = std::max<uint64_t>(1024, std::min<uint64_t>(1ull << (3 * cfg->depth), 6ull << (2 * cfg->depth)));
    uint32_t n = 0;
    fc_octree_stats ost;
    for (int attempt = 0; attempt < 2; ++attempt) {
        CU(cudaSetDevice(c->device));
        CU(c->mesh_leaves.ensure(cap * sizeof(OctreeLeaf)));
        int32_t rc = octree_sample_device(c, tape, cfg, c->mesh_leaves.as<OctreeLeaf>(), cap, &n, &ost);
        if (rc == FC_OK) break;
        if (n > cap && attempt == 0) { cap = n; continue; }   // retry once with the exact count
        return rc;
    }
    std::lock_guard<std::mutex> guard(c->mu);
    cudaStream_t s = c->stream;
    MeshScratch m{};
    m.leaves = c->mesh_leaves.as<OctreeLeaf>();
    m.n_leaves = n;
    info->n_leaves = n;
    info->sampler_ms = ost.total_ms;
    c->mesh_n_verts = c->mesh_n_tris = 0;
    if (n == 0) return FC_OK;
    uint32_t hsize = 1024;
    while (hsize < 2u * n) hsize <<= 1;
    const size_t b_keys = size_t(hsize) * 8, b_vals = size_t(hsize) * 4, b_cv = size_t(n) * 4 * sizeof(float3), b_cn = size_t(n) * 4,
                 b_remap = size_t(n) * 16 * 4;
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    CU(c->mesh_scratch.ensure(al(b_keys) + al(b_vals) + al(b_cv) + al(b_cn) + al(b_remap) + 256));
    char* q = c->mesh_scratch.as<char>();
    m.hkeys = reinterpret_cast<unsigned long long*>(q); q += al(b_keys);
    m.hvals = reinterpret_cast<uint32_t*>(q); q += al(b_vals);
    m.cell_verts = reinterpret_cast<float3*>(q); q += al(b_cv);
    m.corner_vert = reinterpret_cast<uint32_t*>(q); q += al(b_cn);
    m.remap = reinterpret_cast<uint32_t*>(q); q += al(b_remap);
    m.counts = reinterpret_cast<uint32_t*>(q);
    m.hmask = hsize - 1;
    cudaEvent_t e0 = get_event(c, 0), e1 = get_event(c, 1);
    CU(cudaEventRecord(e0, s));
    CU(cudaMemsetAsync(m.hkeys, 0xff, b_keys, s));
    CU(cudaMemsetAsync(m.remap, 0, b_remap, s));
    CU(cudaMemsetAsync(m.counts, 0, 64, s));
    const unsigned bl = (n + 127) / 128;
    k_mesh_hash<<<bl, 128, 0, s>>>(m);
    k_mesh_vertices<<<bl, 128, 0, s>>>(m);
    k_mesh_faces<0><<<(n * 3u + 127) / 128, 128, 0, s>>>(m);
    uint32_t cnt[4];
    CU(cudaMemcpyAsync(cnt, m.counts, sizeof cnt, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const uint32_t n_tris = cnt[1];
    // every used slot becomes a vertex: count them on the device, sized by the worst case (5 slots per triangle fan)
    const uint64_t v_cap = std::min<uint64_t>(uint64_t(n) * 16, uint64_t(n_tris) * 5 / 4 + 16);
    CU(c->mesh_verts.ensure(std::max<uint64_t>(v_cap, 1) * sizeof(float3)));
    CU(c->mesh_tris.ensure(std::max<uint64_t>(n_tris, 1) * sizeof(uint3)));
    m.out_verts = c->mesh_verts.as<float3>();
    m.cap_verts = uint32_t(v_cap);
    m.out_tris = c->mesh_tris.as<uint3>();
    m.cap_tris = n_tris;
    k_mesh_assign<<<unsigned((uint64_t(n) * 16 + 255) / 256), 256, 0, s>>>(m);
    k_mesh_faces<1><<<(n * 3u + 127) / 128, 128, 0, s>>>(m);
    CU(cudaEventRecord(e1, s));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(cnt, m.counts, sizeof cnt, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (cnt[0] > v_cap) return fail(FC_ERR_CUDA, "mesh vertex buffer overflow");
    c->mesh_n_verts = cnt[0];
    c->mesh_n_tris = n_tris;
    info->n_vertices = cnt[0];
    info->n_triangles = n_tris;
    info->open_edges = cnt[3];
    cudaEventElapsedTime(&info->mesh_ms, e0, e1);
    return FC_OK;
}

int32_t fc_mesh_read(fc_ctx* c, float* vertices, uint32_t* triangles) {
    if (!c) return fail(FC_ERR_INVALID, "null ctx");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    if (vertices && c->mesh_n_verts)
        CU(cudaMemcpyAsync(vertices, c->mesh_verts.p, size_t(c->mesh_n_verts) * 12, cudaMemcpyDefault, c->stream));
    if (triangles && c->mesh_n_tris)
        CU(cudaMemcpyAsync(triangles, c->mesh_tris.p, size_t(c->mesh_n_tris) * 12, cudaMemcpyDefault, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return FC_OK;
}

int32_t fc_mesh_write_stl(fc_ctx* c, uint8_t* buf, size_t cap, size_t* n_bytes) {
    if (!c) return fail(FC_ERR_INVALID, "null ctx");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t need = 84 + size_t(c->mesh_n_tris) * 50;
    if (n_bytes) *n_bytes = need;
    if (!buf) return FC_OK;
    if (cap < need) return fail(FC_ERR_INVALID, "buffer too small");
    const bool dev = is_device_ptr(buf);
    uint8_t* d = buf;
    if (!dev) {
        CU(c->fx_out.ensure(need));
        d = c->fx_out.as<uint8_t>();
    }
    k_mesh_stl<<<unsigned((std::max<uint32_t>(c->mesh_n_tris, 1) + 127) / 128), 128, 0, c->stream>>>(
        c->mesh_verts.as<float3>(), c->mesh_tris.as<uint3>(), c->mesh_n_tris, d);
    CU(cudaGetLastError());
    if (!dev) CU(cudaMemcpyAsync(buf, d, need, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return FC_OK;
}

}  // extern "C"
