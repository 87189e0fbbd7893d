// Octree sampler leaves of fidget-mesh's Manifold Dual Contouring (sampling half).
#include "interp.cuh"

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
