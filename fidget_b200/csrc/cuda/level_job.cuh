// One unit of work of the interval levels: a parent tile (or, at level 0, a group of 32 root tiles)
// whose children are evaluated by the lanes of one warp, classified, simplified and queued.  Shared
// by k_interval_level (one launch per level) and by the fused 2D kernel of tail2d.cu, where jobs of
// every level are claimed from one dependency-ordered queue inside a single persistent launch
// (FUSED): there a job becomes visible through a ready mark written after its fields, child tapes
// occupy whole 128-byte lines of the arena, and `outstanding` counts the jobs not yet finished.
#pragma once
#include "interp.cuh"

namespace fdev {

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// Waits until job slot `j` carries this render's ready mark, then reads it from L2 (another SM wrote it
// during this launch; L1 may hold a stale copy of the line).  Bounded: a lost job raises error bit 2.
__device__ __forceinline__ TileJob load_job_ready(const TileJob* j, uint32_t epoch, Counters* ctr) {
    uint32_t spins = 0;
    while (ld_volatile_u32(&j->pad) != epoch) {
        __nanosleep(64);
        if (++spins > (1u << 22)) { atomicOr(&ctr->error, 4u); break; }
    }
    __threadfence();
    TileJob o;
    const uint2* q = reinterpret_cast<const uint2*>(j);
    const uint2 a = __ldcg(q), b = __ldcg(q + 1), c = __ldcg(q + 2), d = __ldcg(q + 3), e = __ldcg(q + 4);
    o.x = a.x; o.y = a.y; o.z = b.x; o.pad = b.y;
    o.tape.ptr = reinterpret_cast<const uint2*>((unsigned long long)c.x | ((unsigned long long)c.y << 32));
    o.tape.n_ops = d.x; o.tape.ref_len = d.y; o.tape.n_choices = e.x; o.tape.pad = e.y;
    return o;
}
__device__ __forceinline__ void publish_job(TileJob* dst, TileJob o, uint32_t epoch) {
    o.pad = 0;
    *dst = o;
    __threadfence();
    *reinterpret_cast<volatile uint32_t*>(&dst->pad) = epoch;
}
__device__ __forceinline__ void store_fill(FillRec* dst, const FillRec& fr) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(fr.x, fr.y, fr.value, fr.ready);
}

template <int DIM, bool FUSED>
__device__ __forceinline__ void level_job(const LevelParams& p, uint32_t j, uint32_t n_roots, itv* slots, uint32_t* cs,
                                          uint32_t (*live)[32], int lane, uint32_t epoch) {
    const uint32_t T = p.tile;
    bool cull_open = false, cull_check = false;
    TapeRef tr;
    uint32_t px = 0, py = 0, pz = 0, nchild;
    if (p.root_mode) {
        tr = p.root_tape;
        nchild = min(32u, n_roots - j * 32u);
    } else {
        const TileJob jb = FUSED ? load_job_ready(p.jobs_in + j, epoch, p.ctr) : p.jobs_in[j];
        px = jb.x;
        py = jb.y;
        pz = jb.z;
        tr = jb.tape;
        nchild = p.n_axis * p.n_axis * (DIM == 3 ? p.n_axis : 1u);
        if (DIM == 3 && p.mode != 1u && p.cull) {
            // Every pixel under this parent already holds depth >= its top + 1 (tiles in front, proven inside by coarser
            // levels): nothing inside can show, so none of its children is evaluated.  The reference skips the same
            // tiles in its front-to-back walk (voxel.rs:283-293) -- and more, since it also knows the voxel hits in
            // front, which arrive last here.  One small read of the occlusion map per lane, not the heightmap itself.
            const uint32_t nb = (T * p.n_axis) / 16u, need = pz + T * p.n_axis + 1u;   // blocks per side of the parent
            for (uint32_t q = lane; q < nb * nb; q += 32u) {
                const uint32_t bx = px / 16u + q % nb, by = py / 16u + q / nb;
                if (bx < p.occl_w && by < p.occl_h) cull_open |= __ldcg(p.occl + size_t(by) * p.occl_w + bx) < need;
            }
            cull_check = true;
        }
    }
    const uint2* tape = tr.ptr;

    for (uint32_t chunk = 0; chunk * 32u < nchild; ++chunk) {
        const uint32_t c = chunk * 32u + lane;
        const bool valid = c < nchild;
        uint32_t cx, cy, cz = 0;
        if (p.root_mode) {
            root_corner(p, j * 32u + (valid ? c : 0u), T, cx, cy, cz);
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

        if (DIM == 3 && cull_check && chunk == 0) {
            if (!__any_sync(FULL, cull_open)) {
                if (p.stats && lane == 0) atomicAdd(&p.stats->culled[p.level], (unsigned long long)nchild);
                return;
            }
        }
        ChoicePacker pk;
        pk.base = cs;
        itv r = iv_nan();
        run_interval<!FUSED>(
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
                if (p.occl && T % 16u == 0u)   // the whole blocks this tile covers now hold its depth (the same value the heightmap gets)
                    for (uint32_t q = lane; q < (T / 16u) * (T / 16u); q += 32u) {
                        const uint32_t bx = fx / 16u + q % (T / 16u), by = fy / 16u + q / (T / 16u);
                        if (bx < p.occl_w && by < p.occl_h) atomicMax(p.occl + size_t(by) * p.occl_w + bx, fz + T + 1u);
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
                        fr.ready = epoch;
                        store_fill(p.fills + slot, fr);   // one 16-byte store: the ready mark travels with the record
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
        bool kept = false;
        const bool need = amb && pk.any_nonboth;
        const uint32_t mneed = __ballot_sync(FULL, need);
        if (mneed) {
            const uint32_t total = __popc(mneed);
            // worst-case slot per child; in the fused kernel slots are whole 128-byte lines, so that a line
            // written for one tape is never one an SM may already hold in L1 for another
            const uint32_t slot_ops = FUSED ? ((tr.n_ops + 15u) & ~15u) : tr.n_ops;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&p.ctr->arena_top, (unsigned long long)total * slot_ops + (FUSED ? 15u : 0u));
            base = __shfl_sync(FULL, base, 0);
            if (FUSED) base = (base + 15ull) & ~15ull;   // (the root level's tapes end anywhere)
            if (base + (unsigned long long)total * slot_ops > p.arena_cap) {
                if (lane == 0) atomicOr(&p.ctr->error, 1u);
            } else {
                const uint32_t rank = __popc(mneed & lanemask_lt());
                unsigned long long end = base + (unsigned long long)(rank + 1u) * slot_ops;
                ChoiceUnpacker cu;
                cu.base = cs;
                cu.ci = tr.n_choices;
                uint32_t n_dev, ref_len, nch;
                simplify_lane<!FUSED>(tape, tr.n_ops, need, live, lane, cu, p.arena + end, n_dev, ref_len, nch);
                bool keep = need && ref_len < tr.ref_len;
                kept = keep;
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

        if (DIM == 3 && p.census) {   // exact census: what this launch evaluated, judged later against the final heightmap
            const uint32_t mv = __ballot_sync(FULL, valid);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&p.ctr->n_census, uint32_t(__popc(mv)));
            base = __shfl_sync(FULL, base, 0);
            if (valid) {
                const uint32_t slot = base + __popc(mv & lanemask_lt());
                if (slot < p.cap_census) {
                    CensusRec r;
                    r.x = uint16_t(cx); r.y = uint16_t(cy); r.z = uint16_t(cz);
                    r.level = uint8_t(p.level);
                    r.flags = uint8_t((fill_in ? 1u : (fill_out ? 0u : 2u)) | (kept ? 4u : 0u));
                    p.census[slot] = r;
                } else atomicOr(&p.ctr->error, 2u);
            }
        }

        // queue ambiguous children for the next level
        const uint32_t mamb = __ballot_sync(FULL, amb);
        if (mamb) {
            uint32_t base = 0;
            if (lane == 0) {
                atomicAdd(&p.ctr->outstanding, uint32_t(__popc(mamb)));   // before the jobs become claimable
                if (FUSED) __threadfence();
                base = atomicAdd(&p.ctr->n_jobs[p.level + 1], uint32_t(__popc(mamb)));
            }
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
                    if (FUSED) publish_job(p.jobs_out + slot, o, epoch);   // fields, fence, then the ready mark
                    else { o.pad = epoch; p.jobs_out[slot] = o; }
                } else {
                    atomicOr(&p.ctr->error, 2u);
                    atomicSub(&p.ctr->outstanding, 1u);   // never claimable: do not wait for it
                }
            }
        }
    }

}

}  // namespace fdev
