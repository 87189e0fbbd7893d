// fc_octree_sample: the sampling half of fidget-mesh's Octree::build.
#include "capi_internal.h"

// Device half: runs the sampler into `dout` (device memory, `cap` leaves); *n_out = surface leaves found
// (FC_ERR_INVALID when it exceeds cap).  Takes the context lock.
int32_t octree_sample_device(fc_ctx* c, const fc_tape* tape, const fc_octree_cfg* cfg, OctreeLeaf* dout, uint64_t cap,
                             uint32_t* n_out_p, fc_octree_stats* stats) {
    static_assert(sizeof(fc_octree_leaf) == sizeof(OctreeLeaf) && sizeof(OctreeLeaf) == 348, "leaf layout");
    if (!c || !tape || !cfg || !n_out_p) return fail(FC_ERR_INVALID, "null argument");
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
    *n_out_p = n_out;
    int32_t rc = check_device_errors(c);
    if (n_out > cap) rc = fail(FC_ERR_INVALID, "leaf buffer too small: " + std::to_string(n_out) + " surface leaves");
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

extern "C" {

int32_t fc_octree_sample(fc_ctx* c, const fc_tape* tape, const fc_octree_cfg* cfg, fc_octree_leaf* out, uint64_t cap,
                         uint64_t* n_leaves, fc_octree_stats* stats) {
    if (!c || !tape || !cfg || !n_leaves || (!out && cap)) return fail(FC_ERR_INVALID, "null argument");
    const bool out_dev = is_device_ptr(out);
    OctreeLeaf* dout = reinterpret_cast<OctreeLeaf*>(out);
    if (!out_dev) {
        std::lock_guard<std::mutex> guard(c->mu);
        CU(cudaSetDevice(c->device));
        CU(c->image.ensure(std::max<uint64_t>(cap, 1) * sizeof(OctreeLeaf)));
        dout = c->image.as<OctreeLeaf>();
    }
    uint32_t n_out = 0;
    int32_t rc = octree_sample_device(c, tape, cfg, dout, cap, &n_out, stats);
    *n_leaves = n_out;
    if (!rc && !out_dev && n_out) CU(cudaMemcpy(out, dout, size_t(n_out) * sizeof(OctreeLeaf), cudaMemcpyDeviceToHost));
    return rc;
}

}  // extern "C"
