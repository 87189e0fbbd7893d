// The fused tile renderers: fc_render2d (pixel::render), fc_render3d (voxel::render), fc_merge_slabs.
#include <cstddef>

#include "capi_internal.h"

// TileSizesRef::new (fidget-raster/src/lib.rs:59-66)
int32_t pick_tile_sizes(const uint32_t* ts_in, uint32_t n_in, const uint32_t* dflt, uint32_t n_dflt,
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
static AxisMap axes_of(const fc_tape* t) { return AxisMap{t->ax[0], t->ax[1], t->ax[2]}; }

// ShapeVars: every non-axis input slot needs a value (MissingVar otherwise, shape/mod.rs:586-600)
int32_t bind_vars(const fc_tape* t, const float* values, uint32_t n_values, VarBind& vb) {
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

cudaEvent_t get_event(fc_ctx* c, size_t i) {
    while (c->events.size() <= i) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        c->events.push_back(ev);
    }
    return c->events[i];
}

// Tile interleave of the multi-GPU renders: the XY root tiles (tx, ty) of the band with
// tile_owner(tx, ty, stride) == offset, in row-major order, as ids (ty - row0) * roots_x + tx.
static void owned_tiles(uint32_t roots_x, uint32_t row0, uint32_t row1, uint32_t stride, uint32_t offset,
                        std::vector<uint32_t>& ids) {
    ids.clear();
    for (uint32_t ty = row0; ty < row1; ++ty)
        for (uint32_t tx = 0; tx < roots_x; ++tx)
            if (tile_owner(tx, ty, stride) == offset) ids.push_back((ty - row0) * roots_x + tx);
}
int32_t root_subset(fc_ctx* c, uint32_t roots_x, uint32_t row0, uint32_t row1, uint32_t stride, uint32_t offset,
                    cudaStream_t s, const uint32_t** d_list, uint32_t* n) {
    const uint32_t key[5] = {roots_x, row0, row1, stride, offset};
    if (memcmp(key, c->root_list_key, sizeof key) != 0 || !c->root_list.p) {
        std::vector<uint32_t> ids;
        owned_tiles(roots_x, row0, row1, stride, offset, ids);
        CU(cudaStreamSynchronize(s));   // a previous launch may still read the old list
        CU(c->root_list.ensure(std::max<size_t>(ids.size(), 1) * 4));
        if (!ids.empty()) CU(cudaMemcpy(c->root_list.p, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice));
        memcpy(c->root_list_key, key, sizeof key);
        c->root_list_n = uint32_t(ids.size());
    }
    *d_list = c->root_list.as<uint32_t>();
    *n = c->root_list_n;
    return FC_OK;
}

extern "C" {

int32_t fc_render2d(fc_ctx* c, const fc_tape* tape, const fc_render2d_cfg* cfg, float* out, fc_render_stats* stats) {
    if (!c || !tape || !cfg || !out) return fail(FC_ERR_INVALID, "null argument");
    if (cfg->width == 0 || cfg->height == 0) return fail(FC_ERR_INVALID, "empty image");
    if (tape->info.mem_count) return fail(FC_ERR_UNSUPPORTED, "renderers need a tape without memory spills (<= 255 registers)");
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
    const bool timing = (cfg->flags & FC_FLAG_TIMING) != 0;
    const bool async = (cfg->flags & FC_FLAG_ASYNC) != 0;
    const bool want_stats = stats != nullptr;
    cudaStream_t s = c->stream;
    const uint32_t* d_roots = nullptr;
    uint32_t n_list = 0;
    if (cfg->root_stride > 1) {
        if (cfg->root_offset >= cfg->root_stride) return fail(FC_ERR_INVALID, "root_offset must be below root_stride");
        if (!is_device_ptr(out)) return fail(FC_ERR_UNSUPPORTED, "tile-interleaved renders need a device image");
        if (int32_t lrc = root_subset(c, roots_x, row0, row1, cfg->root_stride, cfg->root_offset, s, &d_roots, &n_list)) return lrc;
    }
    const uint64_t n_roots = d_roots ? uint64_t(n_list) : uint64_t(roots_x) * roots_y;
    const bool serial_fill = env_int("FIDGET_B200_SERIAL_FILL", 0) != 0;

    // ---- scratch ----
    const int bps = env_int("FIDGET_B200_BLOCKS_PER_SM", 6);
    const int grid_blocks = c->sm_count * bps;
    const uint32_t choice_words = (tape->info.choice_count + 15) / 16 + 1;
    CU(c->choice_scratch.ensure(size_t(std::max(grid_blocks, tail_2d_blocks(c->sm_count))) * WARPS_PER_BLOCK * choice_words * 32 * 4));
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
    const uint32_t fmt = cfg->out_format;
    if (fmt > FC_OUT_RGBA8) return fail(FC_ERR_INVALID, "unknown out_format");
    if (fmt != FC_OUT_F32) {
        // the distance image lives in the context; `out` receives the derived format at the end
        if (cfg->root_stride > 1 || cfg->root_row_begin || cfg->root_row_end)
            return fail(FC_ERR_UNSUPPORTED, "out_format other than FC_OUT_F32 needs a whole-frame render");
        CU(c->image.ensure(img_bytes));
        dimg = c->image.as<float>();
    } else if (!out_dev) {
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
    if (++c->epoch == 0) c->epoch = 1;
    // experimental (FC_FLAG_FUSED_TAIL): every level after the root level, the leaf pixels and the fills as ONE
    // persistent launch draining a job queue (tail2d.cu); the default is one launch per stage
    const bool fused = L >= 2 && L - 1 <= TAIL_MAX_LEVELS && ((cfg->flags & FC_FLAG_FUSED_TAIL) || env_int("FIDGET_B200_FUSE", 0));
    Tail2DParams tail{};
    for (int l = 0; l < L; ++l) {
        LevelParams p{};
        p.level = l;
        p.epoch = c->epoch;
        p.tile = ts[l];
        p.n_axis = l ? ts[l - 1] / ts[l] : 0;
        p.is_last = (l == L - 1);
        p.pixel_perfect = cfg->pixel_perfect;
        p.root_mode = (l == 0);
        p.roots_x = roots_x; p.roots_y = roots_y; p.roots_z = 1;
        p.root_x0 = 0; p.root_y0 = row0 * T0; p.root_z0 = 0;
        p.root_list = d_roots; p.n_root_list = n_list;
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
        if (fused) {
            tail.fill_tile[l] = ts[l];
            tail.fills[l] = c->fills[l].as<FillRec>();
            tail.fill_cap[l] = uint32_t(level_tiles[l + 1]);
            if (l) { tail.lv[l - 1] = p; continue; }
        }
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
        if (!fused) {
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
        if (fused) {
            tail.n_levels = L - 1;
            tail.px = q;
            tail.epoch = c->epoch;
            tail.paint_fills = env_int("FIDGET_B200_TAIL_PAINTS", 0) ? 1 : 0;
            auto paint = [&](int l, cudaStream_t fs) {
                FillParams f{};
                f.tile = ts[l];
                f.width = cfg->width; f.height = cfg->height;
                f.fills = c->fills[l].as<FillRec>();
                f.n_fills = &c->counters.as<Counters>()->n_fills[l];
                f.out = dimg;
                launch_fill_2d(f, c->sm_count * 2, fs);
                ++launches;
            };
            if (!tail.paint_fills) {   // level-0 fills are final already: paint them beside the tail
                CU(cudaEventRecord(c->ev_fork[0], s));
                CU(cudaStreamWaitEvent(c->aux_stream, c->ev_fork[0], 0));
                paint(0, c->aux_stream);
            }
            CU(launch_tail_2d(tail, c->sm_count, s));
            if (!tail.paint_fills) {
                for (int l = 1; l < L; ++l) paint(l, s);
                CU(cudaEventRecord(c->ev_join, c->aux_stream));
                CU(cudaStreamWaitEvent(s, c->ev_join, 0));
            }
        } else {
            launch_pixels_2d(q, c->sm_count * env_int("FIDGET_B200_PIXEL_BLOCKS_PER_SM", 8), s);
        }
        ++launches;
    }
    if (!serial_fill && !fused) {
        CU(cudaEventRecord(c->ev_join, c->aux_stream));
        CU(cudaStreamWaitEvent(s, c->ev_join, 0));
    }
    if (timing) CU(cudaEventRecord(get_event(c, ev++), s));
    CU(cudaGetLastError());
    if (fmt != FC_OUT_F32) {
        const size_t fb = fmt == FC_OUT_MASK_U8 ? size_t(cfg->width) * cfg->height
                        : fmt == FC_OUT_BITMAP_1BIT ? size_t((cfg->width + 7) / 8) * cfg->height : img_bytes;
        uint8_t* dfmt = reinterpret_cast<uint8_t*>(out);
        if (!out_dev) {
            CU(c->fx_out.ensure(fb));
            dfmt = c->fx_out.as<uint8_t>();
        }
        if (fmt == FC_OUT_RGBA8) launch_to_rgba(0, dimg, uint64_t(cfg->width) * cfg->height, dfmt, s);
        else launch_to_mask(dimg, cfg->width, cfg->height, dfmt, fmt == FC_OUT_BITMAP_1BIT, s);
        ++launches;
        CU(cudaGetLastError());
        if (!out_dev) CU(cudaMemcpyAsync(out, dfmt, fb, cudaMemcpyDeviceToHost, s));
    } else if (!out_dev) {   // only the rows of the requested band are copied back
        const uint32_t by0 = std::min(row0 * T0, cfg->height), by1 = std::min(row1 * T0, cfg->height);
        if (by1 > by0)
            CU(cudaMemcpyAsync(out + size_t(by0) * cfg->width, dimg + size_t(by0) * cfg->width,
                               size_t(by1 - by0) * cfg->width * 4, cudaMemcpyDeviceToHost, s));
    }
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
            if (fused) {   // [0] = root level, [12] = the fused tail (levels 1.., leaf pixels, fills)
                cudaEventElapsedTime(&ms, c->events[0], c->events[1]);
                stats->stage_ms[0] = ms;
                cudaEventElapsedTime(&ms, c->events[1], c->events[3]);
                stats->stage_ms[12] = ms;
                // when the last job of each list finished inside the fused launch (device clock, ms after its first warp):
                // [1 .. L-1] interval levels, [L] leaf tiles, [13] level-0 fills .. (statistics of the experiment)
                if (h.culled[0]) {
                    for (int k = 1; k <= 2 * (L - 1) + 2 && k < 8; ++k)
                        stats->stage_ms[k] = h.culled[k] > h.culled[0] ? float(double(h.culled[k] - h.culled[0]) * 1e-6) : 0.0f;
                    // [8], [9]: latest start of a level-1 / level-2 job; [10], [11]: the longest such job
                    for (int k = 8; k <= 9; ++k)
                        stats->stage_ms[k] = h.culled[k] > h.culled[0] ? float(double(h.culled[k] - h.culled[0]) * 1e-6) : 0.0f;
                    stats->stage_ms[10] = float(double(h.culled[10]) * 1e-6);
                    stats->stage_ms[11] = float(double(h.culled[11]) * 1e-6);
                }
                cudaEventElapsedTime(&ms, c->events[0], c->events[3]);
                stats->stage_ms[15] = ms;
            } else {
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
    }
    return rc;
}

int32_t fc_render3d(fc_ctx* c, const fc_tape* tape, const fc_render3d_cfg* cfg, fc_geometry_pixel* out,
                    fc_render_stats* stats) {
    if (!c || !tape || !cfg || !out) return fail(FC_ERR_INVALID, "null argument");
    if (cfg->width == 0 || cfg->height == 0 || cfg->depth == 0) return fail(FC_ERR_INVALID, "empty volume");
    if (tape->info.mem_count) return fail(FC_ERR_UNSUPPORTED, "renderers need a tape without memory spills (<= 255 registers)");
    if (tape->info.n_outputs != 1) return fail(FC_ERR_INVALID, "ShapeTape has multiple outputs");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    static const uint32_t DFLT[5] = {128, 64, 32, 16, 8};
    std::vector<uint32_t> ts;
    int32_t rc = pick_tile_sizes(cfg->tile_sizes, cfg->n_tile_sizes, DFLT, 5, std::max(cfg->width, cfg->height), ts);
    if (rc) return rc;
    if (cfg->n_tile_sizes == 0 && !(cfg->flags & (FC_FLAG_FULL_LADDER | FC_FLAG_EXACT_CENSUS)) &&
        !env_int("FIDGET_B200_FULL_LADDER", 0)) {
        // Device ladder: every other size of the default one.  A warp then carries 32 children per pass instead of 8
        // and a whole level of launches, job records and tape writes disappears; the image cannot change, because a
        // child's interval result on its grandparent's tape equals the one on its parent's simplified tape (a choice
        // the parent's region decided is decided the same way on any sub-region, and the pruned branch never
        // contributed to the value) -- tests/test_gpu_parity.py compares the two ladders bit for bit.
        std::vector<uint32_t> fused(1, ts[0]);
        for (size_t i = 0; i + 1 < ts.size();) {
            const size_t nx = std::min(i + 2, ts.size() - 1);
            fused.push_back(ts[nx]);
            i = nx;
        }
        ts.swap(fused);
    }
    const int L = int(ts.size());
    const uint32_t T0 = ts[0];
    const uint32_t roots_x = (cfg->width + T0 - 1) / T0, roots_y_all = (cfg->height + T0 - 1) / T0;
    const uint32_t row0 = cfg->root_row_begin, row1 = cfg->root_row_end ? cfg->root_row_end : roots_y_all;
    if (row0 > row1 || row1 > roots_y_all) return fail(FC_ERR_INVALID, "bad root row band");
    const uint32_t roots_y = row1 - row0;
    const uint32_t band_y0 = std::min(row0 * T0, cfg->height), band_y1 = std::min(row1 * T0, cfg->height);
    const uint32_t z_begin = cfg->z_begin, z_end = cfg->z_end ? cfg->z_end : cfg->depth;
    if (z_begin % T0 || z_begin >= z_end || z_end > ((cfg->depth + T0 - 1) / T0) * T0)
        return fail(FC_ERR_INVALID, "z slab must start on a root-tile boundary inside the volume");
    const uint32_t roots_z = (std::min(z_end, ((cfg->depth + T0 - 1) / T0) * T0) - z_begin + T0 - 1) / T0;
    const bool timing = (cfg->flags & FC_FLAG_TIMING) != 0;
    const bool async = (cfg->flags & FC_FLAG_ASYNC) != 0;
    const bool want_stats = stats != nullptr;
    cudaStream_t s = c->stream;
    const uint32_t* d_roots = nullptr;
    uint32_t n_list = 0;
    if (cfg->root_stride > 1) {
        if (cfg->root_offset >= cfg->root_stride) return fail(FC_ERR_INVALID, "root_offset must be below root_stride");
        if (!is_device_ptr(out)) return fail(FC_ERR_UNSUPPORTED, "tile-interleaved renders need a device image");
        if (T0 % 8) return fail(FC_ERR_UNSUPPORTED, "tile-interleaved renders need a root tile edge that is a multiple of 8");
        if (int32_t lrc = root_subset(c, roots_x, row0, row1, cfg->root_stride, cfg->root_offset, s, &d_roots, &n_list)) return lrc;
    }
    const uint64_t n_roots = (d_roots ? uint64_t(n_list) : uint64_t(roots_x) * roots_y) * roots_z;
    if (n_roots > 0xfffffff0ull) return fail(FC_ERR_UNSUPPORTED, "volume too large");

    // persistent CTAs per SM: 6 for the upper levels, 8 for the last one (many short jobs; measured on prospero and bear:
    // last level 1.10 -> 0.97 ms and 0.72 -> 0.61 ms, the level before it is fastest at 6)
    const int bps = env_int("FIDGET_B200_BLOCKS_PER_SM", 6);
    const int bps_last = std::max(bps, env_int("FIDGET_B200_LAST_LEVEL_BLOCKS_PER_SM", 8));
    const int grid_blocks = c->sm_count * bps, grid_blocks_last = c->sm_count * bps_last;
    const uint32_t choice_words = (tape->info.choice_count + 15) / 16 + 1;
    CU(c->choice_scratch.ensure(size_t(grid_blocks_last) * WARPS_PER_BLOCK * choice_words * 32 * 4));
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
    const bool exact_census = (cfg->flags & FC_FLAG_EXACT_CENSUS) != 0;
    uint64_t cap_census = 0;
    if (exact_census) {
        if (!stats) return fail(FC_ERR_INVALID, "FC_FLAG_EXACT_CENSUS needs a stats struct");
        if (cfg->width % T0 || cfg->height % T0 || d_roots || row0 || row1 != roots_y_all || z_begin || z_end < cfg->depth)
            return fail(FC_ERR_UNSUPPORTED, "the exact census needs a whole-volume render of an image whose sides are multiples of the root tile");
        cap_census = n_roots;
        for (int l = 1; l < L; ++l) { const uint64_t r = ts[l - 1] / ts[l]; cap_census += level_cap[l] * r * r * r; }
        cap_census = std::min<uint64_t>(cap_census, 64ull << 20);
        CU(c->census.ensure(cap_census * sizeof(CensusRec)));
    }
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
    CU(cudaMemsetAsync(c->heightmap.as<char>() + size_t(band_y0) * cfg->width * 8, 0, size_t(band_y1 - band_y0) * cfg->width * 8, s));
    // occlusion map (16 x 16 pixel blocks); used when every tile size down to 16 is a multiple of 16
    const uint32_t occl_w = (cfg->width + 15) / 16, occl_h = (cfg->height + 15) / 16;
    bool use_occl = !env_int("FIDGET_B200_NO_CULL", 0);
    for (int l = 0; l < L; ++l) if (ts[l] >= 16 && ts[l] % 16) use_occl = false;
    if (use_occl) {
        CU(c->occl.ensure(size_t(occl_w) * occl_h * 4));
        CU(cudaMemsetAsync(c->occl.p, 0, size_t(occl_w) * occl_h * 4, s));
    }
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
        p.root_x0 = 0; p.root_y0 = row0 * T0; p.root_z0 = z_begin;
        p.root_list = d_roots; p.n_root_list = n_list;
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
        p.occl = use_occl ? c->occl.as<uint32_t>() : nullptr;
        p.occl_w = occl_w;
        p.occl_h = occl_h;
        p.cull = (use_occl && l >= 1 && ts[l - 1] >= 16u && ts[l - 1] <= 64u) ? 1u : 0u;   // parents made of 1, 4 or 16 blocks
        p.census = exact_census ? c->census.as<CensusRec>() : nullptr;
        p.cap_census = uint32_t(cap_census);
        p.vb = vb;
        int blocks = (l == L - 1 && l > 0) ? grid_blocks_last : grid_blocks;
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
        launch_voxels_3d(q, c->sm_count * env_int("FIDGET_B200_VOXEL_BLOCKS_PER_SM", 12), s);
        ++launches;
    }
    if (exact_census) {
        // the counts collected so far describe what the device evaluated; replace them by the reference's census,
        // judged against the final heightmap (k_census_3d)
        Stats* ds = c->stats.as<Stats>();
        CU(cudaMemsetAsync(ds, 0, offsetof(Stats, grads), s));                        // evaluated .. simplified, pixels
        CensusParams cp{};
        cp.recs = c->census.as<CensusRec>();
        cp.n_recs = &c->counters.as<Counters>()->n_census;
        cp.cap = uint32_t(cap_census);
        for (int l = 0; l < L; ++l) cp.tile[l] = ts[l];
        cp.last_level = L - 1;
        cp.heightmap = c->heightmap.as<unsigned long long>();
        cp.width = cfg->width;
        cp.stats = ds;
        launch_census_3d(cp, c->sm_count * 8, s);
        ++launches;
    }
    if (timing) CU(cudaEventRecord(get_event(c, ev++), s));
    {
        NormalParams q{};
        q.width = cfg->width; q.height = cfg->height; q.depth = cfg->depth;
        q.y0 = band_y0; q.y1 = band_y1;
        q.root_list = d_roots; q.n_root_list = n_list; q.roots_x = roots_x; q.root_tile = T0;
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
    if (!out_dev && band_y1 > band_y0)
        CU(cudaMemcpyAsync(reinterpret_cast<char*>(out) + size_t(band_y0) * cfg->width * 16,
                           static_cast<char*>(dimg) + size_t(band_y0) * cfg->width * 16, size_t(band_y1 - band_y0) * cfg->width * 16,
                           cudaMemcpyDeviceToHost, s));
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

uint32_t fc_tiles_per_rank(uint32_t width, uint32_t height, uint32_t root_tile, uint32_t n_ranks) {
    if (!root_tile || !n_ranks) return 0;
    const uint32_t rx = (width + root_tile - 1) / root_tile, ry = (height + root_tile - 1) / root_tile;
    uint32_t best = 0;
    std::vector<uint32_t> ids;
    for (uint32_t r = 0; r < n_ranks; ++r) {
        owned_tiles(rx, 0, ry, n_ranks, r, ids);
        best = std::max(best, uint32_t(ids.size()));
    }
    return best;
}

static int32_t tiles_copy(fc_ctx* c, const void* src, void* dst, uint32_t width, uint32_t height, uint32_t px_bytes,
                          uint32_t T, uint32_t n_ranks, int rank /* -1: unpack all */) {
    if (!c || !src || !dst) return fail(FC_ERR_INVALID, "null argument");
    if (px_bytes != 4 && px_bytes != 16) return fail(FC_ERR_INVALID, "px_bytes must be 4 or 16");
    if (!T || !n_ranks || (rank >= 0 && uint32_t(rank) >= n_ranks)) return fail(FC_ERR_INVALID, "bad tile interleave");
    if (!is_device_ptr(src) || !is_device_ptr(dst)) return fail(FC_ERR_INVALID, "fc_tiles_pack/unpack take device pointers");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const uint32_t rx = (width + T - 1) / T, ry = (height + T - 1) / T;
    const uint32_t per = fc_tiles_per_rank(width, height, T, n_ranks);
    // slot of every image tile inside the gathered buffer: owner * per + index in the owner's list
    const uint32_t key[4] = {rx, ry, n_ranks, T};
    if (memcmp(key, c->tile_slots_key, sizeof key) != 0 || !c->tile_slots.p) {
        std::vector<uint32_t> slots(size_t(rx) * ry), ids;
        for (uint32_t r = 0; r < n_ranks; ++r) {
            owned_tiles(rx, 0, ry, n_ranks, r, ids);
            for (size_t k = 0; k < ids.size(); ++k) slots[ids[k]] = r * per + uint32_t(k);
        }
        CU(cudaStreamSynchronize(c->stream));
        CU(c->tile_slots.ensure(slots.size() * 4));
        CU(cudaMemcpy(c->tile_slots.p, slots.data(), slots.size() * 4, cudaMemcpyHostToDevice));
        memcpy(c->tile_slots_key, key, sizeof key);
    }
    launch_tiles_copy(src, dst, width, height, px_bytes, T, rx, ry, c->tile_slots.as<uint32_t>(), n_ranks, per, rank, c->stream);
    CU(cudaGetLastError());
    return FC_OK;
}
int32_t fc_tiles_pack(fc_ctx* c, const void* image, uint32_t width, uint32_t height, uint32_t px_bytes, uint32_t root_tile,
                      uint32_t n_ranks, uint32_t rank, void* packed) {
    return tiles_copy(c, image, packed, width, height, px_bytes, root_tile, n_ranks, int(rank));
}
int32_t fc_tiles_unpack(fc_ctx* c, const void* gathered, uint32_t width, uint32_t height, uint32_t px_bytes, uint32_t root_tile,
                        uint32_t n_ranks, void* image) {
    return tiles_copy(c, gathered, image, width, height, px_bytes, root_tile, n_ranks, -1);
}

}  // extern "C"
