/* A client of the C ABI alone (what a cgo / JNI / Rust-FFI binding does): loads a .vm model with the
 * host front end (fh_*), uploads its bytecode (fc_tape_create), renders it with fc_render2d and writes
 * the inside/outside bitmap as a PBM file plus a checksum of the RawDistancePixel words.
 *
 *   gcc -std=c99 -Iinclude -Ifidget_b200/csrc/host examples/render2d.c -Lfidget_b200 -lfidget_cuda \
 *       -Wl,-rpath,$PWD/fidget_b200 -o render2d && ./render2d models/prospero.vm 1024 out.pbm
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fidget_cuda.h"
#include "host_capi.h"

#define CHECK_FH(call) do { if ((call) != 0) { fprintf(stderr, "%s: %s\n", #call, fh_last_error()); return 1; } } while (0)
#define CHECK_FC(call) do { if ((call) != FC_OK) { fprintf(stderr, "%s: %s\n", #call, fc_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s model.vm size [out.pbm]\n", argv[0]); return 2; }
    const uint32_t size = (uint32_t)atoi(argv[2]);
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* text = (char*)malloc((size_t)len + 1);
    if (fread(text, 1, (size_t)len, f) != (size_t)len) { fclose(f); return 1; }
    text[len] = 0;
    fclose(f);

    /* host front end: text -> expression graph -> SSA + register tape -> bytecode */
    fh_context* ctx = NULL;
    fh_tape* tape = NULL;
    uint32_t root = 0;
    CHECK_FH(fh_context_new(&ctx));
    CHECK_FH(fh_context_from_text(ctx, text, &root));
    CHECK_FH(fh_tape_build(ctx, &root, 1, 255, &tape));
    fh_tape_info ti;
    CHECK_FH(fh_tape_get_info(tape, &ti));
    size_t n_words = 0;
    uint8_t reg_count = 0;
    uint32_t mem_count = 0;
    CHECK_FH(fh_tape_bytecode(tape, 1, NULL, 0, &n_words, &reg_count, &mem_count));
    uint32_t* words = (uint32_t*)malloc(n_words * 4);
    CHECK_FH(fh_tape_bytecode(tape, 1, words, n_words, &n_words, &reg_count, &mem_count));

    /* device side */
    fc_ctx* cuda = NULL;
    fc_tape* dtape = NULL;
    CHECK_FC(fc_ctx_create(0, &cuda));
    CHECK_FC(fc_tape_create(cuda, words, n_words, reg_count, mem_count, ti.n_vars, ti.output_count, ti.choice_count, &dtape));
    CHECK_FC(fc_tape_set_axes(dtape, ti.var_x, ti.var_y, ti.var_z));

    fc_render2d_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.width = cfg.height = size;
    /* RegionSize::screen_to_world (render/region.rs:87-108) with an identity world_to_model */
    const float s = 2.0f / (float)size;
    const float m[16] = {s, 0, 0, -((float)size / 2.0f) * s,
                         0, -s, 0, ((float)size / 2.0f - 1.0f) * s,
                         0, 0, 1, 0,
                         0, 0, 0, 1};
    memcpy(cfg.mat, m, sizeof m);
    float* image = (float*)malloc((size_t)size * size * 4);
    fc_render_stats st;
    CHECK_FC(fc_render2d(cuda, dtape, &cfg, image, &st));

    /* RawDistancePixel::inside (pixel.rs:183-193) */
    uint64_t inside = 0, sum = 1469598103934665603ull;
    for (size_t i = 0; i < (size_t)size * size; ++i) {
        uint32_t bits;
        memcpy(&bits, &image[i], 4);
        const int is_nan = (bits & 0x7fffffffu) > 0x7f800000u;
        const int fill = is_nan && (bits & (0xFFu << 9)) == (0xF6u << 9);
        inside += (uint64_t)(fill ? (bits & 1u) : (image[i] < 0.0f));
        sum = (sum ^ bits) * 1099511628211ull;
    }
    printf("%s %ux%u: %u clauses, inside %llu px, fnv1a %016llx, %u kernel launches\n", argv[1], size, size, ti.asm_len,
           (unsigned long long)inside, (unsigned long long)sum, st.kernel_launches);
    if (argc > 3) {
        FILE* o = fopen(argv[3], "wb");
        if (!o) { perror(argv[3]); return 1; }
        fprintf(o, "P1\n%u %u\n", size, size);
        for (uint32_t y = 0; y < size; ++y) {
            for (uint32_t x = 0; x < size; ++x) {
                uint32_t bits;
                memcpy(&bits, &image[(size_t)y * size + x], 4);
                const int is_nan = (bits & 0x7fffffffu) > 0x7f800000u;
                const int fill = is_nan && (bits & (0xFFu << 9)) == (0xF6u << 9);
                fputc((fill ? (int)(bits & 1u) : (image[(size_t)y * size + x] < 0.0f)) ? '1' : '0', o);
            }
            fputc('\n', o);
        }
        fclose(o);
    }
    fc_tape_release(dtape);
    fc_ctx_destroy(cuda);
    fh_tape_free(tape);
    fh_context_free(ctx);
    free(image); free(words); free(text);
    return 0;
}
