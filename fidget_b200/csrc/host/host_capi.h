// C ABI of the host-side tape front end (fh_*).  See tape.h for the mapping
// to the reference's host crates.  All functions return 0 on success or a
// negative error code; fh_last_error() gives the message for this thread.
#pragma once
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fh_context fh_context;   // expression arena (Context)
typedef struct fh_tape fh_tape;         // SSA + register tape + var map (VmData)

const char* fh_last_error(void);

int32_t fh_context_new(fh_context** out);
void fh_context_free(fh_context* ctx);
// Parses `.vm` text (Context::from_text); *root receives the last node
int32_t fh_context_from_text(fh_context* ctx, const char* text, uint32_t* root);
int32_t fh_constant(fh_context* ctx, float v, uint32_t* node);
// kind: 0=X 1=Y 2=Z 3=fresh anonymous variable (id returned in *var_id)
int32_t fh_var(fh_context* ctx, int32_t kind, uint32_t* node, uint64_t* var_id);
// opcode numbering = bytecode numbering (Neg=3 .. Ln=20)
int32_t fh_unary(fh_context* ctx, uint8_t opcode, uint32_t a, uint32_t* node);
// binary builders apply the same folding rules as Context::{add,mul,...}
// (Add=21 .. Or=32)
int32_t fh_binary(fh_context* ctx, uint8_t opcode, uint32_t a, uint32_t b, uint32_t* node);
int32_t fh_context_len(const fh_context* ctx, uint32_t* n);

int32_t fh_tape_build(const fh_context* ctx, const uint32_t* roots, uint32_t n_roots,
                      uint32_t n_regs, fh_tape** out);
void fh_tape_free(fh_tape* t);

typedef struct fh_tape_info {
    uint32_t ssa_len;       // SsaTape::len
    uint32_t asm_len;       // RegTape::len == Function::size()
    uint32_t slot_count;    // RegTape::slot_count
    uint32_t choice_count;
    uint32_t output_count;
    uint32_t n_vars;
    uint32_t n_regs;
    int32_t var_x, var_y, var_z;   // input slot of X/Y/Z or -1
} fh_tape_info;
int32_t fh_tape_get_info(const fh_tape* t, fh_tape_info* info);
// var kind (0..3) and id for input slot i
int32_t fh_tape_var(const fh_tape* t, uint32_t i, int32_t* kind, uint64_t* id);

// Packed bytecode (fidget_bytecode::Bytecode::new).  Call with words=NULL to
// query the word count.
int32_t fh_tape_bytecode(const fh_tape* t, int32_t repack, uint32_t* words, size_t cap,
                         size_t* n_words, uint8_t* reg_count, uint32_t* mem_count);
// Wire / on-disk form of a tape (the counterpart of serde on VmData, fidget-core/src/vm/data.rs:64): a
// little-endian blob that carries exactly what fc_tape_create needs.
//   "FTAP" | version u32 = 1 | reg_count u32 | mem_count u32 | n_vars u32 | n_outputs u32 | choice_count u32
//   | axis slots i32[3] (X, Y, Z; -1 = unused) | n_words u64 | bytecode words u32[n_words]
// Call with buf = NULL to query the size.
int32_t fh_tape_serialize(const fh_tape* t, uint8_t* buf, size_t cap, size_t* n_bytes);
// Human-readable dump of the register tape in evaluation order (or the SSA
// tape root-first when ssa != 0); returns bytes needed incl. NUL.
size_t fh_tape_dump(const fh_tape* t, int32_t ssa, char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
