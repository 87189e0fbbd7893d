// Host-side check of the interpreters' dispatch tables (interp.cuh): compiled with nvcc, run on the CPU by
// tests/test_dispatch_table.py.  Prints one line per (opcode, form) byte: "<byte> <interval handler> <f32 handler>".
#include <cstdio>

#include "interp.cuh"

int main() {
    constexpr fdev::DopTable iv = fdev::make_dop_table(false), f32 = fdev::make_dop_table(true);
    std::printf("H_COUNT %u OP_COUNT %u\n", unsigned(fdev::H_COUNT), unsigned(fdev::OP_COUNT));
    for (int i = 0; i < 256; ++i) std::printf("%d %u %u\n", i, unsigned(iv.h[i]), unsigned(f32.h[i]));
    std::printf("names H_GENERIC=%u H_ADD_RR=%u H_SUB_RR=%u H_MUL_RR=%u H_MIN_RR=%u H_MAX_RR=%u H_NEG=%u H_ABS=%u H_SQRT=%u "
                "H_SQUARE=%u H_COPY_REG=%u H_COPY_IMM=%u H_DIV_RR=%u H_EXP=%u\n",
                unsigned(fdev::H_GENERIC), unsigned(fdev::H_ADD_RR), unsigned(fdev::H_SUB_RR), unsigned(fdev::H_MUL_RR),
                unsigned(fdev::H_MIN_RR), unsigned(fdev::H_MAX_RR), unsigned(fdev::H_NEG), unsigned(fdev::H_ABS),
                unsigned(fdev::H_SQRT), unsigned(fdev::H_SQUARE), unsigned(fdev::H_COPY_REG), unsigned(fdev::H_COPY_IMM),
                unsigned(fdev::H_DIV_RR), unsigned(fdev::H_EXP));
    std::printf("ops OP_COPY=%u OP_NEG=%u OP_ABS=%u OP_SQRT=%u OP_SQUARE=%u OP_EXP=%u OP_ADD=%u OP_SUB=%u OP_MUL=%u OP_DIV=%u "
                "OP_MIN=%u OP_MAX=%u\n",
                unsigned(fdev::OP_COPY), unsigned(fdev::OP_NEG), unsigned(fdev::OP_ABS), unsigned(fdev::OP_SQRT),
                unsigned(fdev::OP_SQUARE), unsigned(fdev::OP_EXP), unsigned(fdev::OP_ADD), unsigned(fdev::OP_SUB),
                unsigned(fdev::OP_MUL), unsigned(fdev::OP_DIV), unsigned(fdev::OP_MIN), unsigned(fdev::OP_MAX));
    return 0;
}
