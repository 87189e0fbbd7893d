// Host-side tape front end; see tape.h for the reference map.
#include "tape.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <map>
#include <sstream>
#include <stdexcept>

namespace fhost {

static const char* OP_NAMES[OP_COUNT] = {
    "Output", "Input", "Copy", "Neg", "Abs", "Recip", "Sqrt", "Square", "Floor",
    "Ceil", "Round", "Not", "Rand", "Sin", "Cos", "Tan", "Asin", "Acos", "Atan",
    "Exp", "Ln", "Add", "Sub", "Mul", "Div", "Atan2", "Compare", "Mix", "Mod",
    "Min", "Max", "And", "Or", "Mem"};

const char* opcode_name(uint8_t op) { return op < OP_COUNT ? OP_NAMES[op] : "?"; }

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

bool Clause::operator==(const Clause& o) const {
    return op == o.op && form == o.form && out == o.out && a == o.a && b == o.b &&
           idx == o.idx && f2u(imm) == f2u(o.imm);
}

std::string to_string(const Clause& c) {
    std::ostringstream s;
    s << opcode_name(c.op);
    if (c.op == OP_OUTPUT) s << " out[" << c.idx << "] <- " << c.a;
    else if (c.op == OP_INPUT) s << " " << c.out << " <- in[" << c.idx << "]";
    else if (c.op == OP_MEM && c.form == F_RI) s << " " << c.out << " <- mem[" << c.idx << "]";
    else if (c.op == OP_MEM) s << " mem[" << c.idx << "] <- " << c.a;
    else if (c.op == OP_COPY && c.form == F_RI) s << " " << c.out << " <- " << c.imm;
    else if (c.op == OP_COPY || is_unary(c.op)) s << " " << c.out << " <- " << c.a;
    else if (c.form == F_RR) s << " " << c.out << " <- " << c.a << ", " << c.b;
    else if (c.form == F_RI) s << " " << c.out << " <- " << c.a << ", #" << c.imm;
    else s << " " << c.out << " <- #" << c.imm << ", " << c.a;
    return s.str();
}

////////////////////////////////////////////////////////////////////////////
// VarMap

int VarMap::get(const Var& v) const {
    for (size_t i = 0; i < order.size(); ++i)
        if (order[i] == v) return int(i);
    return -1;
}
void VarMap::insert(const Var& v) {
    if (get(v) < 0) order.push_back(v);
}

////////////////////////////////////////////////////////////////////////////
// Scalar semantics used for constant folding
// (fidget-core/src/context/op.rs:48-94, types/float.rs:66-142, rng/mod.rs:8-33)

static uint32_t rng_hash(uint32_t v) {
    uint32_t state = v * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28) + 4)) ^ state) * 277803737u;
    return (word >> 22) ^ word;
}
static float rng_rand(uint32_t seed) {
    uint32_t h = rng_hash(seed);
    return u2f((h >> 9) | 0x3f800000u) - 1.0f;
}
static uint32_t rng_mix(uint32_t a, uint32_t b) { return rng_hash(a + rng_hash(b)); }

float Context::eval_unary(uint8_t op, float a) {
    switch (op) {
        case OP_NEG: return -a;
        case OP_ABS: return std::fabs(a);
        case OP_RECIP: return 1.0f / a;
        case OP_SQRT: return std::sqrt(a);
        case OP_SQUARE: return a * a;
        case OP_FLOOR: return std::floor(a);
        case OP_CEIL: return std::ceil(a);
        case OP_ROUND: return std::round(a);
        case OP_SIN: return std::sin(a);
        case OP_COS: return std::cos(a);
        case OP_TAN: return std::tan(a);
        case OP_ASIN: return std::asin(a);
        case OP_ACOS: return std::acos(a);
        case OP_ATAN: return std::atan(a);
        case OP_EXP: return std::exp(a);
        case OP_LN: return std::log(a);
        case OP_NOT: return a == 0.0f ? 1.0f : 0.0f;
        case OP_RAND: return rng_rand(f2u(a));
    }
    throw std::runtime_error("bad unary opcode");
}

float Context::eval_binary(uint8_t op, float a, float b) {
    switch (op) {
        case OP_ADD: return a + b;
        case OP_SUB: return a - b;
        case OP_MUL: return a * b;
        case OP_DIV: return a / b;
        case OP_ATAN2: return std::atan2(a, b);
        case OP_MIN:
            if (a < b) return a;
            if (b < a) return b;
            return (std::isnan(a) || std::isnan(b)) ? NAN : b;
        case OP_MAX:
            if (a > b) return a;
            if (b > a) return b;
            return (std::isnan(a) || std::isnan(b)) ? NAN : b;
        case OP_COMPARE:
            if (a < b) return -1.0f;
            if (a > b) return 1.0f;
            if (a == b) return 0.0f;
            return NAN;
        case OP_MOD: {
            float r = std::fmod(a, b);
            return r < 0.0f ? r + std::fabs(b) : r;
        }
        case OP_AND: return a == 0.0f ? a : b;
        case OP_OR: return a != 0.0f ? a : b;
        case OP_MIX: return u2f(rng_mix(f2u(a), f2u(b)));
    }
    throw std::runtime_error("bad binary opcode");
}

////////////////////////////////////////////////////////////////////////////
// Context

Node Context::intern(const Op& op) {
    // Key mirrors `Op: Hash + Eq` with OrderedFloat constants
    // (NaN == NaN, +0 == -0; first insertion wins).
    char buf[64];
    switch (op.kind) {
        case K_CONST: {
            uint32_t bits = std::isnan(op.value) ? 0x7fc00000u
                          : (op.value == 0.0f ? 0u : f2u(op.value));
            snprintf(buf, sizeof buf, "c%08x", bits);
            break;
        }
        case K_VAR: snprintf(buf, sizeof buf, "v%d:%llu", int(op.var.kind),
                             (unsigned long long)(op.var.kind == Var::V ? op.var.id : 0)); break;
        case K_UNARY: snprintf(buf, sizeof buf, "u%d:%u", int(op.op), op.a); break;
        case K_BINARY: snprintf(buf, sizeof buf, "b%d:%u:%u", int(op.op), op.a, op.b); break;
    }
    auto it = dedup_.find(buf);
    if (it != dedup_.end()) return it->second;
    Node n = Node(ops_.size());
    ops_.push_back(op);
    dedup_.emplace(buf, n);
    return n;
}

Node Context::constant(float f) { return intern(Op{K_CONST, 0, 0, 0, f, Var{}}); }
Node Context::var(Var v) { return intern(Op{K_VAR, 0, 0, 0, 0.f, v}); }
Node Context::fresh_var() { return var(Var{Var::V, next_var_++}); }

bool Context::is_const(Node n, float* v) const {
    if (ops_.at(n).kind != K_CONST) return false;
    if (v) *v = ops_[n].value;
    return true;
}

Node Context::unary(uint8_t op, Node a) {
    float c;
    if (is_const(a, &c)) return constant(eval_unary(op, c));
    return intern(Op{K_UNARY, op, a, 0, 0.f, Var{}});
}

Node Context::binary(uint8_t op, Node a, Node b) {
    float ca, cb;
    if (is_const(a, &ca) && is_const(b, &cb)) return constant(eval_binary(op, ca, cb));
    return intern(Op{K_BINARY, op, a, b, 0.f, Var{}});
}

Node Context::binary_commutative(uint8_t op, Node a, Node b) {
    return binary(op, std::min(a, b), std::max(a, b));
}

// `match get_const(n) { Ok(k) => .. }` with a float-literal pattern compares
// with `==`, so -0.0 matches 0.0 and NaN matches nothing.
static bool const_is(const Context& c, Node n, float k) {
    float v;
    return c.is_const(n, &v) && v == k;
}

Node Context::add(Node a, Node b) {
    if (a == b) return mul(a, constant(2.0f));
    if (const_is(*this, a, 0.0f)) return b;
    if (const_is(*this, b, 0.0f)) return a;
    return binary_commutative(OP_ADD, a, b);
}
Node Context::mul(Node a, Node b) {
    if (a == b) return square(a);
    if (const_is(*this, a, 1.0f)) return b;
    if (const_is(*this, b, 1.0f)) return a;
    if (const_is(*this, a, 0.0f)) return a;
    if (const_is(*this, b, 0.0f)) return b;
    return binary_commutative(OP_MUL, a, b);
}
Node Context::min(Node a, Node b) { return a == b ? a : binary_commutative(OP_MIN, a, b); }
Node Context::max(Node a, Node b) { return a == b ? a : binary_commutative(OP_MAX, a, b); }
Node Context::and_(Node a, Node b) {
    float v;
    if (is_const(a, &v)) return v == 0.0f ? a : b;
    return binary(OP_AND, a, b);
}
Node Context::or_(Node a, Node b) {
    float v;
    if (is_const(a, &v)) return v != 0.0f ? a : b;
    if (is_const(b, &v) && v == 0.0f) return a;
    return binary(OP_OR, a, b);
}
Node Context::sub(Node a, Node b) {
    if (const_is(*this, a, 0.0f)) return neg(b);
    if (const_is(*this, b, 0.0f)) return a;
    return binary(OP_SUB, a, b);
}
Node Context::div(Node a, Node b) {
    if (const_is(*this, a, 0.0f)) return a;
    if (const_is(*this, b, 1.0f)) return a;
    return binary(OP_DIV, a, b);
}
Node Context::atan2(Node y, Node x) { return binary(OP_ATAN2, y, x); }
Node Context::compare(Node a, Node b) { return binary(OP_COMPARE, a, b); }
Node Context::mix(Node a, Node b) { return binary(OP_MIX, a, b); }
Node Context::modulo(Node a, Node b) { return binary(OP_MOD, a, b); }

Node Context::from_text(const std::string& text) {
    std::map<std::string, Node> seen;
    bool any = false;
    Node last = 0;
    std::istringstream in(text);
    std::string line;
    static const std::map<std::string, uint8_t> UNARY = {
        {"abs", OP_ABS}, {"neg", OP_NEG}, {"sqrt", OP_SQRT}, {"square", OP_SQUARE},
        {"floor", OP_FLOOR}, {"ceil", OP_CEIL}, {"round", OP_ROUND}, {"sin", OP_SIN},
        {"cos", OP_COS}, {"tan", OP_TAN}, {"asin", OP_ASIN}, {"acos", OP_ACOS},
        {"atan", OP_ATAN}, {"ln", OP_LN}, {"not", OP_NOT}, {"rand", OP_RAND},
        {"exp", OP_EXP}};
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ls(line);
        std::string name, opcode;
        if (!(ls >> name >> opcode)) throw std::runtime_error("malformed line: " + line);
        auto pop = [&]() -> Node {
            std::string t;
            if (!(ls >> t)) throw std::runtime_error("missing argument: " + line);
            auto it = seen.find(t);
            if (it == seen.end()) throw std::runtime_error("unknown variable " + t);
            return it->second;
        };
        Node node;
        auto un = UNARY.find(opcode);
        if (opcode == "const") {
            std::string t;
            if (!(ls >> t)) throw std::runtime_error("missing constant: " + line);
            char* end = nullptr;
            float f = strtof(t.c_str(), &end);
            if (end == t.c_str() || *end) throw std::runtime_error("bad constant " + t);
            node = constant(f);
        } else if (opcode == "var-x") node = x();
        else if (opcode == "var-y") node = y();
        else if (opcode == "var-z") node = z();
        else if (un != UNARY.end()) node = unary(un->second, pop());
        else {
            // two-argument forms; arguments are read left to right
            static const std::map<std::string, int> BIN = {
                {"add", 0}, {"mul", 1}, {"min", 2}, {"max", 3}, {"div", 4}, {"atan2", 5},
                {"sub", 6}, {"compare", 7}, {"mod", 8}, {"and", 9}, {"or", 10}, {"mix", 11}};
            auto bi = BIN.find(opcode);
            if (bi == BIN.end()) throw std::runtime_error("unknown opcode " + opcode);
            Node a = pop();
            Node b = pop();
            switch (bi->second) {
                case 0: node = add(a, b); break;
                case 1: node = mul(a, b); break;
                case 2: node = min(a, b); break;
                case 3: node = max(a, b); break;
                case 4: node = div(a, b); break;
                case 5: node = atan2(a, b); break;
                case 6: node = sub(a, b); break;
                case 7: node = compare(a, b); break;
                case 8: node = modulo(a, b); break;
                case 9: node = and_(a, b); break;
                case 10: node = or_(a, b); break;
                default: node = mix(a, b); break;
            }
        }
        seen[name] = node;
        last = node;
        any = true;
    }
    if (!any) throw std::runtime_error("empty file");
    return last;
}

////////////////////////////////////////////////////////////////////////////
// SSA flattening

SsaTape SsaTape::build(const Context& ctx, const std::vector<Node>& roots, VarMap* vars_out) {
    const uint32_t NONE = 0xFFFFFFFFu;
    size_t n = ctx.size();
    // mapping: node -> register index, or "immediate"
    std::vector<uint32_t> reg(n, NONE);
    std::vector<uint8_t> declared(n, 0);
    std::vector<uint32_t> parent_count(n, 0);
    uint32_t slot_count = 0;
    VarMap vars;

    auto children = [&](const Context::Op& op, Node out[2]) -> int {
        if (op.kind == Context::K_BINARY) { out[0] = op.a; out[1] = op.b; return 2; }
        if (op.kind == Context::K_UNARY) { out[0] = op.a; return 1; }
        return 0;
    };

    // Pass 1: declare nodes (DFS from a stack), count parents, discover vars
    std::vector<Node> todo(roots.begin(), roots.end());
    while (!todo.empty()) {
        Node node = todo.back();
        todo.pop_back();
        if (node >= n) throw std::runtime_error("bad node");
        if (declared[node]) continue;
        declared[node] = 1;
        const auto& op = ctx.get(node);
        if (op.kind != Context::K_CONST) {
            if (op.kind == Context::K_VAR) vars.insert(op.var);
            reg[node] = slot_count++;
        }
        Node ch[2];
        int nc = children(op, ch);
        for (int i = 0; i < nc; ++i) {
            parent_count[ch[i]]++;
            todo.push_back(ch[i]);
        }
    }

    SsaTape t;
    t.output_count = uint32_t(roots.size());
    for (size_t i = 0; i < roots.size(); ++i) {
        Clause c;
        c.op = OP_OUTPUT;
        c.idx = uint32_t(i);
        if (reg[roots[i]] != NONE) {
            c.a = reg[roots[i]];
            t.tape.push_back(c);
        } else {
            uint32_t o = slot_count++;
            c.a = o;
            t.tape.push_back(c);
            Clause k;
            k.op = OP_COPY;
            k.form = F_RI;
            k.out = o;
            k.imm = ctx.get(roots[i]).value;
            t.tape.push_back(k);
        }
    }

    // Pass 2: emit a node once all of its parents have been emitted
    std::vector<uint8_t> seen(n, 0);
    todo.assign(roots.begin(), roots.end());
    while (!todo.empty()) {
        Node node = todo.back();
        todo.pop_back();
        if (parent_count[node] > 0 || seen[node]) continue;
        seen[node] = 1;
        const auto& op = ctx.get(node);
        Node ch[2];
        int nc = children(op, ch);
        for (int i = 0; i < nc; ++i) {
            todo.push_back(ch[i]);
            parent_count[ch[i]]--;
        }
        if (reg[node] == NONE) continue;  // constants become immediates
        Clause c;
        c.out = reg[node];
        switch (op.kind) {
            case Context::K_VAR:
                c.op = OP_INPUT;
                c.idx = uint32_t(vars.get(op.var));
                break;
            case Context::K_UNARY:
                if (reg[op.a] == NONE) throw std::runtime_error("cannot handle f(imm)");
                c.op = op.op;
                c.a = reg[op.a];
                break;
            case Context::K_BINARY: {
                c.op = op.op;
                bool la = reg[op.a] != NONE, lb = reg[op.b] != NONE;
                if (is_choice(op.op)) t.choice_count++;
                if (la && lb) {
                    c.form = F_RR;
                    c.a = reg[op.a];
                    c.b = reg[op.b];
                } else if (la) {
                    c.form = F_RI;
                    c.a = reg[op.a];
                    c.imm = ctx.get(op.b).value;
                } else if (lb) {
                    // (imm, reg): commutative ops fold into the RegImm form
                    bool comm = op.op == OP_ADD || op.op == OP_MUL || op.op == OP_MIN || op.op == OP_MAX;
                    if (op.op == OP_AND || op.op == OP_OR)
                        throw std::runtime_error("And/Or with immediate lhs must be collapsed");
                    c.form = comm ? F_RI : F_IR;
                    c.a = reg[op.b];
                    c.imm = ctx.get(op.a).value;
                } else {
                    throw std::runtime_error("cannot handle f(imm, imm)");
                }
                break;
            }
            default: break;
        }
        t.tape.push_back(c);
    }
    if (vars_out) *vars_out = vars;
    return t;
}

////////////////////////////////////////////////////////////////////////////
// Register allocation

RegAlloc::Lru::Lru(uint32_t n) : prev(n), next(n) {
    for (uint32_t i = 0; i < n; ++i) {
        next[i] = uint8_t((i + 1) % n);
        prev[i] = uint8_t(i == 0 ? n - 1 : i - 1);
    }
}
void RegAlloc::Lru::remove(uint8_t i) {
    uint8_t p = prev[i], nx = next[i];
    next[p] = nx;
    prev[nx] = p;
}
void RegAlloc::Lru::insert_before(uint8_t i, uint8_t nx) {
    uint8_t p = prev[nx];
    next[p] = i;
    prev[nx] = i;
    next[i] = nx;
    prev[i] = p;
}
void RegAlloc::Lru::poke(uint8_t i) {
    if (head == i) return;
    if (prev[head] != i) {
        remove(i);
        insert_before(i, head);
    }
    head = i;
}
uint8_t RegAlloc::Lru::pop() {
    head = prev[head];
    return head;
}

RegAlloc::RegAlloc(uint32_t n_regs, size_t ssa_len)
    : N(n_regs), allocations(ssa_len, UNASSIGNED), registers(n_regs, UNASSIGNED), lru(n_regs) {
    if (N == 0 || N > 255) throw std::runtime_error("register count must be 1..255");
    for (int r = int(N) - 1; r >= 0; --r) spare_registers.push_back(uint8_t(r));
}

RegTape RegAlloc::finalize() {
    RegTape t = std::move(out);
    out = RegTape{};
    return t;
}

uint32_t RegAlloc::get_memory() {
    if (!spare_memory.empty()) {
        uint32_t p = spare_memory.back();
        spare_memory.pop_back();
        return p;
    }
    uint32_t m = out.slot_count++;
    assert(m >= N);
    return m;
}

RegAlloc::Alloc RegAlloc::get_allocation(uint32_t n) {
    uint32_t i = allocations[n];
    if (i < N) {
        lru.poke(uint8_t(i));
        return {A_REG, i};
    }
    if (i == UNASSIGNED) return {A_NONE, 0};
    return {A_MEM, i};
}

uint8_t RegAlloc::get_register() {
    if (!spare_registers.empty()) {
        uint8_t r = spare_registers.back();
        spare_registers.pop_back();
        out.slot_count = std::max(out.slot_count, uint32_t(r) + 1);
        assert(registers[r] == UNASSIGNED);
        lru.poke(r);
        return r;
    }
    // Evict the least recently used register to a memory slot
    uint8_t r = lru.pop();
    uint32_t mem = get_memory();
    uint32_t prev_node = registers[r];
    allocations[prev_node] = mem;
    registers[r] = UNASSIGNED;
    Clause ld;
    ld.op = OP_MEM;
    ld.form = F_RI;
    ld.out = r;
    ld.idx = mem;
    out.tape.push_back(ld);
    return r;
}

void RegAlloc::rebind_register(uint32_t n, uint8_t reg) {
    assert(allocations[n] >= N);
    assert(registers[reg] != UNASSIGNED);
    allocations[registers[reg]] = UNASSIGNED;
    registers[reg] = n;
    allocations[n] = reg;
}
void RegAlloc::bind_register(uint32_t n, uint8_t reg) {
    assert(allocations[n] >= N);
    assert(registers[reg] == UNASSIGNED);
    registers[reg] = n;
    allocations[n] = reg;
}
void RegAlloc::release_reg(uint8_t reg) {
    uint32_t node = registers[reg];
    assert(node != UNASSIGNED);
    registers[reg] = UNASSIGNED;
    spare_registers.push_back(reg);
    allocations[node] = UNASSIGNED;
}
void RegAlloc::push_store(uint8_t reg, uint32_t mem) {
    Clause st;
    st.op = OP_MEM;
    st.form = F_IR;
    st.a = reg;
    st.idx = mem;
    out.tape.push_back(st);
    spare_memory.push_back(mem);
}

uint8_t RegAlloc::get_out_reg(uint32_t o) {
    Alloc a = get_allocation(o);
    switch (a.k) {
        case A_REG: return uint8_t(a.v);
        case A_MEM: {
            uint8_t r = get_register();
            push_store(r, a.v);
            bind_register(o, r);
            return r;
        }
        default: throw std::runtime_error("cannot have unassigned output");
    }
}

void RegAlloc::emit(Clause c, uint32_t o, uint32_t a, uint32_t b) {
    c.out = o;
    c.a = a;
    c.b = b;
    out.tape.push_back(c);
}

// One register argument (unary, CopyReg, RegImm / ImmReg forms)
void RegAlloc::op_one_arg(const Clause& c) {
    uint8_t rx = get_out_reg(c.out);
    Alloc a = get_allocation(c.a);
    switch (a.k) {
        case A_REG:
            assert(rx != a.v);
            emit(c, rx, a.v, 0);
            release_reg(rx);
            break;
        case A_MEM: {
            uint8_t ra = get_register();
            push_store(ra, a.v);
            emit(c, rx, ra, 0);
            release_reg(rx);
            bind_register(c.a, ra);
            break;
        }
        case A_NONE:
            emit(c, rx, rx, 0);
            rebind_register(c.a, rx);
            break;
    }
}

void RegAlloc::op_two_args(const Clause& c) {
    uint32_t lhs = c.a, rhs = c.b;
    uint8_t rx = get_out_reg(c.out);
    Alloc L = get_allocation(lhs);
    Alloc R = get_allocation(rhs);
    if (L.k == A_REG && R.k == A_REG) {
        emit(c, rx, L.v, R.v);
        release_reg(rx);
    } else if (L.k == A_MEM && R.k == A_REG) {
        uint8_t ra = get_register();
        push_store(ra, L.v);
        emit(c, rx, ra, R.v);
        release_reg(rx);
        bind_register(lhs, ra);
    } else if (L.k == A_REG && R.k == A_MEM) {
        uint8_t ra = get_register();
        push_store(ra, R.v);
        emit(c, rx, L.v, ra);
        release_reg(rx);
        bind_register(rhs, ra);
    } else if (L.k == A_MEM && R.k == A_MEM && lhs == rhs) {
        uint8_t ra = get_register();
        push_store(ra, L.v);
        emit(c, rx, ra, ra);
        release_reg(rx);
        bind_register(lhs, ra);
    } else if (L.k == A_MEM && R.k == A_MEM) {
        uint8_t ra = get_register();
        uint8_t rb = get_register();
        push_store(ra, L.v);
        push_store(rb, R.v);
        emit(c, rx, ra, rb);
        release_reg(rx);
        bind_register(lhs, ra);
        bind_register(rhs, rb);
    } else if (L.k == A_NONE && R.k == A_REG) {
        emit(c, rx, rx, R.v);
        rebind_register(lhs, rx);
    } else if (L.k == A_REG && R.k == A_NONE) {
        emit(c, rx, L.v, rx);
        rebind_register(rhs, rx);
    } else if (L.k == A_NONE && R.k == A_NONE && lhs == rhs) {
        emit(c, rx, rx, rx);
        rebind_register(lhs, rx);
    } else if (L.k == A_NONE && R.k == A_NONE) {
        uint8_t ra = get_register();
        emit(c, rx, rx, ra);
        rebind_register(lhs, rx);
        bind_register(rhs, ra);
    } else if (L.k == A_NONE && R.k == A_MEM) {
        uint8_t ra = get_register();
        assert(ra != rx && lhs != rhs);
        push_store(ra, R.v);
        emit(c, rx, rx, ra);
        rebind_register(lhs, rx);
        bind_register(rhs, ra);
    } else {  // L mem, R none
        uint8_t ra = get_register();
        assert(ra != rx && lhs != rhs);
        push_store(ra, L.v);
        emit(c, rx, ra, rx);
        bind_register(lhs, ra);
        rebind_register(rhs, rx);
    }
}

void RegAlloc::op_out_only(const Clause& c) {
    uint8_t rx = get_out_reg(c.out);
    emit(c, rx, 0, 0);
    release_reg(rx);
}

void RegAlloc::op_output(const Clause& c) {
    Alloc a = get_allocation(c.a);
    switch (a.k) {
        case A_REG: emit(c, 0, a.v, 0); break;
        case A_MEM: {
            uint8_t ra = get_register();
            push_store(ra, a.v);
            emit(c, 0, ra, 0);
            bind_register(c.a, ra);
            break;
        }
        case A_NONE: {
            uint8_t ra = get_register();
            emit(c, 0, ra, 0);
            bind_register(c.a, ra);
            break;
        }
    }
}

void RegAlloc::op(const Clause& c) {
    if (c.op == OP_OUTPUT) op_output(c);
    else if (c.op == OP_INPUT || (c.op == OP_COPY && c.form == F_RI)) op_out_only(c);
    else if (c.op == OP_COPY || is_unary(c.op)) op_one_arg(c);
    else if (is_binary(c.op)) {
        if (c.form == F_RR) op_two_args(c);
        else op_one_arg(c);
    } else throw std::runtime_error("bad opcode in SSA tape");
}

RegTape allocate_registers(const SsaTape& ssa, uint32_t n_regs) {
    RegAlloc alloc(n_regs, ssa.tape.size());
    for (const auto& c : ssa.tape) alloc.op(c);
    return alloc.finalize();
}

TapeData TapeData::build(const Context& ctx, const std::vector<Node>& roots, uint32_t n_regs) {
    TapeData d;
    d.n_regs = n_regs;
    d.ssa = SsaTape::build(ctx, roots, &d.vars);
    d.asm_ = allocate_registers(d.ssa, n_regs);
    return d;
}

////////////////////////////////////////////////////////////////////////////
// Bytecode

static void visit_regs(const Clause& c, uint32_t N, std::vector<uint32_t>& regs) {
    (void)N;
    regs.clear();
    if (c.op == OP_OUTPUT) regs.push_back(c.a);
    else if (c.op == OP_INPUT) regs.push_back(c.out);
    else if (c.op == OP_MEM) regs.push_back(c.form == F_RI ? c.out : c.a);
    else if (c.op == OP_COPY && c.form == F_RI) regs.push_back(c.out);
    else if (c.op == OP_COPY || is_unary(c.op) || c.form != F_RR) {
        regs.push_back(c.out);
        regs.push_back(c.a);
    } else {
        regs.push_back(c.out);
        regs.push_back(c.a);
        regs.push_back(c.b);
    }
}

Bytecode make_bytecode(const RegTape& t, uint32_t n_regs, bool repack) {
    std::vector<uint32_t> regs;
    std::vector<int> map(256, -1);
    if (repack) {
        std::vector<size_t> counts(256, 0);
        for (const auto& c : t.tape) {
            visit_regs(c, n_regs, regs);
            for (uint32_t r : regs) counts[r]++;
        }
        std::vector<std::pair<size_t, int>> order;
        for (int r = 0; r < 256; ++r)
            if (counts[r]) order.push_back({counts[r], r});
        std::sort(order.begin(), order.end(), [](auto& x, auto& y) {
            if (x.first != y.first) return x.first > y.first;
            return x.second < y.second;
        });
        for (size_t i = 0; i < order.size(); ++i) map[order[i].second] = int(i);
    } else {
        for (int r = 0; r < 256; ++r) map[r] = r;
    }

    Bytecode bc;
    bc.words = {0xFFFFFFFFu, 0u};
    auto reg = [&](uint32_t r) -> uint8_t {
        int m = map[r];
        if (m < 0 || m == 255) throw std::runtime_error("register 255 is reserved");
        bc.reg_count = std::max<uint8_t>(bc.reg_count, uint8_t(m + 1));
        return uint8_t(m);
    };
    for (auto it = t.tape.rbegin(); it != t.tape.rend(); ++it) {
        const Clause& c = *it;
        uint8_t w[4] = {c.op, 0xFF, 0xFF, 0xFF};
        uint32_t imm = 0xFF000000u;
        if (c.op == OP_INPUT) { w[1] = reg(c.out); imm = c.idx; }
        else if (c.op == OP_OUTPUT) { w[1] = reg(c.a); imm = c.idx; }
        else if (c.op == OP_MEM) {
            uint32_t slot = c.idx - n_regs;
            bc.mem_count = std::max(bc.mem_count, slot + 1);
            imm = slot;
            if (c.form == F_RI) w[1] = reg(c.out);
            else w[2] = reg(c.a);
        } else if (c.op == OP_COPY && c.form == F_RI) { w[1] = reg(c.out); imm = f2u(c.imm); }
        else if (c.op == OP_COPY || is_unary(c.op)) { w[1] = reg(c.out); w[2] = reg(c.a); }
        else if (c.form == F_RI) { w[1] = reg(c.out); w[2] = reg(c.a); imm = f2u(c.imm); }
        else if (c.form == F_IR) { w[1] = reg(c.out); w[3] = reg(c.a); imm = f2u(c.imm); }
        else { w[1] = reg(c.out); w[2] = reg(c.a); w[3] = reg(c.b); }
        bc.words.push_back(uint32_t(w[0]) | uint32_t(w[1]) << 8 | uint32_t(w[2]) << 16 | uint32_t(w[3]) << 24);
        bc.words.push_back(imm);
    }
    bc.words.push_back(0xFFFFFFFFu);
    bc.words.push_back(0xFFFFFFFFu);
    return bc;
}

}  // namespace fhost
