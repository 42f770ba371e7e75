// ministark_examples.hpp — the reference's example AIRs on the C++ host layer (ministark_host.hpp).
//
//   fib         examples/fib/main.rs                 (AIR in ministark_host.hpp::fib_air_config; trace generator here)
//   brainfuck   examples/brainfuck/{vm,tables,constraints,air}.rs — the VM that produces the 17 base columns, the
//               48 constraints and the evaluation-argument hints.  C++ twin of ministark_b200/examples/brainfuck.py;
//               CPU-tested against it (tests/test_cpp_host.py): identical base trace, degrees, blow-up, and the
//               compiled composition program evaluates to the same values.
#pragma once
#include <string>

#include "ministark_host.hpp"

namespace mshost {

// ---------------------------------------------------------------- examples/fib: gen_trace (main.rs:175-222)
// column-major 8 x n Montgomery words; returns the last value of column 7 (the claim)
inline u64 fib_gen_trace(u64 n, std::vector<u64> &trace) {
    trace.assign(8 * n, 0);
    u64 v[8] = {1, 2, 0, 0, 0, 0, 0, 0}, last = 0;
    for (int i = 2; i < 8; i++) v[i] = mulm(v[i - 2], v[i - 1]);
    for (u64 r = 0; r < n; r++) {
        for (int c = 0; c < 8; c++) trace[(u64)c * n + r] = to_mont(v[c]);
        last = v[7];
        u64 w[8];
        w[0] = mulm(v[6], v[7]);
        w[1] = mulm(v[7], w[0]);
        for (int i = 2; i < 8; i++) w[i] = mulm(w[i - 2], w[i - 1]);
        for (int i = 0; i < 8; i++) v[i] = w[i];
    }
    return last;
}

// ---------------------------------------------------------------- examples/brainfuck
namespace bf {

// column indices (tables.rs): base 0..16, extension 17..25
enum : u32 { CYCLE, IP, CURR_INSTR, NEXT_INSTR, MP, MEM_VAL, MEM_VAL_INV, DUMMY, M_CYCLE, M_MP, M_MEM_VAL, M_DUMMY, I_IP, I_CURR_INSTR,
             I_NEXT_INSTR, IN_VALUE, OUT_VALUE, P_INSTR_PERM, P_MEM_PERM, P_IN_EVAL, P_OUT_EVAL, M_PERM, I_PROC_PERM, I_PROG_EVAL, IN_EVAL,
             OUT_EVAL };
enum : u64 { CH_A, CH_B, CH_C, CH_D, CH_E, CH_F, CH_ALPHA, CH_BETA, CH_GAMMA, CH_DELTA, CH_ETA };
enum : u64 { H_INSTRUCTION, H_INPUT, H_INPUT_OFFSET, H_OUTPUT, H_OUTPUT_OFFSET };
constexpr u64 OPCODES[8] = {'>', '<', '+', '-', '.', ',', '[', ']'};   // OpCode::VALUES order (vm.rs:23-33)

inline std::vector<u64> compile(const std::string &source) {   // vm.rs:49-66
    std::vector<u64> program;
    std::vector<size_t> stack;
    for (char ch : source) {
        if (std::string("><+-.,[]").find(ch) == std::string::npos) continue;
        program.push_back((u64)ch);
        if (ch == '[') {
            program.push_back(0);
            stack.push_back(program.size() - 1);
        } else if (ch == ']') {
            const size_t last = stack.back();
            stack.pop_back();
            program.push_back(last + 1);
            program[last] = program.size();
        }
    }
    return program;
}

struct VmTrace {
    u64 n = 0;                      // padded number of rows
    std::vector<u64> base;          // 17 x n column-major, CANONICAL integers (to_mont on upload)
    Bytes output;
    u64 at(u32 col, u64 row) const { return base[(u64)col * n + row]; }
};

// vm.rs:68-336: run the program, build processor / memory / instruction / input / output tables, pad to a power of two
inline VmTrace simulate(const std::string &source, const Bytes &input = {}) {
    const std::vector<u64> program = compile(source);
    auto get = [&](u64 i) { return i < program.size() ? program[i] : (u64)0; };
    std::vector<u8> tape(1024, 0);
    u64 cycle = 0, ip = 0, mp = 0, mem_val = 0, curr = program.at(0), nxt = get(1);
    size_t in_pos = 0;
    std::vector<std::array<u64, 8>> proc;
    std::vector<std::array<u64, 3>> instr;
    std::vector<u64> in_rows, out_rows;
    VmTrace t;
    for (u64 i = 0; i < program.size(); i++) instr.push_back({i, program[i], get(i + 1)});
    u64 inv_lut[256];
    inv_lut[0] = 0;
    for (u64 v = 1; v < 256; v++) inv_lut[v] = invm(v);
    auto push_state = [&]() {
        proc.push_back({cycle, ip, curr, nxt, mp, mem_val, inv_lut[mem_val], (u64)(curr == 0)});
        instr.push_back({ip, curr, nxt});
    };
    while (ip < program.size()) {
        push_state();
        switch ((char)curr) {
            case '[': ip = mem_val == 0 ? program[ip + 1] : ip + 2; break;
            case ']': ip = mem_val != 0 ? program[ip + 1] : ip + 2; break;
            case '<': ip++; mp--; break;
            case '>': ip++; mp++; break;
            case '+': ip++; tape.at(mp)++; break;
            case '-': ip++; tape.at(mp)--; break;
            case '.': ip++; t.output.push_back(tape.at(mp)); out_rows.push_back(tape.at(mp)); break;
            case ',': ip++; tape.at(mp) = input.at(in_pos++); in_rows.push_back(tape.at(mp)); break;
            default: throw std::runtime_error("unrecognized instruction");
        }
        cycle++;
        curr = get(ip);
        nxt = get(ip + 1);
        mem_val = tape.at(mp);
    }
    push_state();
    std::stable_sort(instr.begin(), instr.end(), [](const auto &a, const auto &b) { return a[0] < b[0]; });
    // derive_memory_rows (vm.rs:338-381): rows of real instructions sorted by (address, cycle), dummy rows fill cycle gaps
    std::vector<std::array<u64, 4>> mem;
    for (const auto &r : proc)
        if (r[CURR_INSTR] != 0) mem.push_back({r[CYCLE], r[MP], r[MEM_VAL], 0});
    std::stable_sort(mem.begin(), mem.end(), [](const auto &a, const auto &b) { return std::tie(a[1], a[0]) < std::tie(b[1], b[0]); });
    std::vector<std::array<u64, 4>> filled;
    for (size_t k = 0; k < mem.size(); k++) {
        filled.push_back(mem[k]);
        if (k + 1 < mem.size() && mem[k][1] == mem[k + 1][1])
            for (u64 cy = mem[k][0] + 1; cy < mem[k + 1][0]; cy++) filled.push_back({cy, mem[k][1], mem[k][2], 1});
    }
    mem.swap(filled);
    const u64 longest = std::max({(u64)proc.size(), (u64)mem.size(), (u64)instr.size(), (u64)in_rows.size(), (u64)out_rows.size()});
    const u64 n = ceil_power_of_two(longest);
    while (proc.size() < n) {
        const auto l = proc.back();
        proc.push_back({l[CYCLE] + 1, l[IP], 0, 0, l[MP], l[MEM_VAL], l[MEM_VAL_INV], 1});
    }
    while (mem.size() < n) {
        const auto l = mem.back();
        mem.push_back({l[0] + 1, l[1], l[2], 1});
    }
    const u64 last_ip = instr.back()[0];
    while (instr.size() < n) instr.push_back({last_ip, 0, 0});
    in_rows.resize(n, 0);
    out_rows.resize(n, 0);
    t.n = n;
    t.base.assign(17 * n, 0);
    for (u64 r = 0; r < n; r++) {
        for (u32 c = 0; c < 8; c++) t.base[(u64)c * n + r] = proc[r][c];
        for (u32 c = 0; c < 4; c++) t.base[(u64)(8 + c) * n + r] = mem[r][c];
        for (u32 c = 0; c < 3; c++) t.base[(u64)(12 + c) * n + r] = instr[r][c];
        t.base[(u64)IN_VALUE * n + r] = in_rows[r];
        t.base[(u64)OUT_VALUE * n + r] = out_rows[r];
    }
    return t;
}

// a * b * (3c + 5) + O(a) cycles in three nested count-down loops; no cell ever exceeds max(a, b, c)
inline std::string cycle_burner(unsigned a, unsigned b, unsigned c) {
    return std::string(a, '+') + "[>" + std::string(b, '+') + "[>" + std::string(c, '+') + "[-]<-]<-]";
}

// ---- constraints (constraints.rs), assembled as air.rs:77-135
inline Expr instr_zerofier(Graph &g, const Expr &instr) {
    Expr prod;
    for (int i = 0; i < 8; i++) {
        Expr t = instr - Constant(g, OPCODES[i]);
        prod = i == 0 ? t : prod * t;
    }
    return prod;
}
inline Expr if_not_instr(Graph &g, u64 which, const Expr &ind) {
    Expr prod;
    bool first = true;
    for (int i = 0; i < 8; i++)
        if (OPCODES[i] != which) {
            Expr t = ind - Constant(g, OPCODES[i]);
            prod = first ? t : prod * t;
            first = false;
        }
    return prod;
}

inline std::vector<Expr> constraints(Graph &g, u64 trace_len) {
    auto cur = [&](u32 c) { return Trace(g, c, 0); };
    auto nx = [&](u32 c) { return Trace(g, c, 1); };
    auto CH = [&](u64 i) { return Challenge(g, i); };
    auto H = [&](u64 i) { return Hint(g, i); };
    const Expr one = Constant(g, 1), two = one + one;
    // --- processor base
    std::vector<Expr> proc_boundary = {cur(CYCLE), cur(IP), cur(MP), cur(MEM_VAL), cur(MEM_VAL_INV), cur(DUMMY)};
    const Expr mem_val_is_zero = cur(MEM_VAL) * cur(MEM_VAL_INV) - one;
    const Expr ip_step = nx(IP) - cur(IP) - one, same_mp = nx(MP) - cur(MP), same_val = nx(MEM_VAL) - cur(MEM_VAL);
    struct Triple { Expr e[3]; bool has[3]; };
    auto triple = [&](u64 op) {
        Triple t{{ip_step, same_mp, same_val}, {true, true, true}};
        switch ((char)op) {
            case '>': t.e[1] = nx(MP) - cur(MP) - one; t.has[2] = false; break;
            case '<': t.e[1] = nx(MP) - cur(MP) + one; t.has[2] = false; break;
            case '+': t.e[2] = nx(MEM_VAL) - cur(MEM_VAL) - one; break;
            case '-': t.e[2] = nx(MEM_VAL) - cur(MEM_VAL) + one; break;
            case '.': t.has[2] = false; break;
            case ',': break;
            case '[': t.e[0] = cur(MEM_VAL) * (nx(IP) - cur(IP) - two) + mem_val_is_zero * (nx(IP) - cur(NEXT_INSTR)); break;
            default: t.e[0] = mem_val_is_zero * (nx(IP) - cur(IP) - two) + cur(MEM_VAL) * (nx(IP) - cur(NEXT_INSTR));
        }
        return t;
    };
    Expr acc[3];
    bool have[3] = {false, false, false};
    for (int i = 0; i < 8; i++) {
        const Expr deselector = if_not_instr(g, OPCODES[i], cur(CURR_INSTR));
        const Triple t = triple(OPCODES[i]);
        for (int k = 0; k < 3; k++)
            if (t.has[k]) {
                Expr term = deselector * t.e[k] * cur(CURR_INSTR);
                acc[k] = have[k] ? acc[k] + term : term;
                have[k] = true;
            }
    }
    std::vector<Expr> transition = {acc[0], acc[1], acc[2],
                                    nx(CYCLE) - cur(CYCLE) - one,
                                    cur(MEM_VAL) * mem_val_is_zero,
                                    cur(MEM_VAL_INV) * mem_val_is_zero,
                                    (nx(DUMMY) - one) * nx(DUMMY),
                                    instr_zerofier(g, cur(CURR_INSTR)) * (cur(DUMMY) - one) + cur(CURR_INSTR) * cur(DUMMY)};
    // --- processor extension
    auto instr_fp = [&](const Expr &ip, const Expr &ci, const Expr &ni) { return CH(CH_ALPHA) - CH(CH_A) * ip - CH(CH_B) * ci - CH(CH_C) * ni; };
    auto mem_fp = [&](const Expr &cy, const Expr &mp, const Expr &mv) { return CH(CH_BETA) - CH(CH_D) * cy - CH(CH_E) * mp - CH(CH_F) * mv; };
    const Expr i_fp = instr_fp(cur(I_IP), cur(I_CURR_INSTR), cur(I_NEXT_INSTR)), p_fp = instr_fp(cur(IP), cur(CURR_INSTR), cur(NEXT_INSTR));
    const Expr m_fp = mem_fp(cur(M_CYCLE), cur(M_MP), cur(M_MEM_VAL)), pm_fp = mem_fp(cur(CYCLE), cur(MP), cur(MEM_VAL));
    std::vector<Expr> pext_boundary = {cur(P_IN_EVAL), cur(P_OUT_EVAL)};
    std::vector<Expr> pext_terminal = {
        cur(I_CURR_INSTR) * (cur(DUMMY) - one) * (cur(I_PROC_PERM) * i_fp - cur(P_INSTR_PERM) * p_fp) +
            instr_zerofier(g, cur(I_CURR_INSTR)) * (cur(DUMMY) - one) * (cur(I_PROC_PERM) - cur(P_INSTR_PERM) * p_fp) +
            cur(I_CURR_INSTR) * cur(DUMMY) * (cur(I_PROC_PERM) * i_fp - cur(P_INSTR_PERM)) +
            instr_zerofier(g, cur(I_CURR_INSTR)) * cur(DUMMY) * (cur(I_PROC_PERM) - cur(P_INSTR_PERM)),
        (cur(M_DUMMY) - one) * (cur(DUMMY) - one) * (cur(M_PERM) * m_fp - cur(P_MEM_PERM) * pm_fp) +
            cur(M_DUMMY) * (cur(DUMMY) - one) * (cur(M_PERM) - cur(P_MEM_PERM) * pm_fp) +
            (cur(M_DUMMY) - one) * cur(DUMMY) * (cur(M_PERM) * m_fp - cur(P_MEM_PERM)) +
            cur(M_DUMMY) * cur(DUMMY) * (cur(M_PERM) - cur(P_MEM_PERM)),
        cur(P_IN_EVAL) - H(H_INPUT),
        cur(P_OUT_EVAL) - H(H_OUTPUT)};
    const Expr ci = cur(CURR_INSTR);
    std::vector<Expr> pext_transition = {
        ci * (cur(P_INSTR_PERM) * p_fp - nx(P_INSTR_PERM)) + cur(DUMMY) * (cur(P_INSTR_PERM) - nx(P_INSTR_PERM)),
        // (a product where a sum is meant, as in constraints.rs:214-224: vacuous, restated as is)
        ci * (cur(P_MEM_PERM) * pm_fp - nx(P_MEM_PERM)) * cur(DUMMY) * (cur(P_MEM_PERM) - nx(P_MEM_PERM)),
        ci * if_not_instr(g, ',', ci) * (nx(P_IN_EVAL) - CH(CH_GAMMA) * cur(P_IN_EVAL) - nx(MEM_VAL)) + (ci - Constant(g, ',')) * (nx(P_IN_EVAL) - cur(P_IN_EVAL)),
        ci * if_not_instr(g, '.', ci) * (nx(P_OUT_EVAL) - cur(P_OUT_EVAL) * CH(CH_DELTA) - cur(MEM_VAL)) + (ci - Constant(g, '.')) * (nx(P_OUT_EVAL) - cur(P_OUT_EVAL))};
    // --- memory
    std::vector<Expr> mem_boundary = {cur(M_CYCLE), cur(M_MP), cur(M_MEM_VAL)};
    const Expr dmp = nx(M_MP) - cur(M_MP);
    std::vector<Expr> mem_transition = {(dmp - one) * dmp,
                                        dmp * nx(M_MEM_VAL),
                                        (nx(M_DUMMY) - one) * nx(M_DUMMY),
                                        dmp * cur(M_DUMMY),
                                        (nx(M_MEM_VAL) - cur(M_MEM_VAL)) * cur(M_DUMMY),
                                        (dmp - one) * (nx(M_CYCLE) - cur(M_CYCLE) - one)};
    std::vector<Expr> mext_transition = {(nx(M_PERM) - cur(M_PERM) * m_fp) * (cur(M_DUMMY) - one) + (nx(M_PERM) - cur(M_PERM)) * cur(M_DUMMY)};
    // --- instruction
    std::vector<Expr> instr_boundary = {cur(I_IP)};
    const Expr dip = nx(I_IP) - cur(I_IP);
    std::vector<Expr> instr_transition = {(dip - one) * dip, (dip - one) * (nx(I_CURR_INSTR) - cur(I_CURR_INSTR)), (dip - one) * (nx(I_NEXT_INSTR) - cur(I_NEXT_INSTR))};
    std::vector<Expr> iext_boundary = {cur(I_PROG_EVAL) - CH(CH_A) * cur(I_IP) - CH(CH_B) * cur(I_CURR_INSTR) - CH(CH_C) * cur(I_NEXT_INSTR)};
    std::vector<Expr> iext_terminal = {cur(I_PROG_EVAL) - H(H_INSTRUCTION)};
    const Expr next_fp = instr_fp(nx(I_IP), nx(I_CURR_INSTR), nx(I_NEXT_INSTR));
    std::vector<Expr> iext_transition = {
        cur(I_CURR_INSTR) * (cur(I_IP) - nx(I_IP) + one) * (nx(I_PROC_PERM) - cur(I_PROC_PERM) * next_fp) +
            instr_zerofier(g, cur(I_CURR_INSTR)) * (nx(I_PROC_PERM) - cur(I_PROC_PERM)) + (cur(I_IP) - nx(I_IP)) * (cur(I_PROC_PERM) - nx(I_PROC_PERM)),
        (dip - one) * (nx(I_PROG_EVAL) - cur(I_PROG_EVAL)) +
            dip * (nx(I_PROG_EVAL) - cur(I_PROG_EVAL) * CH(CH_ETA) - CH(CH_A) * nx(I_IP) - CH(CH_B) * nx(I_CURR_INSTR) - CH(CH_C) * nx(I_NEXT_INSTR))};
    // --- input / output
    std::vector<Expr> in_boundary = {cur(IN_EVAL) - cur(IN_VALUE)}, out_boundary = {cur(OUT_EVAL) - cur(OUT_VALUE)};
    std::vector<Expr> in_terminal = {cur(IN_EVAL) - H(H_INPUT) * H(H_INPUT_OFFSET)}, out_terminal = {cur(OUT_EVAL) - H(H_OUTPUT) * H(H_OUTPUT_OFFSET)};
    std::vector<Expr> in_transition = {cur(IN_EVAL) * CH(CH_GAMMA) + nx(IN_VALUE) - nx(IN_EVAL)};
    std::vector<Expr> out_transition = {cur(OUT_EVAL) * CH(CH_DELTA) + nx(OUT_VALUE) - nx(OUT_EVAL)};

    const unsigned log_n = 63 - (unsigned)__builtin_clzll(trace_len);
    const Expr x = X(g), first = Constant(g, 1), last = Constant(g, powm(domain_generator(log_n), trace_len - 1));
    const Expr but_last = (x - last) / (x.pow(trace_len) - one);
    std::vector<Expr> out;
    for (const auto *grp : {&transition, &pext_transition, &mem_transition, &mext_transition, &instr_transition, &iext_transition, &in_transition, &out_transition})
        for (const Expr &c : *grp) out.push_back(c * but_last);
    for (const auto *grp : {&proc_boundary, &pext_boundary, &mem_boundary, &instr_boundary, &iext_boundary, &in_boundary, &out_boundary})
        for (const Expr &c : *grp) out.push_back(c / (x - first));
    for (const auto *grp : {&pext_terminal, &iext_terminal, &in_terminal, &out_terminal})
        for (const Expr &c : *grp) out.push_back(c / (x - last));
    return out;
}

// gen_hints (air.rs:34-75): instruction / input / output evaluation arguments and offsets
inline std::vector<Fq> gen_hints(u64 trace_len, const std::string &source, const Bytes &input, const Bytes &output, const std::vector<Fq> &ch) {
    auto io_terminal = [&](const Bytes &symbols, const Fq &challenge) {
        Fq acc;
        for (u8 s : symbols) acc = fq_add(fq_mul(challenge, acc), Fq((u64)s));
        return std::make_pair(acc, fq_pow(challenge, trace_len - symbols.size()));
    };
    const auto in = io_terminal(input, ch.at(CH_GAMMA)), out = io_terminal(output, ch.at(CH_DELTA));
    std::vector<u64> program = compile(source);
    program.push_back(0);
    Fq acc;
    for (u64 ip = 0; ip < program.size(); ip++) {
        const u64 nxt = ip + 1 < program.size() ? program[ip + 1] : 0;
        acc = fq_mul(acc, ch.at(CH_ETA));
        acc = fq_add(acc, fq_add(fq_scale(ch.at(CH_A), ip), fq_add(fq_scale(ch.at(CH_B), program[ip]), fq_scale(ch.at(CH_C), nxt))));
    }
    return {acc, in.first, in.second, out.first, out.second};
}

inline AirConfig air_config(const std::string &source, const Bytes &input, const Bytes &output) {
    AirConfig cfg;
    cfg.num_base_columns = 17;
    cfg.num_extension_columns = 9;
    cfg.fq_is_fp = false;
    cfg.constraints = [](Graph &g, u64 n) { return constraints(g, n); };
    cfg.gen_hints = [=](u64 n, const std::vector<Fq> &, const std::vector<Fq> &ch) { return gen_hints(n, source, input, output, ch); };
    return cfg;
}

}  // namespace bf
}  // namespace mshost
