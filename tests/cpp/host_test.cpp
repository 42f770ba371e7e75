// CPU-only driver for include/ministark_host.hpp (tests/test_cpp_host.py): prints JSON for the Python side to compare.
//   host_test coin <hexseed> <lanes>          scripted draws / reseeds of PublicCoin
//   host_test fib <log_n>                     Air bookkeeping + compiled composition program of examples/fib
//   host_test expr <nbase> <lde_step> <log_ce>   postfix expression on stdin -> compiled program
//   host_test proof <seed> <lanes> <with_ext> synthetic Proof -> wire bytes (hex)
#include <cstdio>
#include <iostream>
#include <sstream>

#include "ministark_examples.hpp"
#include "ministark_verifier.hpp"

using namespace mshost;

static std::string hex(const Bytes &b) {
    static const char *d = "0123456789abcdef";
    std::string s;
    for (u8 c : b) { s.push_back(d[c >> 4]); s.push_back(d[c & 15]); }
    return s;
}
static Bytes unhex(const std::string &s) {
    Bytes b;
    for (size_t i = 0; i + 1 < s.size(); i += 2) b.push_back((u8)std::stoul(s.substr(i, 2), nullptr, 16));
    return b;
}
static std::string fq_json(const Fq &v) {
    std::ostringstream o;
    o << "[" << v.c[0] << "," << v.c[1] << "," << v.c[2] << "]";
    return o.str();
}
static void print_program(const Program &p) {
    std::cout << "{\"nregs\":" << p.nregs << ",\"code\":[";
    for (size_t i = 0; i < p.code.size(); i++)
        std::cout << (i ? "," : "") << "[" << p.code[i][0] << "," << p.code[i][1] << "," << p.code[i][2] << "," << p.code[i][3] << "]";
    std::cout << "],\"consts\":[";
    for (size_t i = 0; i < p.consts.size(); i++)
        std::cout << (i ? "," : "") << "[" << p.consts[i][0] << "," << p.consts[i][1] << "," << p.consts[i][2] << "]";
    std::cout << "],\"bindings\":[";
    for (size_t i = 0; i < p.bindings.size(); i++)
        std::cout << (i ? "," : "") << "[" << p.bindings[i].slot << "," << (int)p.bindings[i].kind << "," << p.bindings[i].index << "]";
    std::cout << "]}";
}

struct Lcg {   // the same generator the Python test uses
    u64 s;
    u64 next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return s >> 11; }
    u64 fp() { return next() % P; }
    Fq fq(int lanes) {
        if (lanes != 3) return Fq(fp());
        const u64 a = fp(), b = fp(), c = fp();
        return Fq(a, b, c);
    }
    Bytes digest() { Bytes b(32); for (auto &x : b) x = (u8)next(); return b; }
};

int main(int argc, char **argv) {
    const std::string cmd = argc > 1 ? argv[1] : "";
    if (cmd == "coin") {
        PublicCoin c(unhex(argv[2]), atoi(argv[3]));
        std::cout << "{\"draws\":[";
        for (int step = 0; step < 24; step++) {
            Fq v = c.draw();
            std::cout << (step ? "," : "") << fq_json(v);
            if (step % 5 == 1) c.reseed_with_digest(sha256({Bytes{(u8)step}}));
            if (step % 7 == 2) c.reseed_with_field_elements({v, v});
            if (step % 9 == 3) c.reseed_with_int((u64)step * 1234567);
        }
        std::cout << "],\"queries\":[";
        auto q = c.draw_queries(32, (u64)1 << 23);
        for (size_t i = 0; i < q.size(); i++) std::cout << (i ? "," : "") << q[i];
        std::cout << "],\"queries96\":[";
        q = c.draw_queries(5, 96);
        for (size_t i = 0; i < q.size(); i++) std::cout << (i ? "," : "") << q[i];
        std::cout << "],\"pow\":[";
        for (u64 n = 1; n < 64; n++) std::cout << (n > 1 ? "," : "") << (c.verify_proof_of_work(3, n) ? 1 : 0);
        std::cout << "],\"seed\":\"" << hex(c.seed) << "\"}\n";
    } else if (cmd == "fib") {
        const unsigned log_n = (unsigned)atoi(argv[2]);
        Air air(fib_air_config(), (u64)1 << log_n, ProofOptions{32, 4, 8, 8, 64});
        std::cout << "{\"ce_blowup\":" << air.ce_blowup_factor << ",\"num_challenges\":" << air.num_challenges()
                  << ",\"num_coeffs\":" << air.num_composition_constraint_coeffs() << ",\"trace_arguments\":[";
        auto ta = air.trace_arguments();
        for (size_t i = 0; i < ta.size(); i++) std::cout << (i ? "," : "") << "[" << ta[i].first << "," << ta[i].second << "]";
        std::cout << "],\"degrees\":[";
        for (size_t i = 0; i < air.constraints.size(); i++) {
            auto d = degree(air.g, air.constraints[i].id, ((u64)1 << log_n) - 1);
            std::cout << (i ? "," : "") << "[" << d.first << "," << d.second << "]";
        }
        std::cout << "],\"program\":";
        print_program(air.composition_program(8));
        std::cout << "}\n";
    } else if (cmd == "bf") {
        // host_test bf hello | host_test bf burner a b c | host_test bf src <program> <input>
        const std::string kind = argv[2];
        const std::string hello = "++++++++++[>+++++++>++++++++++>+++>+<<<<-]>++.>+.+++++++..+++.>++.<<+++++++++++++++.>.+++.------.--------.";
        std::string src = kind == "hello" ? hello : (kind == "burner" ? bf::cycle_burner(atoi(argv[3]), atoi(argv[4]), atoi(argv[5])) : std::string(argv[3]));
        Bytes input;
        if (kind == "src" && argc > 4) for (const char *c = argv[4]; *c; c++) input.push_back((u8)*c);
        const bf::VmTrace t = bf::simulate(src, input);
        Sha256 hs;
        for (u64 w : t.base) { Bytes le; put_u64_le(le, to_mont(w)); hs.update(le); }
        std::cout << "{\"n\":" << t.n << ",\"output\":\"" << hex(t.output) << "\",\"base_sha256\":\"" << hex(hs.finish()) << "\"";
        if (t.n <= 4096) {
            Air air(bf::air_config(src, input, t.output), t.n, ProofOptions{19, 16, 20, 16, 16});
            std::cout << ",\"ce_blowup\":" << air.ce_blowup_factor << ",\"nconstraints\":" << air.constraints.size() << ",\"num_challenges\":"
                      << air.num_challenges() << ",\"num_coeffs\":" << air.num_composition_constraint_coeffs() << ",\"trace_arguments\":[";
            auto ta = air.trace_arguments();
            for (size_t i = 0; i < ta.size(); i++) std::cout << (i ? "," : "") << "[" << ta[i].first << "," << ta[i].second << "]";
            std::cout << "],\"degrees\":[";
            for (size_t i = 0; i < air.constraints.size(); i++) {
                auto d = degree(air.g, air.constraints[i].id, t.n - 1);
                std::cout << (i ? "," : "") << "[" << d.first << "," << d.second << "]";
            }
            // hints for fixed challenges (i+1, i+2, i+3)
            std::vector<Fq> ch;
            for (u64 i = 0; i < 11; i++) ch.push_back(Fq(i + 1, i + 2, i + 3));
            std::cout << "],\"hints\":[";
            auto hints = bf::gen_hints(t.n, src, input, t.output, ch);
            for (size_t i = 0; i < hints.size(); i++) std::cout << (i ? "," : "") << fq_json(hints[i]);
            std::cout << "],\"program\":";
            print_program(air.composition_program(17));
        }
        std::cout << "}\n";
    } else if (cmd == "verify") {
        // host_test verify fib <claim> <bits>            proof hex on stdin
        // host_test verify bf <source> <output hex> <bits>
        std::string hexproof;
        std::cin >> hexproof;
        const Bytes proof = unhex(hexproof);
        try {
            std::vector<u64> pos;
            if (std::string(argv[2]) == "fib") {
                pos = verify(fib_air_config(), proof, {Fq(strtoull(argv[3], nullptr, 10))}, {}, (u32)atoi(argv[4]));
            } else {
                const std::string src = argv[3];
                const Bytes output = unhex(argv[4]);
                pos = verify(bf::air_config(src, {}, output), proof, {}, [&] {
                    Bytes o;
                    put_u64_le(o, src.size()); o.insert(o.end(), src.begin(), src.end());
                    put_u64_le(o, 0);
                    put_u64_le(o, output.size()); o.insert(o.end(), output.begin(), output.end());
                    return o; }(), (u32)atoi(argv[5]));
            }
            std::cout << "ok";
            for (u64 p : pos) std::cout << " " << p;
            std::cout << "\n";
        } catch (const std::exception &e) {
            std::cout << "error: " << e.what() << "\n";
        }
    } else if (cmd == "expr") {
        Graph g;
        std::vector<Expr> st;
        std::string tok;
        auto pop = [&]() { Expr e = st.back(); st.pop_back(); return e; };
        while (std::cin >> tok) {
            if (tok == "x") st.push_back(X(g));
            else if (tok == "c") { u64 v; std::cin >> v; st.push_back(Constant(g, v)); }
            else if (tok == "q") { u64 a, b, c; std::cin >> a >> b >> c; st.push_back(ConstantQ(g, Fq(a, b, c))); }
            else if (tok == "ch") { u64 i; std::cin >> i; st.push_back(Challenge(g, i)); }
            else if (tok == "h") { u64 i; std::cin >> i; st.push_back(Hint(g, i)); }
            else if (tok == "cc") { u64 i; std::cin >> i; st.push_back(CompositionCoeff(g, i)); }
            else if (tok == "t") { u64 c; long long o; std::cin >> c >> o; st.push_back(Trace(g, c, o)); }
            else if (tok == "neg") { Expr a = pop(); st.push_back(-a); }
            else if (tok == "pow") { u64 e; std::cin >> e; Expr a = pop(); st.push_back(a.pow(e)); }
            else if (tok == "dup") { int k; std::cin >> k; st.push_back(st[st.size() - 1 - k]); }
            else {
                Expr b = pop(), a = pop();
                st.push_back(tok == "add" ? a + b : (tok == "mul" ? a * b : a / b));
            }
        }
        print_program(compile_program(g, st.back().id, (u32)atoi(argv[2]), (u64)atoll(argv[3]), atoi(argv[4])));
        std::cout << "\n";
    } else if (cmd == "proof") {
        Lcg r{(u64)atoll(argv[2])};
        const int lanes = atoi(argv[3]);
        const bool ext = atoi(argv[4]) != 0;
        auto view = [&](int k) {
            MerkleView v;
            for (int i = 0; i < k; i++) v.nodes.push_back(r.digest());
            for (int i = 0; i < 3; i++) v.initial_leaves.push_back(r.digest());
            for (int i = 0; i < 2; i++) v.sibling_leaves.push_back(r.digest());
            v.height = 17;
            return v;
        };
        Proof p;
        p.options = ProofOptions{19, 16, 20, 16, 16};
        p.trace_len = 2048;
        p.base_trace_commitment = r.digest();
        p.has_extension = ext;
        if (ext) p.extension_trace_commitment = r.digest();
        p.composition_trace_commitment = r.digest();
        for (int l = 0; l < 2; l++) {
            LayerProof lp;
            for (int i = 0; i < 16 >> l; i++) lp.flattenend_rows.push_back(r.fq(lanes));
            lp.merkle_proof = view(5 - 5 * l);
            lp.commitment = r.digest();
            p.fri_proof.layers.push_back(lp);
        }
        for (int i = 0; i < 2; i++) p.fri_proof.remainder_coeffs.push_back(r.fq(lanes));
        p.pow_nonce = r.next();
        for (int i = 0; i < 6; i++) p.trace_queries.base_trace_values.push_back(r.fq(1));
        if (ext) for (int i = 0; i < 4; i++) p.trace_queries.extension_trace_values.push_back(r.fq(lanes));
        for (int i = 0; i < 4; i++) p.trace_queries.composition_trace_values.push_back(r.fq(lanes));
        p.trace_queries.base_trace_proof = view(7);
        p.trace_queries.has_extension = ext;
        if (ext) p.trace_queries.extension_trace_proof = view(6);
        p.trace_queries.composition_trace_proof = view(4);
        for (int i = 0; i < 5; i++) p.execution_trace_ood_evals.push_back(r.fq(lanes));
        for (int i = 0; i < 2; i++) p.composition_trace_ood_evals.push_back(r.fq(lanes));
        std::cout << hex(p.to_bytes(lanes)) << "\n";
    } else {
        fprintf(stderr, "usage: host_test coin|fib|expr|proof ...\n");
        return 2;
    }
    return 0;
}
