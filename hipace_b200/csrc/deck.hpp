// Input-deck reader: the subset of AMReX ParmParse + the HiPACE++ math parser that the hot-path
// decks use (src/utils/Parser.H:36-51, 316-395): "prefix.key = v1 v2 ..." lines, '#' comments,
// later entries override earlier ones (the reference takes overrides on the command line),
// every numeric value is an expression over + - * / ^ ( ), my_constants.* and the built-in
// constants, and function-valued entries such as plasma.density(x,y,z) compile to a small RPN
// byte-code that a device kernel can evaluate per particle.
#pragma once
#include <cmath>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace hpb {

enum RpnOp : int { OP_CONST = 0, OP_VAR, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_NEG,
                   OP_SQRT, OP_EXP, OP_LOG, OP_SIN, OP_COS, OP_TANH, OP_ABS, OP_LT, OP_GT,
                   OP_MIN, OP_MAX };
struct RpnInstr { int op; int var; double val; };
constexpr int kMaxRpn = 64;

struct Deck {
    std::map<std::string, std::vector<std::string>> kv;
    // keys the reader has asked for (has / find / num / str ...): whatever is left over after
    // read_deck is an option this implementation does not know, and the run must not go on with
    // different physics than the deck asks for (the reference aborts on unused parameters too,
    // amrex::ParmParse::QueryUnusedInputs + hipace's "unused_param" check)
    mutable std::set<std::string> used;

    void parse(const std::string &text)
    {
        std::istringstream is(text);
        std::string line;
        while (std::getline(is, line)) {
            const size_t h = line.find('#');
            if (h != std::string::npos) line.erase(h);
            const size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            std::string key = trim(line.substr(0, eq));
            std::string val = line.substr(eq + 1);
            if (key.empty()) continue;
            std::vector<std::string> toks;
            size_t i = 0;
            while (i < val.size()) {
                while (i < val.size() && isspace((unsigned char)val[i])) ++i;
                if (i >= val.size()) break;
                if (val[i] == '"') {
                    const size_t e = val.find('"', i + 1);
                    toks.push_back(val.substr(i + 1, (e == std::string::npos ? val.size() : e) - i - 1));
                    i = (e == std::string::npos) ? val.size() : e + 1;
                } else {
                    size_t e = i;
                    while (e < val.size() && !isspace((unsigned char)val[e])) ++e;
                    toks.push_back(val.substr(i, e - i));
                    i = e;
                }
            }
            kv[key] = toks;
        }
    }

    bool has(const std::string &k) const { used.insert(k); return kv.count(k) != 0; }
    const std::vector<std::string> *find(const std::string &k, const std::string &alt = "") const
    {
        used.insert(k);
        if (!alt.empty()) used.insert(alt);
        auto it = kv.find(k);
        if (it != kv.end()) return &it->second;
        if (!alt.empty()) { it = kv.find(alt); if (it != kv.end()) return &it->second; }
        return nullptr;
    }
    // keys of the deck nobody asked for, except those under the given prefixes
    std::vector<std::string> unused(const std::vector<std::string> &allow_prefix) const
    {
        std::vector<std::string> out;
        for (auto &e : kv) {
            if (used.count(e.first)) continue;
            bool ok = false;
            for (auto &p : allow_prefix) if (e.first.compare(0, p.size(), p) == 0) ok = true;
            if (!ok) out.push_back(e.first);
        }
        return out;
    }
    std::string str(const std::string &k, const std::string &def, const std::string &alt = "") const
    {
        auto v = find(k, alt);
        return (v && !v->empty()) ? (*v)[0] : def;
    }
    std::vector<std::string> strs(const std::string &k) const
    {
        auto v = find(k);
        return v ? *v : std::vector<std::string>{};
    }
    double num(const std::string &k, double def, const std::string &alt = "") const
    {
        // a scalar: the whole right-hand side is ONE expression, blanks included ("0.01 * ne");
        // the reference's getWithParser joins the tokens the same way (utils/Parser.H)
        auto v = find(k, alt);
        if (!v || v->empty()) return def;
        std::string joined;
        for (auto &t : *v) joined += t;
        return eval(joined);
    }
    std::vector<double> nums(const std::string &k, const std::vector<double> &def,
                             const std::string &alt = "") const
    {
        auto v = find(k, alt);
        if (!v) return def;
        std::vector<double> out;
        for (auto &t : *v) out.push_back(eval(t));
        return out;
    }

    // ---- expression parser -----------------------------------------------------------------
    double eval(const std::string &expr) const
    {
        std::vector<RpnInstr> code;
        compile(expr, {}, code);
        return run(code, nullptr);
    }

    void compile(const std::string &expr, const std::vector<std::string> &vars,
                 std::vector<RpnInstr> &code) const
    {
        P p{this, expr, 0, &vars, &code, 0};
        p.expr();
        p.skip();
        if (p.i != expr.size()) throw std::runtime_error("parser: trailing characters in '" + expr + "'");
        if (code.size() > (size_t)kMaxRpn) throw std::runtime_error("parser: expression too long: " + expr);
    }

    static double run(const std::vector<RpnInstr> &code, const double *vars)
    {
        double st[kMaxRpn];
        int sp = 0;
        for (auto &c : code) {
            switch (c.op) {
            case OP_CONST: st[sp++] = c.val; break;
            case OP_VAR: st[sp++] = vars ? vars[c.var] : 0.0; break;
            case OP_ADD: --sp; st[sp - 1] += st[sp]; break;
            case OP_SUB: --sp; st[sp - 1] -= st[sp]; break;
            case OP_MUL: --sp; st[sp - 1] *= st[sp]; break;
            case OP_DIV: --sp; st[sp - 1] /= st[sp]; break;
            case OP_POW: --sp; st[sp - 1] = std::pow(st[sp - 1], st[sp]); break;
            case OP_LT: --sp; st[sp - 1] = st[sp - 1] < st[sp] ? 1.0 : 0.0; break;
            case OP_GT: --sp; st[sp - 1] = st[sp - 1] > st[sp] ? 1.0 : 0.0; break;
            case OP_MIN: --sp; st[sp - 1] = std::fmin(st[sp - 1], st[sp]); break;
            case OP_MAX: --sp; st[sp - 1] = std::fmax(st[sp - 1], st[sp]); break;
            case OP_NEG: st[sp - 1] = -st[sp - 1]; break;
            case OP_SQRT: st[sp - 1] = std::sqrt(st[sp - 1]); break;
            case OP_EXP: st[sp - 1] = std::exp(st[sp - 1]); break;
            case OP_LOG: st[sp - 1] = std::log(st[sp - 1]); break;
            case OP_SIN: st[sp - 1] = std::sin(st[sp - 1]); break;
            case OP_COS: st[sp - 1] = std::cos(st[sp - 1]); break;
            case OP_TANH: st[sp - 1] = std::tanh(st[sp - 1]); break;
            case OP_ABS: st[sp - 1] = std::fabs(st[sp - 1]); break;
            }
        }
        return sp > 0 ? st[sp - 1] : 0.0;
    }

private:
    static std::string trim(const std::string &s)
    {
        size_t a = 0, b = s.size();
        while (a < b && isspace((unsigned char)s[a])) ++a;
        while (b > a && isspace((unsigned char)s[b - 1])) --b;
        return s.substr(a, b - a);
    }

    struct P {
        const Deck *d; const std::string &s; size_t i; const std::vector<std::string> *vars;
        std::vector<RpnInstr> *code; int depth;
        void skip() { while (i < s.size() && isspace((unsigned char)s[i])) ++i; }
        void emit(int op, int var = 0, double val = 0.) { code->push_back({op, var, val}); }
        void expr()
        {
            cmp();
        }
        void cmp()
        {
            sum();
            skip();
            while (i < s.size() && (s[i] == '<' || s[i] == '>')) {
                const char c = s[i++];
                sum();
                emit(c == '<' ? OP_LT : OP_GT);
                skip();
            }
        }
        void sum()
        {
            term();
            skip();
            while (i < s.size() && (s[i] == '+' || s[i] == '-')) {
                const char c = s[i++];
                term();
                emit(c == '+' ? OP_ADD : OP_SUB);
                skip();
            }
        }
        void term()
        {
            unary();
            skip();
            while (i < s.size() && (s[i] == '*' || s[i] == '/')) {
                const char c = s[i++];
                unary();
                emit(c == '*' ? OP_MUL : OP_DIV);
                skip();
            }
        }
        void unary()
        {
            skip();
            if (i < s.size() && s[i] == '-') { ++i; unary(); emit(OP_NEG); return; }
            if (i < s.size() && s[i] == '+') { ++i; unary(); return; }
            power();
        }
        void power()
        {
            atom();
            skip();
            if (i < s.size() && (s[i] == '^' || (s[i] == '*' && i + 1 < s.size() && s[i + 1] == '*'))) {
                i += (s[i] == '^') ? 1 : 2;
                unary();            // right associative
                emit(OP_POW);
            }
        }
        void atom()
        {
            skip();
            if (i >= s.size()) throw std::runtime_error("parser: unexpected end in '" + s + "'");
            if (s[i] == '(') {
                ++i; expr(); skip();
                if (i >= s.size() || s[i] != ')') throw std::runtime_error("parser: missing ) in '" + s + "'");
                ++i;
                return;
            }
            if (isdigit((unsigned char)s[i]) || s[i] == '.') {
                size_t n = 0;
                const double v = std::stod(s.substr(i), &n);
                i += n;
                emit(OP_CONST, 0, v);
                return;
            }
            if (isalpha((unsigned char)s[i]) || s[i] == '_') {
                size_t e = i;
                while (e < s.size() && (isalnum((unsigned char)s[e]) || s[e] == '_')) ++e;
                const std::string name = s.substr(i, e - i);
                i = e;
                skip();
                if (i < s.size() && s[i] == '(') {       // function call
                    ++i; expr(); skip();
                    int nargs = 1;
                    while (i < s.size() && s[i] == ',') { ++i; expr(); skip(); ++nargs; }
                    if (i >= s.size() || s[i] != ')') throw std::runtime_error("parser: missing ) in '" + s + "'");
                    ++i;
                    static const std::map<std::string, int> f1 = {
                        {"sqrt", OP_SQRT}, {"exp", OP_EXP}, {"log", OP_LOG}, {"sin", OP_SIN},
                        {"cos", OP_COS}, {"tanh", OP_TANH}, {"abs", OP_ABS}};
                    static const std::map<std::string, int> f2 = {{"min", OP_MIN}, {"max", OP_MAX}, {"pow", OP_POW}};
                    if (nargs == 1 && f1.count(name)) { emit(f1.at(name)); return; }
                    if (nargs == 2 && f2.count(name)) { emit(f2.at(name)); return; }
                    throw std::runtime_error("parser: unknown function " + name);
                }
                for (size_t v = 0; v < vars->size(); ++v)
                    if ((*vars)[v] == name) { emit(OP_VAR, (int)v); return; }
                static const std::map<std::string, double> consts = {
                    {"pi", 3.14159265358979323846}, {"clight", 299792458.}, {"epsilon0", 8.8541878128e-12},
                    {"mu0", 1.25663706212e-06}, {"q_e", 1.602176634e-19}, {"m_e", 9.1093837015e-31},
                    {"m_p", 1.67262192369e-27}, {"hbar", 1.054571817e-34}, {"r_e", 2.817940326204929e-15},
                    {"true", 1.}, {"false", 0.}};
                if (consts.count(name)) { emit(OP_CONST, 0, consts.at(name)); return; }
                auto it = d->kv.find("my_constants." + name);
                if (it != d->kv.end() && !it->second.empty()) {
                    if (++depth > 16) throw std::runtime_error("parser: recursive my_constants");
                    // the expression may contain blanks ("1. / kp_inv"): the reader split it
                    std::string joined;
                    for (auto &t : it->second) joined += t;
                    emit(OP_CONST, 0, d->eval(joined));
                    return;
                }
                throw std::runtime_error("parser: unknown symbol '" + name + "' in '" + s + "'");
            }
            throw std::runtime_error("parser: bad character in '" + s + "'");
        }
    };
};

}  // namespace hpb
