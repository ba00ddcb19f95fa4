// Device-side evaluation of the RPN byte-code that hpb::Deck::compile() produces for
// function-valued deck entries (plasma density(x,y,z), beam external_E/B(x,y,z,t)); the role of
// amrex::ParserExecutor in the reference (src/utils/Parser.H:316-395).
#pragma once
#include "deck.hpp"

struct DevRpn { int n; hpb::RpnInstr code[hpb::kMaxRpn]; };

__host__ __device__ inline double rpn_eval(const DevRpn &p, double x, double y, double z, double t = 0.)
{
    double st[24];
    int sp = 0;
    for (int k = 0; k < p.n; ++k) {
        const hpb::RpnInstr c = p.code[k];
        switch (c.op) {
        case hpb::OP_CONST: st[sp++] = c.val; break;
        case hpb::OP_VAR: st[sp++] = c.var == 0 ? x : (c.var == 1 ? y : (c.var == 2 ? z : t)); break;
        case hpb::OP_ADD: --sp; st[sp - 1] += st[sp]; break;
        case hpb::OP_SUB: --sp; st[sp - 1] -= st[sp]; break;
        case hpb::OP_MUL: --sp; st[sp - 1] *= st[sp]; break;
        case hpb::OP_DIV: --sp; st[sp - 1] /= st[sp]; break;
        case hpb::OP_POW: --sp; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
        case hpb::OP_LT: --sp; st[sp - 1] = st[sp - 1] < st[sp] ? 1.0 : 0.0; break;
        case hpb::OP_GT: --sp; st[sp - 1] = st[sp - 1] > st[sp] ? 1.0 : 0.0; break;
        case hpb::OP_MIN: --sp; st[sp - 1] = fmin(st[sp - 1], st[sp]); break;
        case hpb::OP_MAX: --sp; st[sp - 1] = fmax(st[sp - 1], st[sp]); break;
        case hpb::OP_NEG: st[sp - 1] = -st[sp - 1]; break;
        case hpb::OP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
        case hpb::OP_EXP: st[sp - 1] = exp(st[sp - 1]); break;
        case hpb::OP_LOG: st[sp - 1] = log(st[sp - 1]); break;
        case hpb::OP_SIN: st[sp - 1] = sin(st[sp - 1]); break;
        case hpb::OP_COS: st[sp - 1] = cos(st[sp - 1]); break;
        case hpb::OP_TANH: st[sp - 1] = tanh(st[sp - 1]); break;
        case hpb::OP_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
        }
    }
    return sp > 0 ? st[sp - 1] : 0.0;
}

inline void rpn_from_code(DevRpn &dst, const std::vector<hpb::RpnInstr> &code)
{
    dst.n = (int)code.size();
    for (size_t k = 0; k < code.size(); ++k) dst.code[k] = code[k];
}
