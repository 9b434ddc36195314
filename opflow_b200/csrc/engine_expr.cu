// engine_expr.cu -- expression IR (parsed from the functor-type signature), Expr::prepare() range algebra,
// FieldAssigner::assign and rangeReduce on the device.
//   reference: src/Core/Expr/Expression.hpp:99-105 (prepare / contains), every Op::prepare (cited per node below),
//              src/Core/Loops/FieldAssigner.hpp:26-86, src/Core/Loops/RangeFor.hpp:87-121.
#include "engine.hpp"
#include <array>
#include <map>
#include <mutex>
#include <unordered_map>
#include <cstdint>
#include <cstdlib>
#include <algorithm>

namespace opfe {

    enum Kind {
        K_F, K_S, K_PAR,// K_PAR: cell-parity leaf Par<c> of the red-black smoother (engine_solver.cu); a scalar-like node
        K_ADD, K_SUB, K_MUL, K_DIV, K_MIN, K_MAX, K_POW, K_LT, K_LE, K_GT, K_GE, K_EQ, K_NE, K_AND, K_OR,
        K_NEG, K_POS, K_NOT, K_SQRT, K_ABS, K_EXP, K_LOG, K_SIN, K_COS, K_TAN, K_TANH, K_POW2,
        K_COND, K_UNI, K_BIN, K_FLC2N, K_FLN2C, K_CONV,// K_UNI / K_BIN: point-wise math functors without any special preparation (AMDS.hpp:40-89)
        K_D2C, K_D1C, K_D1DN, K_D1UP, K_WENODN, K_WENOUP, K_INTPC2N, K_INTPN2C
    };
    struct KindInfo {
        const char* name;
        int kind, nchild;
        bool has_axis;
    };
    static const KindInfo kinds[] = {
            {"F", K_F, 0, false},         {"S", K_S, 0, false},         {"Par", K_PAR, 0, false},         {"Add", K_ADD, 2, false},     {"Sub", K_SUB, 2, false},
            {"Mul", K_MUL, 2, false},     {"Div", K_DIV, 2, false},     {"Min", K_MIN, 2, false},     {"Max", K_MAX, 2, false},
            {"Pow", K_POW, 2, false},     {"Lt", K_LT, 2, false},       {"Le", K_LE, 2, false},       {"Gt", K_GT, 2, false},
            {"Ge", K_GE, 2, false},       {"Eq", K_EQ, 2, false},       {"Ne", K_NE, 2, false},       {"And", K_AND, 2, false},
            {"Or", K_OR, 2, false},       {"Neg", K_NEG, 1, false},     {"Pos", K_POS, 1, false},     {"Not", K_NOT, 1, false},
            {"Sqrt", K_SQRT, 1, false},   {"Abs", K_ABS, 1, false},     {"Exp", K_EXP, 1, false},     {"Log", K_LOG, 1, false},
            {"Sin", K_SIN, 1, false},     {"Cos", K_COS, 1, false},     {"Tan", K_TAN, 1, false},     {"Tanh", K_TANH, 1, false},
            {"Pow2", K_POW2, 1, false},   {"Cond", K_COND, 3, false},   {"D2C", K_D2C, 1, true},      {"D1C", K_D1C, 1, true},
            {"D1Dn", K_D1DN, 1, true},    {"D1Up", K_D1UP, 1, true},    {"WenoDn", K_WENODN, 1, true}, {"WenoUp", K_WENOUP, 1, true},
            {"IntpC2N", K_INTPC2N, 1, true}, {"IntpN2C", K_INTPN2C, 1, true}, {"Conv", K_CONV, 1, false},
            {"Adapt1", K_UNI, 1, false}, {"Adapt2", K_BIN, 2, false},// UniOpAdaptor / BinOpAdaptor<functor>: Adapt1<functorName,E>
            {"FlCentralC2N", K_FLC2N, 2, true}, {"FlCentralN2C", K_FLN2C, 2, true}, {"FlQuickC2N", K_FLC2N, 2, true}, {"FlQuickN2C", K_FLN2C, 2, true}, {"FlCuiC2N", K_FLC2N, 2, true}, {"FlCuiN2C", K_FLN2C, 2, true}, {"FlFrommC2N", K_FLC2N, 2, true}, {"FlFrommN2C", K_FLN2C, 2, true}, {"FlLuiC2N", K_FLC2N, 2, true}, {"FlLuiN2C", K_FLN2C, 2, true}, {"FlMinmodC2N", K_FLC2N, 2, true}, {"FlMinmodN2C", K_FLN2C, 2, true}, {"FlSuperbeeC2N", K_FLC2N, 2, true}, {"FlSuperbeeN2C", K_FLN2C, 2, true}, {"FlMusclC2N", K_FLC2N, 2, true}, {"FlMusclN2C", K_FLN2C, 2, true}, {"FlHarmonicC2N", K_FLC2N, 2, true}, {"FlHarmonicN2C", K_FLN2C, 2, true}, {"FlAlbadaC2N", K_FLC2N, 2, true}, {"FlAlbadaN2C", K_FLN2C, 2, true}, 
            {"Exp2", K_UNI, 1, false}, {"Expm1", K_UNI, 1, false}, {"Log10", K_UNI, 1, false}, {"Log2", K_UNI, 1, false}, {"Log1p", K_UNI, 1, false}, {"Cbrt", K_UNI, 1, false}, {"ASin", K_UNI, 1, false}, {"ACos", K_UNI, 1, false}, {"ATan", K_UNI, 1, false}, {"Sinh", K_UNI, 1, false}, {"Cosh", K_UNI, 1, false}, {"ASinh", K_UNI, 1, false}, {"ACosh", K_UNI, 1, false}, {"ATanh", K_UNI, 1, false}, {"Erf", K_UNI, 1, false}, {"Erfc", K_UNI, 1, false}, {"TGamma", K_UNI, 1, false}, {"LGamma", K_UNI, 1, false}, {"Ceil", K_UNI, 1, false}, {"Floor", K_UNI, 1, false}, {"Trunc", K_UNI, 1, false}, {"Round", K_UNI, 1, false}, {"LRound", K_UNI, 1, false}, {"LLRound", K_UNI, 1, false}, {"NearbyInt", K_UNI, 1, false}, {"Rint", K_UNI, 1, false}, {"LRint", K_UNI, 1, false}, {"LLRint", K_UNI, 1, false}, {"ILogb", K_UNI, 1, false}, {"Logb", K_UNI, 1, false}, {"FMod", K_BIN, 2, false}, {"Remainder", K_BIN, 2, false}, {"FDim", K_BIN, 2, false}, {"Hypot", K_BIN, 2, false}, {"ATan2", K_BIN, 2, false}, {"Ldexp", K_BIN, 2, false}, {"Scalbn", K_BIN, 2, false}, {"Scalbln", K_BIN, 2, false}, {"Nextafter", K_BIN, 2, false}, {"Nexttoward", K_BIN, 2, false}, {"Copysing", K_BIN, 2, false}};

    struct Node {
        int kind = 0, axis = -1, leaf = -1, nchild = 0;
        int child[3] = {-1, -1, -1};
        int conv[4] = {1, 1, 1, 0};// Conv<n0,n1,n2,k0,E>: kernel extents and the first scalar slot of its entries
        // prepared properties
        bool scalar = false;
        int loc[D3] = {0, 0, 0};
        Range acc, local, logical;
        const opf_field_s* src = nullptr;// field whose props were inherited (mesh etc.)
    };
    struct Tree {
        std::vector<Node> nodes;// preorder: node index == device-side B
        int nfields = 0, nscalars = 0;
    };

    struct Parser {
        const char* s;
        size_t pos = 0;
        std::string err;
        void ws() {
            while (s[pos] == ' ') ++pos;
        }
        bool parse_int(int& v) {
            ws();
            bool neg = false;
            if (s[pos] == '-') {
                neg = true;
                ++pos;
            }
            if (s[pos] < '0' || s[pos] > '9') return false;
            v = 0;
            while (s[pos] >= '0' && s[pos] <= '9') v = v * 10 + (s[pos++] - '0');
            if (neg) v = -v;
            return true;
        }
        int parse_node(Tree& t) {
            ws();
            size_t b = pos;
            while ((s[pos] >= 'A' && s[pos] <= 'Z') || (s[pos] >= 'a' && s[pos] <= 'z') || (s[pos] >= '0' && s[pos] <= '9')) ++pos;
            std::string name(s + b, pos - b);
            const KindInfo* ki = nullptr;
            for (const auto& k : kinds)
                if (name == k.name) ki = &k;
            if (!ki) {
                err = "unknown node '" + name + "'";
                return -1;
            }
            ws();
            if (s[pos] != '<') {
                err = "expected '<' after " + name;
                return -1;
            }
            ++pos;
            const int me = (int) t.nodes.size();
            t.nodes.emplace_back();
            t.nodes[me].kind = ki->kind;
            t.nodes[me].nchild = ki->nchild;
            if (ki->kind == K_F || ki->kind == K_S || ki->kind == K_PAR) {
                int v;
                if (!parse_int(v) || v < 0) {
                    err = "bad leaf index";
                    return -1;
                }
                t.nodes[me].leaf = v;
                t.nodes[me].scalar = ki->kind != K_F;
                if (ki->kind == K_F) t.nfields = std::max(t.nfields, v + 1);
                else if (ki->kind == K_S)
                    t.nscalars = std::max(t.nscalars, v + 1);
            } else {
                if (name == "Adapt1" || name == "Adapt2") {// the functor's name is part of the signature (registry key), not of the tree
                    ws();
                    const size_t b2 = pos;
                    while ((s[pos] >= 'A' && s[pos] <= 'Z') || (s[pos] >= 'a' && s[pos] <= 'z') || (s[pos] >= '0' && s[pos] <= '9') || s[pos] == '_') ++pos;
                    ws();
                    if (pos == b2 || s[pos] != ',') {
                        err = "Adapt<functorName, operands...> expected";
                        return -1;
                    }
                    ++pos;
                }
                if (ki->kind == K_CONV) {// Conv<n0, n1, n2, k0, E>
                    for (int q = 0; q < 4; ++q) {
                        int v;
                        if (!parse_int(v) || v < 0 || (q < 3 && (v < 1 || v > 2 * opf::WR + 1 || v % 2 == 0))) {
                            err = "Conv<n0,n1,n2,k0,E>: odd extents 1..7 and a scalar slot expected";
                            return -1;
                        }
                        t.nodes[me].conv[q] = v;
                        ws();
                        if (s[pos] != ',') {
                            err = "expected ','";
                            return -1;
                        }
                        ++pos;
                    }
                    t.nscalars = std::max(t.nscalars, t.nodes[me].conv[3] + t.nodes[me].conv[0] * t.nodes[me].conv[1] * t.nodes[me].conv[2]);
                }
                if (ki->has_axis) {
                    int v;
                    if (!parse_int(v) || v < 0 || v >= D3) {
                        err = "bad axis";
                        return -1;
                    }
                    t.nodes[me].axis = v;
                    ws();
                    if (s[pos] != ',') {
                        err = "expected ','";
                        return -1;
                    }
                    ++pos;
                }
                for (int c = 0; c < ki->nchild; ++c) {
                    if (c) {
                        ws();
                        if (s[pos] != ',') {
                            err = "expected ',' in " + name;
                            return -1;
                        }
                        ++pos;
                    }
                    int ch = parse_node(t);
                    if (ch < 0) return -1;
                    t.nodes[me].child[c] = ch;
                }
            }
            ws();
            if (s[pos] != '>') {
                err = "expected '>' closing " + name;
                return -1;
            }
            ++pos;
            return me;
        }
    };

    static int parse_signature(const char* sig, Tree& t) {
        Parser p{sig};
        if (p.parse_node(t) != 0) return fail(OPF_ERR_INVALID, "signature '%s': %s", sig, p.err.c_str());
        p.ws();
        if (sig[p.pos] != 0) return fail(OPF_ERR_INVALID, "signature '%s': trailing characters", sig);
        if ((int) t.nodes.size() > OPF_MAX_NODES) return fail(OPF_ERR_UNSUPPORTED, "expression has %zu nodes (max %d)", t.nodes.size(), OPF_MAX_NODES);
        if (t.nfields > OPF_MAX_FIELDS || t.nscalars > OPF_MAX_SCALARS) return fail(OPF_ERR_UNSUPPORTED, "too many leaves");
        return OPF_OK;
    }

    static void inherit(Node& n, const Node& a) {
        n.scalar = a.scalar;
        for (int d = 0; d < D3; ++d) n.loc[d] = a.loc[d];
        n.acc = a.acc;
        n.local = a.local;
        n.logical = a.logical;
        n.src = a.src;
    }

    // post-order prepare (Expression.hpp:99-103)
    static int prepare_node(Tree& t, int id, const opf_field_t* fields, int nfields) {
        Node& n = t.nodes[id];
        for (int c = 0; c < n.nchild; ++c)
            if (int rc = prepare_node(t, n.child[c], fields, nfields)) return rc;
        auto ch = [&](int c) -> Node& { return t.nodes[n.child[c]]; };
        switch (n.kind) {
            case K_F: {
                if (n.leaf >= nfields || !fields[n.leaf]) return fail(OPF_ERR_INVALID, "expression needs field argument %d", n.leaf);
                const opf_field_s* f = fields[n.leaf];
                for (int d = 0; d < D3; ++d) n.loc[d] = f->loc[d];
                n.acc = f->accessible;
                n.local = f->local;
                n.logical = f->logical;
                n.src = f;
                break;
            }
            case K_S:
            case K_PAR: n.scalar = true; break;
            case K_COND: {
                // CondOp::prepare (Conditional.hpp:45-70): props from arg2, ranges = common of the three
                Node &c0 = ch(0), &c1 = ch(1), &c2 = ch(2);
                if (c1.scalar || c2.scalar || c0.scalar) {
                    // scalar operands carry no ranges
                    const Node* firstf = !c1.scalar ? &c1 : (!c2.scalar ? &c2 : (!c0.scalar ? &c0 : nullptr));
                    if (!firstf) {
                        n.scalar = true;
                        break;
                    }
                    inherit(n, *firstf);
                    for (Node* o : {&c0, &c1, &c2})
                        if (!o->scalar) {
                            n.acc = common(n.acc, o->acc);
                            n.local = common(n.local, o->local);
                        }
                    break;
                }
                for (int d = 0; d < D3; ++d)
                    if (c0.loc[d] != c1.loc[d] || c0.loc[d] != c2.loc[d]) return fail(OPF_ERR_LOC, "conditional(): operands' loc differ");
                inherit(n, c1);
                n.acc = common(common(c0.acc, c1.acc), c2.acc);
                n.local = common(common(c0.local, c1.local), c2.local);
                break;
            }
            case K_D2C:// D2SecondOrderCentered::prepare (D2SecondOrderCentered.hpp:187-204)
                if (ch(0).scalar) return fail(OPF_ERR_INVALID, "stencil operator applied to a scalar");
                inherit(n, ch(0));
                n.acc.start[n.axis]++, n.acc.end[n.axis]--;
                n.logical.start[n.axis]++, n.logical.end[n.axis]--;
                n.local.start[n.axis]++, n.local.end[n.axis]--;
                break;
            case K_D1C:// D1FirstOrderCentered::prepare (D1FirstOrderCentered.hpp:38-63)
                if (ch(0).scalar) return fail(OPF_ERR_INVALID, "stencil operator applied to a scalar");
                inherit(n, ch(0));
                if (ch(0).loc[n.axis] == OPF_LOC_CENTER) {
                    n.loc[n.axis] = OPF_LOC_CORNER;
                    n.acc.start[n.axis]++, n.local.start[n.axis]++, n.logical.start[n.axis]++;
                } else {
                    n.loc[n.axis] = OPF_LOC_CENTER;
                    n.acc.end[n.axis]--, n.local.end[n.axis]--, n.logical.end[n.axis]--;
                }
                break;
            case K_D1DN:// D1FirstOrderBiasedDownwind::prepare (D1FirstOrderBiasedDownwind.hpp:68-84)
                if (ch(0).scalar) return fail(OPF_ERR_INVALID, "stencil operator applied to a scalar");
                inherit(n, ch(0));
                n.acc.start[n.axis]++, n.logical.start[n.axis]++, n.local.start[n.axis]++;
                break;
            case K_D1UP:// D1FirstOrderBiasedUpwind::prepare (D1FirstOrderBiasedUpwind.hpp:68-84)
                if (ch(0).scalar) return fail(OPF_ERR_INVALID, "stencil operator applied to a scalar");
                inherit(n, ch(0));
                n.acc.end[n.axis]--, n.logical.end[n.axis]--, n.local.end[n.axis]--;
                break;
            case K_WENODN:
            case K_WENOUP:// D1WENO53*::prepare (D1WENO53Downwind.hpp:98-109): logicalRange is NOT shrunk
                if (ch(0).scalar) return fail(OPF_ERR_INVALID, "stencil operator applied to a scalar");
                inherit(n, ch(0));
                n.acc.start[n.axis] += 3, n.acc.end[n.axis] -= 3;
                n.local.start[n.axis] += 3, n.local.end[n.axis] -= 3;
                break;
            case K_INTPC2N:// D1Linear<d,Cen2Cor>::prepare (D1Linear.hpp:50-59)
                if (ch(0).scalar) return fail(OPF_ERR_INVALID, "stencil operator applied to a scalar");
                inherit(n, ch(0));
                n.loc[n.axis] = OPF_LOC_CORNER;
                n.acc.start[n.axis]++, n.local.start[n.axis]++, n.logical.start[n.axis]++;
                break;
            case K_FLC2N:// D1FluxLimiterImpl<K, d, Cen2Cor>::prepare (D1FluxLimiter.hpp:155-171): props from arg2 (the interpolated field)
                if (ch(1).scalar) return fail(OPF_ERR_INVALID, "flux-limiter interpolator applied to a scalar");
                if (ch(1).loc[n.axis] != OPF_LOC_CENTER) return fail(OPF_ERR_LOC, "D1FluxLimiterIntp<Cen2Cor>: operand located in corner in dimension %d", n.axis);
                inherit(n, ch(1));
                n.loc[n.axis] = OPF_LOC_CORNER;
                n.acc.start[n.axis] += 2, n.acc.end[n.axis] -= 1;
                n.local.start[n.axis] += 2, n.local.end[n.axis] -= 1;
                n.logical.start[n.axis] += 2, n.logical.end[n.axis] -= 1;
                break;
            case K_FLN2C:// D1FluxLimiterImpl<K, d, Cor2Cen>::prepare (D1FluxLimiter.hpp:188-203)
                if (ch(1).scalar) return fail(OPF_ERR_INVALID, "flux-limiter interpolator applied to a scalar");
                if (ch(1).loc[n.axis] != OPF_LOC_CORNER) return fail(OPF_ERR_LOC, "D1FluxLimiterIntp<Cor2Cen>: operand located in center in dimension %d", n.axis);
                inherit(n, ch(1));
                n.loc[n.axis] = OPF_LOC_CENTER;
                n.acc.start[n.axis] += 1, n.acc.end[n.axis] -= 2;
                n.local.start[n.axis] += 1, n.local.end[n.axis] -= 2;
                n.logical.start[n.axis] += 1, n.logical.end[n.axis] -= 2;
                break;
            case K_CONV:// Convolution<ns...>::prepare (Convolution.hpp:66-82): every range shrinks by n/2 per axis
                if (ch(0).scalar) return fail(OPF_ERR_INVALID, "convolution applied to a scalar");
                inherit(n, ch(0));
                for (int d = 0; d < D3; ++d) {
                    const int h = n.conv[d] / 2;
                    n.acc.start[d] += h, n.acc.end[d] -= h;
                    n.local.start[d] += h, n.local.end[d] -= h;
                    n.logical.start[d] += h, n.logical.end[d] -= h;
                }
                break;
            case K_INTPN2C:// D1Linear<d,Cor2Cen>::prepare (D1Linear.hpp:60-69)
                if (ch(0).scalar) return fail(OPF_ERR_INVALID, "stencil operator applied to a scalar");
                inherit(n, ch(0));
                n.loc[n.axis] = OPF_LOC_CENTER;
                n.acc.end[n.axis]--, n.logical.end[n.axis]--, n.local.end[n.axis]--;
                break;
            default:
                if (n.nchild == 1) {// UniOp prepare (UniOpDefMacros.hpp.in:16-26)
                    inherit(n, ch(0));
                } else {// BinOp prepare (BinOpDefMacros.hpp.in:19-80)
                    Node &a = ch(0), &b = ch(1);
                    if (a.scalar && b.scalar) n.scalar = true;
                    else if (a.scalar) inherit(n, b);
                    else if (b.scalar) inherit(n, a);
                    else {
                        for (int d = 0; d < D3; ++d)
                            if (a.loc[d] != b.loc[d]) return fail(OPF_ERR_LOC, "binary operator: operands' loc not same (axis %d)", d);
                        inherit(n, a);// logicalRange stays arg1's
                        n.acc = common(a.acc, b.acc);
                        n.local = common(a.local, b.local);
                    }
                }
        }
        return OPF_OK;
    }

    // read footprint of every field leaf relative to the evaluation index (needs prepared loc)
    static void footprint(const Tree& t, int id, const int lo_in[D3], const int hi_in[D3], int lo[][D3], int hi[][D3], bool* used,
                          int mlo[D3], int mhi[D3]) {
        const Node& n = t.nodes[id];
        if (n.kind == K_F) {
            for (int d = 0; d < D3; ++d) {
                lo[n.leaf][d] = used[n.leaf] ? std::min(lo[n.leaf][d], lo_in[d]) : lo_in[d];
                hi[n.leaf][d] = used[n.leaf] ? std::max(hi[n.leaf][d], hi_in[d]) : hi_in[d];
            }
            used[n.leaf] = true;
            return;
        }
        int l[D3], h[D3];
        for (int d = 0; d < D3; ++d) l[d] = lo_in[d], h[d] = hi_in[d];
        if (n.kind == K_FLC2N || n.kind == K_FLN2C) {// the advecting field is read in place, the interpolated one at four taps
            const int a = n.axis, ra = n.kind == K_FLC2N ? -2 : -1;
            mlo[a] = std::min(mlo[a], lo_in[a] - 3);
            mhi[a] = std::max(mhi[a], hi_in[a] + 3);
            footprint(t, n.child[0], lo_in, hi_in, lo, hi, used, mlo, mhi);
            l[a] += ra, h[a] += ra + 3;
            footprint(t, n.child[1], l, h, lo, hi, used, mlo, mhi);
            return;
        }
        if (n.kind == K_CONV)
            for (int d = 0; d < D3; ++d) l[d] -= n.conv[d] / 2, h[d] += n.conv[d] / 2;
        if (n.axis >= 0) {
            const int a = n.axis;
            const bool center = t.nodes[n.child[0]].loc[a] == OPF_LOC_CENTER;
            int ra = 0, rb = 0;// operand taps
            int ma = 0, mb = 0;// mesh-array taps (dx / x indices)
            switch (n.kind) {
                case K_D2C: ra = -1, rb = 1, ma = -1, mb = center ? 1 : 0; break;
                case K_D1C:
                    if (center) ra = -1, rb = 0, ma = -1, mb = 0;
                    else
                        ra = 0, rb = 1, ma = 0, mb = 0;
                    break;
                case K_D1DN: ra = -1, rb = 0, ma = -1, mb = center ? 0 : -1; break;
                case K_D1UP: ra = 0, rb = 1, ma = 0, mb = center ? 1 : 0; break;
                case K_WENODN: ra = -3, rb = 2; break;
                case K_WENOUP: ra = -2, rb = 3; break;
                case K_INTPC2N: ra = -1, rb = 0, ma = -1, mb = 0; break;
                case K_INTPN2C: ra = 0, rb = 1; break;
            }
            mlo[a] = std::min(mlo[a], lo_in[a] + ma - 1);// rdxh/rdxc reach one further
            mhi[a] = std::max(mhi[a], hi_in[a] + mb + 1);
            l[a] += ra;
            h[a] += rb;
        }
        for (int c = 0; c < n.nchild; ++c) footprint(t, n.child[c], l, h, lo, hi, used, mlo, mhi);
    }

    // largest stencil reach of an expression along any axis (sum of the reaches of nested stencil operators)
    static int node_radius(const Tree& t, int id) {
        const Node& n = t.nodes[id];
        int own = 0;
        switch (n.kind) {
            case K_D2C: case K_D1C: case K_D1DN: case K_D1UP: case K_INTPC2N: case K_INTPN2C: own = 1; break;
            case K_WENODN: case K_WENOUP: own = 3; break;
            case K_FLC2N: case K_FLN2C: own = 2; break;
            case K_CONV: own = std::max(n.conv[0], std::max(n.conv[1], n.conv[2])) / 2; break;
            default: own = 0;
        }
        int sub = 0;
        for (int c = 0; c < n.nchild; ++c) sub = std::max(sub, node_radius(t, n.child[c]));
        return own + sub;
    }
    int signature_radius(const char* sig) {
        Tree t;
        std::string k;
        for (const char* p = sig; *p; ++p)
            if (*p != ' ') k.push_back(*p);
        if (parse_signature(k.c_str(), t)) return 1;
        return std::max(1, node_radius(t, 0));
    }


    // offsets (relative to the evaluated cell) at which an expression reads its UNKNOWN leaves (mask bit k <=> field leaf k): the
    // footprint of the matrix row the expression stands for.  Location-dependent stencils (D1C) are taken at their widest.
    static void unknown_taps(const Tree& t, int id, unsigned mask, const std::vector<std::array<int, 3>>& at, std::vector<std::array<int, 3>>& out) {
        const Node& n = t.nodes[id];
        if (n.kind == K_F) {
            if ((mask >> n.leaf) & 1u) out.insert(out.end(), at.begin(), at.end());
            return;
        }
        if (n.kind == K_S || n.kind == K_PAR) return;
        auto shifted = [&](int axis, int lo, int hi) {
            std::vector<std::array<int, 3>> r;
            for (const auto& o : at)
                for (int d = lo; d <= hi; ++d) {
                    auto q = o;
                    q[axis] += d;
                    r.push_back(q);
                }
            std::sort(r.begin(), r.end());
            r.erase(std::unique(r.begin(), r.end()), r.end());
            return r;
        };
        switch (n.kind) {
            case K_D2C: case K_D1C: return unknown_taps(t, n.child[0], mask, shifted(n.axis, -1, 1), out);
            case K_D1DN: case K_INTPC2N: return unknown_taps(t, n.child[0], mask, shifted(n.axis, -1, 0), out);
            case K_D1UP: case K_INTPN2C: return unknown_taps(t, n.child[0], mask, shifted(n.axis, 0, 1), out);
            case K_WENODN: return unknown_taps(t, n.child[0], mask, shifted(n.axis, -3, 2), out);
            case K_WENOUP: return unknown_taps(t, n.child[0], mask, shifted(n.axis, -2, 3), out);
            case K_FLC2N:
                unknown_taps(t, n.child[0], mask, at, out);
                return unknown_taps(t, n.child[1], mask, shifted(n.axis, -2, 1), out);
            case K_FLN2C:
                unknown_taps(t, n.child[0], mask, at, out);
                return unknown_taps(t, n.child[1], mask, shifted(n.axis, -1, 2), out);
            case K_CONV: {
                std::vector<std::array<int, 3>> r = at;
                for (int d = 0; d < 3; ++d) {
                    std::vector<std::array<int, 3>> q;
                    for (const auto& o : r)
                        for (int k = -n.conv[d] / 2; k <= n.conv[d] / 2; ++k) {
                            auto v = o;
                            v[d] += k;
                            q.push_back(v);
                        }
                    r.swap(q);
                }
                return unknown_taps(t, n.child[0], mask, r, out);
            }
            default:
                for (int c = 0; c < n.nchild; ++c) unknown_taps(t, n.child[c], mask, at, out);
        }
    }
    int signature_unknown_taps(const char* sig, unsigned mask, std::vector<std::array<int, 3>>& out) {
        Tree t;
        std::string k;
        for (const char* p = sig; *p; ++p)
            if (*p != ' ') k.push_back(*p);
        if (parse_signature(k.c_str(), t)) return 1;
        out.clear();
        unknown_taps(t, 0, mask, {{0, 0, 0}}, out);
        std::sort(out.begin(), out.end());
        out.erase(std::unique(out.begin(), out.end()), out.end());
        return 0;
    }

    // ------------------------------------------------------------------------------------------ registry
    struct Registry {
        std::mutex mu;
        std::unordered_map<std::string, opf_expr_launcher> map;
        std::vector<std::string> builtin;
    };
    static Registry& registry() {
        static Registry r;
        return r;
    }
    static std::string strip(const char* s) {
        std::string o;
        for (; *s; ++s)
            if (*s != ' ') o.push_back(*s);
        return o;
    }
    void register_builtin(const char* sig, opf_expr_launcher fn) {
        Registry& r = registry();
        std::lock_guard<std::mutex> g(r.mu);
        std::string k = strip(sig);
        if (!r.map.count(k)) r.builtin.push_back(k);
        r.map[k] = fn;
    }
    static opf_expr_launcher find_launcher(const std::string& k) {
        Registry& r = registry();
        std::lock_guard<std::mutex> g(r.mu);
        auto it = r.map.find(k);
        return it == r.map.end() ? nullptr : it->second;
    }

    // ------------------------------------------------------------------------------------------ plan
    struct Plan {
        Tree tree;
        opf_expr_launcher fn = nullptr;
        std::string sig;
    };
    static std::unordered_map<std::string, std::unique_ptr<Plan>>& plan_cache() {
        static std::unordered_map<std::string, std::unique_ptr<Plan>> c;
        return c;
    }
    static int get_plan(const char* signature, Plan** out, bool need_launcher) {
        std::string k = strip(signature);
        auto& c = plan_cache();
        auto it = c.find(k);
        if (it == c.end()) {
            auto p = std::make_unique<Plan>();
            p->sig = k;
            if (int rc = parse_signature(k.c_str(), p->tree)) return rc;
            it = c.emplace(k, std::move(p)).first;
        }
        Plan* p = it->second.get();
        if (need_launcher && !p->fn) {
            p->fn = find_launcher(k);
            if (!p->fn)
                return fail(OPF_ERR_UNSUPPORTED,
                            "expression '%s' has no compiled device kernel: compile the program with nvcc against <OpFlow> "
                            "(registers it) or add it to opflow_b200/csrc/builtin_exprs.cu",
                            k.c_str());
        }
        *out = p;
        return OPF_OK;
    }

    static int fill_args(Plan& p, const opf_field_t* fields, int nfields, const double* scalars, int nscalars, const Range& r,
                         const opf_mesh_s* mesh_hint, opf::ExprArgs& a, int& alias0) {
        Tree& t = p.tree;
        if (nfields < t.nfields) return fail(OPF_ERR_INVALID, "expression '%s' needs %d fields, got %d", p.sig.c_str(), t.nfields, nfields);
        if (nscalars < t.nscalars) return fail(OPF_ERR_INVALID, "expression '%s' needs %d scalars, got %d", p.sig.c_str(), t.nscalars, nscalars);
        if (int rc = prepare_node(t, 0, fields, nfields)) return rc;
        memset(&a, 0, sizeof a);
        const opf_mesh_s* mesh = mesh_hint;
        alias0 = t.nfields > 1;
        for (int k = 0; k < t.nfields; ++k) {
            const opf_field_s* f = fields[k];
            if (!f) return fail(OPF_ERR_INVALID, "field argument %d is null", k);
            if (!f->buf[0]) return fail(OPF_ERR_INVALID, "field '%s' is a plan (opf_field_plan): it has no device storage", f->name.c_str());
            a.f[k].p = f->biased(f->cur);
            a.f[k].s1 = f->pitch1;
            a.f[k].s2 = f->pitch2;
            if (f != fields[0]) alias0 = 0;
            if (!mesh) mesh = f->mesh;
            // operands must live on the same mesh (BinOpDefMacros.hpp.in:21-22 asserts mesh equality)
            if (mesh != f->mesh) {
                for (int d = 0; d < mesh->dim; ++d)
                    if (mesh->ax[d].x != f->mesh->ax[d].x || mesh->ext_range.start[d] != f->mesh->ext_range.start[d])
                        return fail(OPF_ERR_INVALID, "operands of '%s' live on different meshes", p.sig.c_str());
            }
        }
        for (int k = 0; k < t.nscalars; ++k) a.s[k] = scalars[k];
        if (mesh)
            for (int d = 0; d < mesh->dim; ++d) a.ax[d] = mesh_axis_view(mesh, d);
        for (size_t i = 0; i < t.nodes.size(); ++i) {
            const Node& n = t.nodes[i];
            unsigned char bits = 0;
            if (n.axis >= 0) {
                const Node& c = t.nodes[n.child[0]];
                for (int d = 0; d < D3; ++d)
                    if (c.loc[d] == OPF_LOC_CENTER) bits |= 1u << d;
            }
            a.loc[i] = bits;
        }
        // bounds: every tap of every leaf over the evaluation box must lie inside that leaf's storage
        if (r.count() > 0 && t.nfields > 0) {
            int lo[OPF_MAX_FIELDS][D3], hi[OPF_MAX_FIELDS][D3];
            bool used[OPF_MAX_FIELDS] = {false};
            int z[D3] = {0, 0, 0};
            int mlo[D3] = {0, 0, 0}, mhi[D3] = {0, 0, 0};
            footprint(t, 0, z, z, lo, hi, used, mlo, mhi);
            for (int k = 0; k < t.nfields; ++k) {
                if (!used[k]) continue;
                const opf_field_s* f = fields[k];
                for (int d = 0; d < f->dim; ++d)
                    if (r.start[d] + lo[k][d] < f->storage.start[d] || r.end[d] + hi[k][d] > f->storage.end[d])
                        return fail(OPF_ERR_RANGE,
                                    "expression '%s' evaluated over [%d,%d) on axis %d reads field '%s' at offsets [%d,%d] outside its storage "
                                    "[%d,%d): increase setExt/setPadding",
                                    p.sig.c_str(), r.start[d], r.end[d], d, f->name.c_str(), lo[k][d], hi[k][d], f->storage.start[d],
                                    f->storage.end[d]);
            }
            if (mesh)
                for (int d = 0; d < mesh->dim; ++d)
                    if (r.start[d] + mlo[d] < mesh->ext_range.start[d] || r.end[d] - 1 + mhi[d] >= mesh->ext_range.end[d])
                        return fail(OPF_ERR_RANGE, "expression '%s' reads mesh spacings outside the mesh's extended range on axis %d", p.sig.c_str(), d);
        }
        return OPF_OK;
    }
}// namespace opfe

namespace opfe {
    // Decomposition of an assignment over a field whose neighbours all lie along ONE axis (slab decomposition):
    // the planes a neighbour needs (within `padding` of the shared face) are swept first so that their exchange runs under
    // the interior sweep.
    struct SlabPlan {
        int axis = -1;
        bool has_lo = false, has_hi = false;
        Range lo, hi, mid;               // sub-boxes of the written range w
        Range clip_lo, clip_hi, clip_mid;// storage sub-boxes for the matching BC fills
    };
    static bool slab_plan(const opf_field_s* f, const Range& w, SlabPlan& sp) {
        if (f->neighbors.empty() || w.count() <= 0 || f->padding <= 0) return false;
        for (const auto& nb : f->neighbors) {
            int partial = -1, npartial = 0;
            for (int d = 0; d < f->dim; ++d)
                if (!(nb.send.start[d] == f->local.start[d] && nb.send.end[d] == f->local.end[d])) {
                    partial = d;
                    ++npartial;
                }
            if (npartial != 1) return false;// whole-block or edge/corner neighbour: not a slab layout
            if (sp.axis >= 0 && sp.axis != partial) return false;
            sp.axis = partial;
            if (nb.send.start[partial] == f->local.start[partial]) sp.has_lo = true;
            else if (nb.send.end[partial] == f->local.end[partial])
                sp.has_hi = true;
            else
                return false;
        }
        const int ax = sp.axis, pad = f->padding;
        if (f->local.end[ax] - f->local.start[ax] < 2 * pad + 1) return false;// slab too thin to have an interior
        sp.lo = sp.hi = sp.mid = w;
        const int lo_end = sp.has_lo ? std::min(w.end[ax], std::max(w.start[ax], f->local.start[ax] + pad)) : w.start[ax];
        const int hi_start = sp.has_hi ? std::max(lo_end, std::min(w.end[ax], f->local.end[ax] - pad)) : w.end[ax];
        sp.lo.end[ax] = lo_end;
        sp.hi.start[ax] = hi_start;
        sp.mid.start[ax] = lo_end;
        sp.mid.end[ax] = hi_start;
        sp.clip_lo = sp.clip_hi = sp.clip_mid = f->storage;
        sp.clip_lo.end[ax] = lo_end;
        sp.clip_hi.start[ax] = hi_start;
        if (sp.has_lo) sp.clip_mid.start[ax] = lo_end;
        if (sp.has_hi) sp.clip_mid.end[ax] = hi_start;
        return true;
    }
}// namespace opfe

namespace opfe {
    // request of opf_assign_host, picked up by opf_assign_ex once the launch description is ready
    // does updatePadding() have ghost cells to refresh (BC extension / periodic copies with a non-empty box)?
    static bool has_ghost_fills(const opf_field_s* f) {
        for (const auto& op : f->fill1)
            if (op.r.count() > 0) return true;
        for (const auto& op : f->fill2)
            if (op.r.count() > 0) return true;
        return false;
    }
    // Can opf_assign_host pipeline this (dst, input) pair in chunks along the slowest axis?  Undecomposed fields without ghost
    // cells to refresh, or -- *dist -- slab decompositions along that axis shared by both fields: the two boundary chunks are
    // uploaded first, the INPUT's halo planes are exchanged once they are on the device, the chunk sweeps / downloads then stream
    // as on one GPU, and the destination's own halo exchange closes the call.
    static bool hostpipe_qualifies(const opf_field_s* dst, const opf_field_s* inf, bool* dist) {
        *dist = false;
        if (dst->dim < 2 || has_ghost_fills(dst) || has_ghost_fills(inf) || !(inf->local == dst->local)) return false;
        const int ax = dst->dim - 1;
        if (dst->local.end[ax] - dst->local.start[ax] < 16) return false;
        if (dst->neighbors.empty() && inf->neighbors.empty()) return true;
        SlabPlan sp;
        const Range w = common(dst->assignable, dst->local);
        if (!comm_active() || !slab_plan(dst, w, sp) || sp.axis != ax) return false;
        if (inf != dst && (inf->n_ranks != dst->n_ranks || inf->padding != dst->padding || inf->neighbors.size() != dst->neighbors.size())) return false;
        if (dst->local.end[ax] - dst->local.start[ax] < 24) return false;// three chunks of eight planes
        *dist = true;
        return true;
    }
    struct HostPipe {
        opf_field_s* in_field;
        const double* host_in;
        double* host_out;
        bool done;
    };
    static thread_local HostPipe* g_hostpipe = nullptr;

    struct PipeStreams {
        cudaStream_t h2d = nullptr, d2h = nullptr;
        cudaEvent_t start = nullptr, up[32] = {}, done[32] = {};
        double *stage_in = nullptr, *stage_out = nullptr;// dense images of one field's localRange
        long long stage_elems = 0;
    };
    static int pipe_streams(PipeStreams** out) {
        static PipeStreams ps;
        if (!ps.h2d) {
            OPF_CUDA(cudaStreamCreateWithFlags(&ps.h2d, cudaStreamNonBlocking));
            OPF_CUDA(cudaStreamCreateWithFlags(&ps.d2h, cudaStreamNonBlocking));
            OPF_CUDA(cudaEventCreateWithFlags(&ps.start, cudaEventDisableTiming));
            for (int i = 0; i < 32; ++i) {
                OPF_CUDA(cudaEventCreateWithFlags(&ps.up[i], cudaEventDisableTiming));
                OPF_CUDA(cudaEventCreateWithFlags(&ps.done[i], cudaEventDisableTiming));
            }
        }
        *out = &ps;
        return OPF_OK;
    }
}// namespace opfe

using namespace opfe;

// CUtensorMap factory for the TMA tile skeleton (called by the launcher templates, also from user translation units).
// Descriptors are cached by (base, dims, strides, box): the ping-pong buffers of a field alternate between two entries.
extern "C" int opf_internal_tensor_map(void* out128, const void* base, const unsigned long long* dims, const unsigned long long* strides_b,
                                       const unsigned* box, int rank) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    static std::map<std::string, CUtensorMap> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return 1;
        encode = (encode_fn) fn;
    }
    std::string key((const char*) &base, sizeof base);
    key.append((const char*) dims, sizeof(unsigned long long) * rank);
    key.append((const char*) strides_b, sizeof(unsigned long long) * (rank - 1));
    key.append((const char*) box, sizeof(unsigned) * rank);
    auto it = cache.find(key);
    if (it == cache.end()) {
        CUtensorMap m;
        cuuint64_t gd[5], gs[5];
        cuuint32_t bx[5], es[5];
        for (int i = 0; i < rank; ++i) {
            gd[i] = dims[i];
            bx[i] = box[i];
            es[i] = 1;
        }
        for (int i = 0; i < rank - 1; ++i) gs[i] = strides_b[i];
        CUresult rc = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t) rank, const_cast<void*>(base), gd, gs, bx, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return 2;
        if (cache.size() > 4096) cache.clear();
        it = cache.emplace(key, m).first;
    }
    memcpy(out128, &it->second, sizeof(CUtensorMap));
    return 0;
}

extern "C" {

int opf_expr_register(const char* signature, opf_expr_launcher fn) {
    if (!signature || !fn) return fail(OPF_ERR_INVALID, "null argument");
    Registry& r = registry();
    std::lock_guard<std::mutex> g(r.mu);
    r.map[strip(signature)] = fn;
    return OPF_OK;
}
int opf_expr_register_abi(const char* signature, opf_expr_launcher fn, unsigned long long abi) {
    if (abi != OPF_DEVICE_ABI)
        return fail(OPF_ERR_INVALID, "expression '%s' was compiled against another opf_device.cuh (layout stamp %llx, library %llx): rebuild the program",
                    signature ? signature : "", abi, (unsigned long long) OPF_DEVICE_ABI);
    return opf_expr_register(signature, fn);
}
int opf_expr_is_registered(const char* signature) { return signature && find_launcher(strip(signature)) != nullptr; }
int opf_expr_builtin_count(void) { return (int) registry().builtin.size(); }
const char* opf_expr_builtin_name(int i) {
    auto& b = registry().builtin;
    return i >= 0 && i < (int) b.size() ? b[i].c_str() : nullptr;
}

int opf_expr_prepare(const char* signature, const opf_field_t* fields, int nfields, int which, opf_range* out, int* loc) {
    Plan* p;
    if (int rc = get_plan(signature, &p, false)) return rc;
    if (nfields < p->tree.nfields) return fail(OPF_ERR_INVALID, "need %d fields", p->tree.nfields);
    if (int rc = prepare_node(p->tree, 0, fields, nfields)) return rc;
    const Node& n = p->tree.nodes[0];
    if (out) {
        switch (which) {
            case 0: *out = to_c(n.local); break;
            case 2: *out = to_c(n.acc); break;
            case 3: *out = to_c(n.logical); break;
            case 1: {
                Range e;
                e.set_empty(n.src ? n.src->dim : D3);
                *out = to_c(e);
                break;
            }
            default: return fail(OPF_ERR_INVALID, "bad selector");
        }
    }
    if (loc)
        for (int d = 0; d < D3; ++d) loc[d] = n.loc[d];
    return OPF_OK;
}

int opf_assign(opf_field_t dst, int op, const char* signature, const opf_field_t* fields, int nfields, const double* scalars,
               int nscalars) {
    return opf_assign_ex(dst, op, signature, fields, nfields, scalars, nscalars, 0);
}

int opf_assign_ex(opf_field_t dst, int op, const char* signature, const opf_field_t* fields, int nfields, const double* scalars,
                  int nscalars, int flags) {
    if (!dst || !signature) return fail(OPF_ERR_INVALID, "null argument");
    if (op < 0 || op > 4) return fail(OPF_ERR_UNSUPPORTED, "assign op %d is integer-only in the reference (Mod/And/Or/Xor/Shift)", op);
    if (int rc = require_device()) return rc;
    if (!dst->buf[0]) return fail(OPF_ERR_INVALID, "field '%s' is a plan (opf_field_plan): it has no device storage", dst->name.c_str());
    Plan* p;
    if (int rc = get_plan(signature, &p, true)) return rc;
    const Range w = common(dst->assignable, dst->local);// FieldAssigner.hpp:49
    opf::ExprArgs a;
    int alias0 = 0;
    if (int rc = fill_args(*p, fields, nfields, scalars, nscalars, w, dst->mesh, a, alias0)) return rc;
    // src.contains(dst) (FieldAssigner.hpp:28, CartesianField.hpp:793)
    bool alias = false;
    for (int k = 0; k < p->tree.nfields; ++k)
        if (fields[k] == dst) alias = true;
    // a pure point-wise expression of dst itself may be updated in place (no neighbour taps)
    bool has_stencil = false;
    for (const auto& n : p->tree.nodes)
        if (n.axis >= 0 || n.kind == K_CONV) has_stencil = true;
    const bool use_twin = alias && has_stencil;
    if (use_twin)
        if (int rc = field_ensure_twin(dst)) return rc;
    opf::LaunchInfo li;
    memset(&li, 0, sizeof li);
    const int wr = use_twin ? 1 - dst->cur : dst->cur;
    li.dst.p = dst->biased(wr);
    li.dst.s1 = dst->pitch1;
    li.dst.s2 = dst->pitch2;
    li.old = dst->biased(dst->cur);
    for (int d = 0; d < 3; ++d) {
        li.r.lo[d] = w.start[d];
        li.r.hi[d] = w.end[d];
    }
    li.dim = dst->dim;
    li.op = op;
    li.mode = ctx().mode;
    li.alias0 = alias0;
    li.rop = -1;
    // register-window skeleton: vector loads need 16-byte aligned rows at the first evaluated cell
    li.window = opf_internal_opt(OPF_OPT_WINDOW) && dst->dim >= 2 && (w.end[0] - w.start[0]) >= 8;
    li.valign = 0;
    for (int k = 0; k < p->tree.nfields; ++k) {
        const opf_field_s* f = fields[k];
        const bool al = ((reinterpret_cast<uintptr_t>(f->biased(f->cur) + w.start[0]) & 15) == 0) && (f->pitch1 % 2 == 0) && (f->pitch2 % 2 == 0);
        if (al) li.valign |= 1u << k;
    }
    li.dalign = ((reinterpret_cast<uintptr_t>(li.dst.p + w.start[0]) & 15) == 0) && (dst->pitch1 % 2 == 0) && (dst->pitch2 % 2 == 0);
    li.uniform = 1;
    for (int d = 0; d < dst->dim; ++d)
        if (!a.ax[d].uniform) li.uniform = 0;
    li.tma_ok = dst->dim == 3;
    for (int k = 0; k < p->tree.nfields && li.tma_ok; ++k) {
        const opf_field_s* f = fields[alias0 ? 0 : k];
        auto& t = li.tma[k];
        t.base = f->buf[f->cur];
        t.dim[0] = (unsigned long long) f->pitch1;
        t.dim[1] = (unsigned long long) (f->storage.end[1] - f->storage.start[1]);
        t.dim[2] = (unsigned long long) (f->storage.end[2] - f->storage.start[2]);
        t.stride_b[0] = (unsigned long long) f->pitch1 * 8ull;
        t.stride_b[1] = (unsigned long long) f->pitch2 * 8ull;
        t.org[0] = f->storage.start[0] - (int) f->lead;
        t.org[1] = f->storage.start[1];
        t.org[2] = f->storage.start[2];
        if ((reinterpret_cast<uintptr_t>(t.base) & 15) || (t.stride_b[0] & 15) || (t.stride_b[1] & 15) || f->dim != 3) li.tma_ok = 0;
    }
    auto launch_box = [&](const Range& box, cudaStream_t lst = nullptr) -> int {
        if (box.count() <= 0) return OPF_OK;
        if (!lst) lst = ctx().stream;
        opf::LaunchInfo lb = li;
        for (int d = 0; d < 3; ++d) {
            lb.r.lo[d] = box.start[d];
            lb.r.hi[d] = box.end[d];
        }
        // the 16-byte alignment flags were derived for rows starting at w.start[0]: a sub-box shifted by an odd number of cells
        // along axis 0 (decomposition along x) must fall back to scalar loads / stores
        if ((box.start[0] - w.start[0]) & 1) {
            lb.valign = 0;
            lb.dalign = 0;
        }
        const int rc = p->fn(&a, &lb, lst);
        if (rc != 0) return fail(OPF_ERR_CUDA, "launch of '%s' failed: %s", p->sig.c_str(), rc > 0 ? cudaGetErrorString((cudaError_t) rc) : "expression uses an axis the field does not have, or its TMA descriptor could not be encoded");
        ctx().launches++;
        return OPF_OK;
    };
    // ---- opf_assign_host: upload | sweep | download pipelined in slabs along the slowest axis (three streams; PCIe runs in both
    // directions at once and the sweep hides under the copies).  Fields with ghost cells to refresh after the upload, decomposed
    // fields and 1-D fields take the sequential route in opf_assign_host.
    if (g_hostpipe && !g_hostpipe->done) {
        HostPipe& hp = *g_hostpipe;
        const int ax = dst->dim - 1;
        opf_field_s* inf = hp.in_field;
        const int nz = dst->local.end[ax] - dst->local.start[ax];
        bool dist = false;
        const bool simple = hostpipe_qualifies(dst, inf, &dist) && !(flags & OPF_ASSIGN_NO_PADDING);
        if (simple) {
            Context& c = ctx();
            PipeStreams* ps;
            if (int rc = pipe_streams(&ps)) return rc;
            static const int max_chunks = getenv("OPF_PIPE_CHUNKS") ? std::max(2, std::min(32, atoi(getenv("OPF_PIPE_CHUNKS")))) : 32;
            const int nch = std::max(dist ? 3 : 1, std::min(max_chunks, nz / 8));
            int radius = 0;// reach of the expression along the pipelined axis (planes of the next chunk a sweep needs)
            {
                int lo[OPF_MAX_FIELDS][D3], hi[OPF_MAX_FIELDS][D3];
                bool used[OPF_MAX_FIELDS] = {false};
                int z[D3] = {0, 0, 0}, mlo[D3] = {0, 0, 0}, mhi[D3] = {0, 0, 0};
                footprint(p->tree, 0, z, z, lo, hi, used, mlo, mhi);
                for (int k = 0; k < p->tree.nfields; ++k)
                    if (used[k]) radius = std::max(radius, std::max(-lo[k][ax], hi[k][ax]));
            }
            auto zcut = [&](int cidx) { return dst->local.start[ax] + (int) ((long long) nz * cidx / nch); };
            if (radius > zcut(1) - zcut(0)) return fail(OPF_ERR_UNSUPPORTED, "opf_assign_host: stencil reach exceeds the pipeline chunk");
            OPF_CUDA(cudaEventRecord(ps->start, c.stream));
            OPF_CUDA(cudaStreamWaitEvent(ps->h2d, ps->start, 0));
            OPF_CUDA(cudaStreamWaitEvent(ps->d2h, ps->start, 0));
            inf->bc0_clean[inf->cur] = false;
            const long long total = dst->local.count();
            if (total > ps->stage_elems) {
                if (ps->stage_in) cudaFree(ps->stage_in);
                if (ps->stage_out) cudaFree(ps->stage_out);
                ps->stage_in = ps->stage_out = nullptr;
                ps->stage_elems = 0;
                OPF_CUDA(cudaMalloc(&ps->stage_in, sizeof(double) * total));
                OPF_CUDA(cudaMalloc(&ps->stage_out, sizeof(double) * total));
                ps->stage_elems = total;
            }
            // dense layout of localRange: rows of e0, planes of e0*e1 (1 in 2-D); a slab [z0, z1) of the slowest axis is contiguous
            const long long e0 = dst->local.end[0] - dst->local.start[0], e1 = dst->dim == 3 ? dst->local.end[1] - dst->local.start[1] : 1;
            const long long slab = e0 * e1;// doubles per index of the pipelined axis
            auto off = [&](int zc) { return (long long) (zc - dst->local.start[ax]) * slab; };
            auto convert = [&](opf_field_s* f, int which, double* dense, int z0, int z1, bool unpack) -> int {
                Range r = f->local;
                r.start[ax] = z0, r.end[ax] = z1;
                // dense strides follow the field's axes: axis 1 stride e0; axis 2 stride e0*e1 (unused in 2-D where axis 1 is pipelined)
                return dense_convert(f, which, dense + off(z0), r, e0, e0 * e1, unpack, c.stream);
            };
            // updatePadding() of the uploaded input, chunk by chunk: its Corner-Dirichlet boundary nodes take the BC value (step 0 of
            // CartesianField.hpp:351-364; ghost-cell extensions do not exist on fields that qualify for this route)
            auto fill_input = [&](int buf, int z0, int z1) -> int {
                if (inf->fill0.empty()) return OPF_OK;
                Range clip = inf->storage;
                if (z0 > inf->local.start[ax]) clip.start[ax] = z0;
                if (z1 < inf->local.end[ax]) clip.end[ax] = z1;
                const int saved = inf->cur;
                inf->cur = buf;// the fill addresses the buffer being uploaded even after dst (== inf) has flipped to its twin
                inf->bc0_clean[buf] = false;
                const int rc = field_fill_bc(inf, &clip, c.stream);
                inf->cur = saved;
                return rc;
            };
            static const bool dbg = getenv("OPF_PIPE_DEBUG") != nullptr;
            static cudaEvent_t te[6] = {};
            if (dbg && !te[0])
                for (auto& ev : te) cudaEventCreate(&ev);
            if (dbg) cudaEventRecord(te[0], ps->h2d);
            for (int k = 0; k < nch; ++k) {
                // decomposed: the two boundary chunks first (their planes feed the input's halo exchange)
                const int ci = !dist ? k : (k == 0 ? 0 : (k == 1 ? nch - 1 : k - 1));
                OPF_CUDA(cudaMemcpyAsync(ps->stage_in + off(zcut(ci)), hp.host_in + off(zcut(ci)), sizeof(double) * (size_t) (off(zcut(ci + 1)) - off(zcut(ci))),
                                         cudaMemcpyHostToDevice, ps->h2d));
                OPF_CUDA(cudaEventRecord(ps->up[ci], ps->h2d));
            }
            if (dbg) cudaEventRecord(te[1], ps->h2d);
            if (dbg) cudaEventRecord(te[2], c.stream);
            const int in_buf = inf->cur;// the buffer the sweep reads (dst's `cur` flips below when dst is in_field and ping-pongs)
            OPF_CUDA(cudaStreamWaitEvent(c.stream, ps->up[0], 0));
            if (int rc = convert(inf, in_buf, ps->stage_in, zcut(0), zcut(1), true)) return rc;
            if (int rc = fill_input(in_buf, zcut(0), zcut(1))) return rc;
            if (dist) {// input halo: both boundary chunks are on the device -> exchange the planes the neighbours' sweeps tap
                OPF_CUDA(cudaStreamWaitEvent(c.stream, ps->up[nch - 1], 0));
                if (int rc = convert(inf, in_buf, ps->stage_in, zcut(nch - 1), zcut(nch), true)) return rc;
                if (int rc = fill_input(in_buf, zcut(nch - 1), zcut(nch))) return rc;
                if (int rc = halo_exchange(inf, c.stream)) return rc;
            }
            for (int ci = 0; ci < nch; ++ci) {
                if (ci + 1 < nch && !(dist && ci + 1 == nch - 1)) {// the sweep of slab ci taps the first planes of slab ci+1
                    OPF_CUDA(cudaStreamWaitEvent(c.stream, ps->up[ci + 1], 0));
                    if (int rc = convert(inf, in_buf, ps->stage_in, zcut(ci + 1), zcut(ci + 2), true)) return rc;
                    if (int rc = fill_input(in_buf, zcut(ci + 1), zcut(ci + 2))) return rc;
                }
                Range box = w, clip = dst->storage;
                box.start[ax] = std::max(w.start[ax], zcut(ci));
                box.end[ax] = std::min(w.end[ax], zcut(ci + 1));
                if (ci > 0) clip.start[ax] = zcut(ci);
                if (ci < nch - 1) clip.end[ax] = zcut(ci + 1);
                if (int rc = launch_box(box)) return rc;
                if (ci == 0 && use_twin) dst->cur = wr;// fills and downloads address the buffer being written
                if (int rc = field_fill_bc(dst, &clip)) return rc;
                if (int rc = convert(dst, dst->cur, ps->stage_out, zcut(ci), zcut(ci + 1), false)) return rc;
                OPF_CUDA(cudaEventRecord(ps->done[ci], c.stream));
                OPF_CUDA(cudaStreamWaitEvent(ps->d2h, ps->done[ci], 0));
                OPF_CUDA(cudaMemcpyAsync(hp.host_out + off(zcut(ci)), ps->stage_out + off(zcut(ci)), sizeof(double) * (size_t) (off(zcut(ci + 1)) - off(zcut(ci))),
                                         cudaMemcpyDeviceToHost, ps->d2h));
            }
            dst->bc0_clean[dst->cur] = true;
            if (dist) {// updatePadding() of the result: halo planes of the new values (downloads of the last chunks still in flight)
                if (int rc = halo_exchange(dst, c.stream)) return rc;
                if (int rc = field_fill_periodic(dst)) return rc;
            }
            if (dbg) cudaEventRecord(te[3], c.stream);
            if (dbg) cudaEventRecord(te[4], ps->d2h);
            OPF_CUDA(cudaStreamSynchronize(ps->d2h));
            OPF_CUDA(cudaStreamSynchronize(c.stream));
            if (dbg) {
                float a = 0, b = 0, d = 0, e = 0;
                cudaEventElapsedTime(&a, te[0], te[1]);
                cudaEventElapsedTime(&b, te[2], te[3]);
                cudaEventElapsedTime(&d, te[0], te[4]);
                cudaEventElapsedTime(&e, te[0], te[3]);
                fprintf(stderr, "[opf pipe] chunks=%d h2d %.2f ms | compute stream span %.2f ms | start->last d2h %.2f ms | start->compute end %.2f ms\n", nch, a, b, d, e);
            }
            hp.done = true;
            return OPF_OK;
        }
    }
    // ---- slab-decomposed destination: halo exchange overlapped with the interior sweep (replaces the serial
    // pack -> MPI_Isend/Irecv -> Waitall -> unpack of CartesianField.hpp:630-768).  Order:
    //   compute stream: interior sweep, BC fill of its planes ............................................ | wait(comm)
    //   comm stream   : boundary-slab sweeps, BC fill of those planes, pack -> NCCL send/recv -> unpack
    SlabPlan sp;
    if (opf_internal_opt(OPF_OPT_OVERLAP) && !(flags & OPF_ASSIGN_NO_PADDING) && comm_active() && slab_plan(dst, w, sp)) {
        Context& c = ctx();
        // OPF_HALO_DEBUG=1: event time line of the overlapped step (no profiler needed), printed every 50th call
        static const bool hdbg = getenv("OPF_HALO_DEBUG") != nullptr;
        static cudaEvent_t he[8] = {};
        static long long hcalls = 0;
        static double hacc[8] = {};
        if (hdbg && !he[0])
            for (auto& e : he) cudaEventCreate(&e);
        auto mark = [&](int i, cudaStream_t st) {
            if (hdbg) cudaEventRecord(he[i], st);
        };
        // the boundary slabs, their BC fills and the whole exchange run on the high-priority stream, concurrently with the
        // interior sweep on the compute stream (disjoint planes of the destination; both read the old buffer)
        OPF_CUDA(cudaEventRecord(c.ev_compute, c.stream));
        mark(0, c.stream);
        OPF_CUDA(cudaStreamWaitEvent(c.comm_stream, c.ev_compute, 0));
        mark(1, c.comm_stream);
        if (int rc = launch_box(sp.lo, c.comm_stream)) return rc;
        if (int rc = launch_box(sp.hi, c.comm_stream)) return rc;
        mark(2, c.comm_stream);
        if (int rc = launch_box(sp.mid)) return rc;
        mark(5, c.stream);
        if (use_twin) dst->cur = wr;
        if (sp.has_lo)
            if (int rc = field_fill_bc(dst, &sp.clip_lo, c.comm_stream)) return rc;
        if (sp.has_hi)
            if (int rc = field_fill_bc(dst, &sp.clip_hi, c.comm_stream)) return rc;
        mark(3, c.comm_stream);
        if (int rc = halo_exchange(dst, c.comm_stream)) return rc;
        mark(4, c.comm_stream);
        OPF_CUDA(cudaEventRecord(c.ev_comm, c.comm_stream));
        if (int rc = field_fill_bc(dst, &sp.clip_mid)) return rc;
        dst->bc0_clean[dst->cur] = true;
        OPF_CUDA(cudaStreamWaitEvent(c.stream, c.ev_comm, 0));
        mark(6, c.stream);
        const int prc = field_fill_periodic(dst);// unsplit periodic axes: local copies over the whole logical box, exchanged planes included
        if (hdbg) {
            cudaEventSynchronize(he[6]);
            for (int i = 1; i <= 6; ++i) {
                float ms = 0;
                cudaEventElapsedTime(&ms, he[0], he[i]);
                hacc[i] += ms;
            }
            if (++hcalls % 50 == 0) {
                fprintf(stderr, "[opf halo r%d] us from step start: comm-start %.1f | boundary sweeps done %.1f | fills %.1f | exchange done %.1f || interior sweep done %.1f | joined %.1f\n",
                        dst->rank, 1e3 * hacc[1] / 50, 1e3 * hacc[2] / 50, 1e3 * hacc[3] / 50, 1e3 * hacc[4] / 50, 1e3 * hacc[5] / 50, 1e3 * hacc[6] / 50);
                for (auto& a2 : hacc) a2 = 0;
            }
        }
        return prc;
    }
    if (int rc = launch_box(w)) return rc;
    if (use_twin) dst->cur = wr;// ping-pong instead of the reference's temp copy + second sweep
    if (flags & OPF_ASSIGN_NO_PADDING) return OPF_OK;
    return field_update_padding(dst);// CartesianField.hpp:231
}

int opf_assign_host(opf_field_t dst, int op, const char* signature, const opf_field_t* fields, int nfields, const double* scalars,
                    int nscalars, opf_field_t in_field, const double* host_in, double* host_out) {
    if (!dst || !in_field || !host_in || !host_out) return fail(OPF_ERR_INVALID, "null argument");
    HostPipe hp{in_field, host_in, host_out, false};
    g_hostpipe = &hp;
    int rc = OPF_OK;
    const opf_range lr_in = to_c(in_field->local), lr_out = to_c(dst->local);
    // the pipelined route consumes the request inside opf_assign_ex; otherwise: plain upload -> assign -> download
    bool dist_unused = false;
    const bool try_pipe = hostpipe_qualifies(dst, in_field, &dist_unused);
    auto sequential = [&]() -> int {
        // upload -> updatePadding of the input (BC ghosts, halo planes: the uploaded values ARE the field now) -> assign -> download
        if (int r2 = opf_field_upload(in_field, &lr_in, host_in)) return r2;
        if (has_ghost_fills(in_field) || !in_field->neighbors.empty())
            if (int r2 = field_update_padding(in_field)) return r2;
        if (int r2 = opf_assign_ex(dst, op, signature, fields, nfields, scalars, nscalars, 0)) return r2;
        return opf_field_download(dst, &lr_out, host_out);
    };
    if (!try_pipe) {
        g_hostpipe = nullptr;
        return sequential();
    }
    rc = opf_assign_ex(dst, op, signature, fields, nfields, scalars, nscalars, 0);
    g_hostpipe = nullptr;
    if (rc) return rc;
    if (hp.done) return OPF_OK;
    // the launch description did not qualify for the pipelined route (e.g. a block decomposition): opf_assign_ex has then run the
    // plain assignment on the field's OLD contents -- not what this call means.  That cannot happen silently:
    return fail(OPF_ERR_INVALID, "opf_assign_host: internal error, the pipelined route was not taken");
}


// `count` consecutive identical assignments (a time loop whose body is one statement, e.g. FTCS-OMP.cpp:24-27) replayed from a
// CUDA graph: small fields are launch-latency bound (C1: a 7 us kernel every 9.7 us), the graph removes the gaps between launches.
// Two plain steps first (allocations, both ping-pong buffers get their boundary nodes), then an even number of steps is captured
// once per (destination, expression, operands, mode) and replayed; what does not fill a whole graph runs as plain launches.
namespace {
    struct RepeatGraph {
        std::string key;
        int steps = 0;
        cudaGraphExec_t exec = nullptr;
        long long launches = 0;
    };
    std::vector<RepeatGraph>& repeat_cache() {
        static std::vector<RepeatGraph> c;
        return c;
    }
}// namespace
extern "C++" {
namespace opfe {
    void repeat_cache_clear() {
        if (ctx().inited) cudaStreamSynchronize(ctx().stream);
        for (auto& e : repeat_cache())
            if (e.exec) cudaGraphExecDestroy(e.exec);
        repeat_cache().clear();
    }
}// namespace opfe
}
int opf_assign_repeat(opf_field_t dst, int op, const char* signature, const opf_field_t* fields, int nfields, const double* scalars, int nscalars,
                      int count) {
    if (count < 0) return fail(OPF_ERR_INVALID, "negative repeat count");
    if (!dst || !signature) return fail(OPF_ERR_INVALID, "null argument");
    Context& c = ctx();
    auto plain = [&](int n) -> int {
        for (int i = 0; i < n; ++i)
            if (int rc = opf_assign_ex(dst, op, signature, fields, nfields, scalars, nscalars, 0)) return rc;
        return OPF_OK;
    };
    const int unit = 32;// steps per graph
    if (!opf_internal_opt(OPF_OPT_GRAPHS) || count < 2 + unit || !dst->neighbors.empty()) return plain(count);
    std::string key = strip(signature);
    key.append((const char*) &dst, sizeof dst);
    key.append((const char*) &op, sizeof op);
    key.append((const char*) &c.mode, sizeof c.mode);
    key.append((const char*) fields, sizeof(opf_field_t) * nfields);
    key.append((const char*) scalars, sizeof(double) * nscalars);
    for (int k = 0; k < nfields; ++k) key.append((const char*) &fields[k]->cur, sizeof(int));
    RepeatGraph* g = nullptr;
    for (auto& e : repeat_cache())
        if (e.key == key) g = &e;
    int done = 0;
    if (!g) {
        if (int rc = plain(2)) return rc;
        done = 2;
        // the key was taken before the two warm-up steps; ping-pong fields are back in the same state after an even number of steps
        OPF_CUDA(cudaStreamSynchronize(c.stream));
        const long long l0 = c.launches;
        OPF_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
        const int rc = plain(unit);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(c.stream, &graph);
        const long long nl = c.launches - l0;
        c.launches = l0;
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        RepeatGraph ng;
        ng.key = key, ng.steps = unit, ng.launches = nl;
        if (ce == cudaSuccess && graph && cudaGraphInstantiate(&ng.exec, graph, 0) != cudaSuccess) ng.exec = nullptr;
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (repeat_cache().size() > 64) {
            for (auto& e : repeat_cache())
                if (e.exec) cudaGraphExecDestroy(e.exec);
            repeat_cache().clear();
        }
        repeat_cache().push_back(ng);
        g = &repeat_cache().back();
    }
    if (g->exec)
        while (count - done >= g->steps) {
            OPF_CUDA(cudaGraphLaunch(g->exec, c.stream));
            c.launches += g->launches;
            done += g->steps;
        }
    return plain(count - done);
}

int opf_field_assign_field(opf_field_t dst, int op, opf_field_t src) {
    if (!dst || !src) return fail(OPF_ERR_INVALID, "null field");
    if (dst == src && op == OPF_OP_EQ) return OPF_OK;// CartesianField.hpp:188 (this != &other)
    const opf_field_t fs[1] = {src};
    return opf_assign(dst, op, "F<0>", fs, 1, nullptr, 0);
}

// launches the reduction; the result is left in device memory (*dev_result), valid until the next reduction on the stream
static int reduce_launch(int rop, const char* signature, const opf_field_t* fields, int nfields, const double* scalars, int nscalars,
                         const opf_range* range, double** dev_result, bool* empty, double* result_slot = nullptr) {
    if (!signature) return fail(OPF_ERR_INVALID, "null argument");
    if (rop < 0 || rop > 4) return fail(OPF_ERR_INVALID, "bad reduce op");
    if (int rc = require_device()) return rc;
    Plan* p;
    if (int rc = get_plan(signature, &p, true)) return rc;
    if (nfields < 1) return fail(OPF_ERR_INVALID, "reduce needs at least one field");
    // range must be known before the bounds check: prepare first to get the expression's own ranges
    if (int rc = prepare_node(p->tree, 0, fields, nfields)) return rc;
    const Node& root = p->tree.nodes[0];
    Range r = range ? from_c(*range, fields[0]->dim) : common(root.local, root.acc);
    opf::ExprArgs a;
    int alias0 = 0;
    if (int rc = fill_args(*p, fields, nfields, scalars, nscalars, r, fields[0]->mesh, a, alias0)) return rc;
    Context& c = ctx();
    *empty = r.count() <= 0;
    if (*empty) return OPF_OK;
    opf::LaunchInfo li;
    memset(&li, 0, sizeof li);
    for (int d = 0; d < 3; ++d) {
        li.r.lo[d] = r.start[d];
        li.r.hi[d] = r.end[d];
    }
    li.dim = fields[0]->dim;
    li.mode = c.mode;
    li.alias0 = alias0;
    li.rop = rop;
    const long long rows = (long long) (r.end[1] - r.start[1]) * (r.end[2] - r.start[2]);
    const long long items = rows * ((r.end[0] - r.start[0] + opf::RED_SEG - 1) / opf::RED_SEG);
    li.n_partials = (int) std::max<long long>(1, std::min<long long>(items, 4LL * c.sm_count));
    li.partials = c.red_buf;
    li.result = result_slot;
    const int rc = p->fn(&a, &li, c.stream);
    if (rc != 0) return fail(OPF_ERR_CUDA, "reduce launch of '%s' failed", p->sig.c_str());
    c.launches += 2;
    *dev_result = result_slot ? result_slot : c.red_buf + li.n_partials;
    return OPF_OK;
}

int opf_reduce(int rop, const char* signature, const opf_field_t* fields, int nfields, const double* scalars, int nscalars,
               const opf_range* range, double* result) {
    if (!result) return fail(OPF_ERR_INVALID, "null argument");
    double* dev = nullptr;
    bool empty = false;
    if (int rc = reduce_launch(rop, signature, fields, nfields, scalars, nscalars, range, &dev, &empty)) return rc;
    if (empty) {
        *result = rop == OPF_RED_MAX ? -INFINITY : (rop == OPF_RED_MIN ? INFINITY : 0.0);
        return OPF_OK;
    }
    Context& c = ctx();
    OPF_CUDA(cudaMemcpyAsync(c.red_host, dev, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    OPF_CUDA(cudaStreamSynchronize(c.stream));
    *result = c.red_host[0];
    return OPF_OK;
}

}// extern "C"

namespace opfe {
    // device-resident sum of a field over a box (no host synchronisation): used by the multigrid mean projection
    // device-resident dot product over a box: the folded value is written to `slot` (device memory), stream-ordered, no host
    // synchronisation; an empty box writes nothing (callers pre-zero the slot)
    int dot_device(opf_field_s* a, opf_field_s* b, const Range& r, double* slot) {
        opf_field_t F[2] = {a, b};
        opf_range cr = to_c(r);
        bool empty = false;
        double* dev = nullptr;
        return reduce_launch(OPF_RED_SUM, "Mul<F<0>,F<1>>", F, 2, nullptr, 0, &cr, &dev, &empty, slot);
    }
    int reduce_sum_device(opf_field_s* f, const Range& r, double** dev_result) {
        opf_field_t F[1] = {f};
        opf_range cr = to_c(r);
        bool empty = false;
        return reduce_launch(OPF_RED_SUM, "F<0>", F, 1, nullptr, 0, &cr, dev_result, &empty);
    }
}// namespace opfe
