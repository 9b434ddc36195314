/* oracle/opflow_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked, imported or executed by the product path).
 *
 * A scalar, plain-C restatement of the reference's algorithm for the stencil hot path (OpFlow v0.2.7), used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker.  Every function cites the reference
 * file:line it follows (paths relative to the reference tree).  PARITY IS PINNED: tests/test_oracle_pinned.py checks
 * this file against (a) the golden literals of the reference's own tests (CartesianFieldTest, DircBCTest, NeumBCTest,
 * PeriodicBCTest, EvenSplitStrategyTest, CSRMatrixGeneratorTest) and (b) fixtures under tests/golden/ produced by the
 * unmodified reference built by oracle/build_ref.sh (generator: oracle/make_golden.py).
 *
 * Arithmetic: IEEE double, one rounding per operation, the reference's operation order; compile with
 * -O2 -ffp-contract=off (no FMA contraction) so results are bit-comparable with the reference's g++ build.
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXD 3
#define ORC_MAXN 128

enum { LOC_CORNER = 0, LOC_CENTER = 1 };
enum { BC_UNDEFINED = 0, BC_DIRC = 1, BC_NEUM = 2, BC_PERIODIC = 3, BC_INTERNAL = 4, BC_SYMM = 5, BC_ASYMM = 6 };

typedef struct {
    int start[ORC_MAXD], end[ORC_MAXD];
} orc_range;

/* dense field storage: element g at data[(g0-s0) + n0*((g1-s1) + n1*(g2-s2))], s = storage.start
 * (PlainTensor::getOffset, src/DataStructures/Arrays/Tensor/PlainTensor.hpp:204-212) */
typedef struct {
    int dim;
    int loc[ORC_MAXD];
    orc_range storage, local, assignable, accessible, logical;
    int bc_type[ORC_MAXD][2];
    double bc_value[ORC_MAXD][2];
    double* data;
} orc_field;

typedef struct {
    int dim;
    int ext_start[ORC_MAXD], n_ext[ORC_MAXD];
    const double* x[ORC_MAXD];  /* n_ext   */
    const double* dx[ORC_MAXD]; /* n_ext-1 */
} orc_mesh;

static double mesh_x(const orc_mesh* m, int d, int i) { return m->x[d][i - m->ext_start[d]]; }
static double mesh_dx(const orc_mesh* m, int d, int i) { return m->dx[d][i - m->ext_start[d]]; }

static long long f_off(const orc_field* f, const int* g) {
    long long n0 = f->storage.end[0] - f->storage.start[0], n1 = f->storage.end[1] - f->storage.start[1];
    return (g[0] - f->storage.start[0]) + n0 * ((long long) (g[1] - f->storage.start[1]) + n1 * (long long) (g[2] - f->storage.start[2]));
}
static double f_get(const orc_field* f, const int* g) { return f->data[f_off(f, g)]; }

/* ---------------------------------------------------------------------------------------------- mesh
 * MeshBuilder::set1DRange + set1DMesh(min,max,k) / set1DMesh(f,k) + setExtMesh
 * (src/Core/Mesh/Structured/CartesianMesh.hpp:208-303).  x has n+2*pad entries, dx/idx one less.
 * ext_mode: 0 undefined(=symm) 1 symm 2 periodic 3 uniform.  If xs != NULL the node coordinates are xs[0..n). */
int orc_mesh_axis(int n, int start, int pad, int ext_mode, double xmin, double xmax, const double* xs, double* x, double* dx,
                  double* idx) {
    const int rs = start, re = start + n, es = rs - pad, ee = re + pad;
    int i, j;
    for (i = 0; i < ee - es; ++i) x[i] = 0.0;
    for (i = 0; i < ee - es - 1; ++i) dx[i] = idx[i] = 0.0;
    if (xs) {
        for (i = rs; i < re; ++i) x[i - es] = xs[i - rs];
        for (j = rs; j < re - 1; ++j) {
            dx[j - es] = (x[j + 1 - es] - x[j - es]);
            idx[j - es] = 1. / dx[j - es];
        }
    } else {
        for (i = rs; i < re; ++i) x[i - es] = (xmax - xmin) / (n - 1) * (i - rs) + xmin;
        for (j = rs; j < re - 1; ++j) {
            dx[j - es] = (xmax - xmin) / (n - 1);
            idx[j - es] = 1. / dx[j - es];
        }
    }
    switch (ext_mode) {
        case 0:
        case 1:
            for (i = es; i < rs; ++i) {
                dx[i - es] = dx[2 * rs - 1 - i - es];
                idx[i - es] = 1. / dx[i - es];
            }
            for (i = re - 1; i < ee - 1; ++i) {
                dx[i - es] = dx[2 * re - 3 - i - es];
                idx[i - es] = 1. / dx[i - es];
            }
            break;
        case 2:
            for (i = es; i < rs; ++i) {
                dx[i - es] = dx[re - (rs - i) - es];
                idx[i - es] = 1. / dx[i - es];
            }
            for (i = re - 1; i < ee - 1; ++i) {
                dx[i - es] = dx[rs + i - re + 1 - es];
                idx[i - es] = 1. / dx[i - es];
            }
            break;
        default:
            for (i = es; i < rs; ++i) {
                dx[i - es] = dx[rs - es];
                idx[i - es] = 1. / dx[i - es];
            }
            for (i = re - 1; i < ee - 1; ++i) {
                dx[i - es] = dx[re - 2 - es];
                idx[i - es] = 1. / idx[i - es];
            }
    }
    for (i = rs - 1; i >= es; --i) x[i - es] = x[i + 1 - es] - dx[i - es];
    for (i = re; i < ee; ++i) x[i - es] = x[i - 1 - es] + dx[i - 1 - es];
    return ee - es;
}

/* ---------------------------------------------------------------------------------------------- ranges
 * ExprBuilder::calculateRanges + validateRanges without a split strategy
 * (src/Core/Field/MeshBased/Structured/CartesianField.hpp:941-1029).  Fills f->local/assignable/accessible/logical and
 * f->storage = local inflated by padding (:933); returns the effective padding. */
static orc_range rcommon(orc_range a, orc_range b) {
    orc_range r;
    int d;
    for (d = 0; d < ORC_MAXD; ++d) {
        r.start[d] = a.start[d] > b.start[d] ? a.start[d] : b.start[d];
        r.end[d] = a.end[d] < b.end[d] ? a.end[d] : b.end[d];
    }
    return r;
}

int orc_field_ranges(orc_field* f, const int* mesh_start, const int* mesh_end, const int* ext /* [dim][2] */, int padding,
                     const orc_range* local_override /* strategy->splitRange result (cell ranges) or NULL */) {
    int i, d;
    orc_range mr;
    for (d = 0; d < ORC_MAXD; ++d) {
        mr.start[d] = d < f->dim ? mesh_start[d] : 0;
        mr.end[d] = d < f->dim ? mesh_end[d] : 1;
    }
    for (i = 0; i < f->dim; ++i) {
        if (ext[2 * i] > padding) padding = ext[2 * i];
        if (ext[2 * i + 1] > padding) padding = ext[2 * i + 1];
    }
    f->logical = f->assignable = f->local = f->accessible = mr;
    for (i = 0; i < f->dim; ++i) {
        int loc = f->loc[i];
        int type = f->bc_type[i][0];
        if (type == BC_DIRC && loc == LOC_CORNER) f->assignable.start[i]++;
        type = f->bc_type[i][1];
        if (type == BC_DIRC) {
            if (loc == LOC_CORNER) f->assignable.end[i]--;
            else {
                f->accessible.end[i]--;
                f->assignable.end[i]--;
            }
        } else if (type == BC_NEUM || type == BC_UNDEFINED || type == BC_SYMM || type == BC_ASYMM) {
            if (loc == LOC_CENTER) {
                f->accessible.end[i]--;
                f->assignable.end[i]--;
            }
        } else if (type == BC_PERIODIC) {
            f->accessible.end[i]--;
            f->assignable.end[i]--;
        }
        f->logical.start[i] = f->accessible.start[i] - ext[2 * i];
        f->logical.end[i] = f->accessible.end[i] + ext[2 * i + 1];
    }
    if (local_override) {
        f->local = *local_override;
        for (i = 0; i < f->dim; ++i)
            if (f->loc[i] == LOC_CORNER && f->local.end[i] == mr.end[i] - 1) {
                int e = f->local.end[i] + 1;
                f->local.end[i] = e < f->accessible.end[i] ? e : f->accessible.end[i];
            }
    } else
        f->local = f->accessible;
    f->accessible = rcommon(f->accessible, f->logical);
    f->local = rcommon(f->local, f->logical);
    f->assignable = rcommon(f->assignable, f->accessible);
    f->storage = f->local;
    for (i = 0; i < f->dim; ++i) {
        f->storage.start[i] -= padding;
        f->storage.end[i] += padding;
    }
    return padding;
}

/* ---------------------------------------------------------------------------------------------- ghost fill
 * CartesianField::updatePaddingImpl_final, single rank, constant BCs
 * (src/Core/Field/MeshBased/Structured/CartesianField.hpp:349-629). */
static double intp2(double x1, double y1, double x2, double y2) { return (x1 * y2 - x2 * y1) / (x1 - x2); }
/* Interpolator1D::intp(x1,y1,x2,y2,x) src/Math/Interpolator/Interpolator.hpp:21-23,38-40 */
static double intp(double x1, double y1, double x2, double y2, double x) { return intp2(x1 - x, y1, x2 - x, y2); }

static void for_box(const orc_range* r, const orc_range* clip, void (*fn)(const int*, void*), void* ud) {
    int g[3];
    orc_range b = rcommon(*r, *clip);
    for (g[2] = b.start[2]; g[2] < b.end[2]; ++g[2])
        for (g[1] = b.start[1]; g[1] < b.end[1]; ++g[1])
            for (g[0] = b.start[0]; g[0] < b.end[0]; ++g[0]) fn(g, ud);
}

typedef struct {
    orc_field* f;
    const orc_mesh* m;
    int axis, side, kind;
    double bc;
} fill_ctx;

static void fill_cell(const int* g, void* ud) {
    fill_ctx* c = (fill_ctx*) ud;
    orc_field* f = c->f;
    const orc_mesh* m = c->m;
    const int i = c->axis;
    int mir[3] = {g[0], g[1], g[2]};
    const int ls = f->local.start[i], le = f->local.end[i];
    double v;
    if (c->kind == 0) {/* step 0 */
        f->data[f_off(f, g)] = c->bc;
        return;
    }
    if (c->kind == 100) {/* periodic copy :609-629 */
        mir[i] += c->side == 0 ? (f->accessible.end[i] - f->accessible.start[i]) : -(f->accessible.end[i] - f->accessible.start[i]);
        f->data[f_off(f, g)] = f_get(f, mir);
        return;
    }
    if (f->loc[i] == LOC_CORNER) mir[i] = c->side == 0 ? 2 * ls - g[i] : 2 * le - 2 - g[i];
    else
        mir[i] = c->side == 0 ? 2 * ls - 1 - g[i] : 2 * le - 1 - g[i];
    {
        const double um = f_get(f, mir);
        switch (c->kind) {
            case BC_DIRC:
                if (f->loc[i] == LOC_CORNER)
                    v = intp(mesh_x(m, i, c->side == 0 ? ls : le - 1), c->bc, mesh_x(m, i, mir[i]), um, mesh_x(m, i, g[i]));
                else
                    v = intp(mesh_x(m, i, c->side == 0 ? ls : le), c->bc, mesh_x(m, i, mir[i]) + mesh_dx(m, i, mir[i]) / 2., um,
                             mesh_x(m, i, g[i]) + mesh_dx(m, i, g[i]) / 2.);
                break;
            case BC_NEUM:
                if (f->loc[i] == LOC_CORNER) v = um + c->bc * (mesh_x(m, i, g[i]) - mesh_x(m, i, mir[i]));
                else
                    v = um + c->bc * (mesh_x(m, i, g[i]) + mesh_dx(m, i, g[i]) / 2. - mesh_x(m, i, mir[i]) - mesh_dx(m, i, mir[i]) / 2.);
                break;
            case BC_SYMM: v = um; break;
            default: v = -um; break;
        }
    }
    f->data[f_off(f, g)] = v;
}

void orc_update_padding(orc_field* f, const orc_mesh* m) {
    int i, j, side;
    int start[3], end[3];
    fill_ctx c;
    c.f = f;
    c.m = m;
    /* step 0 (:351-364) */
    for (i = 0; i < f->dim; ++i)
        for (side = 0; side < 2; ++side) {
            int at = side == 0 ? f->local.start[i] == f->accessible.start[i] : f->local.end[i] == f->accessible.end[i];
            if (at && f->bc_type[i][side] == BC_DIRC && f->loc[i] == LOC_CORNER) {
                orc_range r = f->local;
                int pos = side == 0 ? f->local.start[i] : f->local.end[i] - 1;
                r.start[i] = pos;
                r.end[i] = pos + 1;
                c.axis = i, c.side = side, c.kind = 0, c.bc = f->bc_value[i][side];
                for_box(&r, &f->storage, fill_cell, &c);
            }
        }
    /* step 1 (:365-606) */
    for (i = 0; i < f->dim; ++i)
        for (side = 0; side < 2; ++side) {
            int at = side == 0 ? f->local.start[i] == f->accessible.start[i] : f->local.end[i] == f->accessible.end[i];
            int t = f->bc_type[i][side];
            if (at && t != BC_UNDEFINED && t != BC_PERIODIC) {
                orc_range r = f->local;
                if (side == 0) start[i] = f->logical.start[i];
                else
                    end[i] = f->logical.end[i];
                for (j = 0; j < i; ++j) {
                    r.start[j] = start[j];
                    r.end[j] = end[j];
                }
                if (side == 0) {
                    r.start[i] = f->logical.start[i];
                    r.end[i] = f->local.start[i];
                } else {
                    r.start[i] = f->local.end[i];
                    r.end[i] = f->logical.end[i];
                }
                c.axis = i, c.side = side, c.kind = t, c.bc = f->bc_value[i][side];
                for_box(&r, &f->storage, fill_cell, &c);
            } else {
                if (side == 0) start[i] = f->local.start[i];
                else
                    end[i] = f->local.end[i];
            }
        }
    /* step 2, single rank (:609-629) */
    for (i = 0; i < f->dim; ++i)
        if (f->bc_type[i][0] == BC_PERIODIC) {
            orc_range r = f->logical;
            r.end[i] = f->accessible.start[i];
            c.axis = i, c.side = 0, c.kind = 100;
            for_box(&r, &f->storage, fill_cell, &c);
            r = f->logical;
            r.start[i] = f->accessible.end[i];
            c.side = 1;
            for_box(&r, &f->storage, fill_cell, &c);
        }
}

/* ---------------------------------------------------------------------------------------------- expression interpreter
 * The signature grammar is the one of include/opflow_b200.h; evaluation follows each Op::eval of the reference. */
typedef struct {
    char name[20];
    int axis, leaf, nchild, child[3];
    int conv[4]; /* Conv<n0,n1,n2,k0,E> */
    int scalar, loc[3];
    orc_range acc, local, logical;
} onode;
typedef struct {
    onode n[ORC_MAXN];
    int count, nfields, nscalars;
    const char* s;
    int pos, err;
} otree;

static int parse_node(otree* t) {
    int me = t->count++, b, l;
    onode* n;
    if (me >= ORC_MAXN) {
        t->err = 1;
        return -1;
    }
    n = &t->n[me];
    memset(n, 0, sizeof *n);
    n->axis = n->leaf = -1;
    b = t->pos;
    while (isalnum((unsigned char) t->s[t->pos])) t->pos++;
    l = t->pos - b;
    if (l <= 0 || l > 19 || t->s[t->pos] != '<') {
        t->err = 1;
        return -1;
    }
    memcpy(n->name, t->s + b, l);
    t->pos++;
    if (!strcmp(n->name, "F") || !strcmp(n->name, "S")) {
        n->leaf = (int) strtol(t->s + t->pos, NULL, 10);
        while (isdigit((unsigned char) t->s[t->pos])) t->pos++;
        if (n->name[0] == 'F') {
            if (n->leaf + 1 > t->nfields) t->nfields = n->leaf + 1;
        } else {
            n->scalar = 1;
            if (n->leaf + 1 > t->nscalars) t->nscalars = n->leaf + 1;
        }
    } else {
        if (!strcmp(n->name, "Conv")) {
            int q;
            for (q = 0; q < 4; ++q) {
                n->conv[q] = (int) strtol(t->s + t->pos, NULL, 10);
                while (isdigit((unsigned char) t->s[t->pos])) t->pos++;
                t->pos++; /* ',' */
            }
            if (n->conv[3] + n->conv[0] * n->conv[1] * n->conv[2] > t->nscalars) t->nscalars = n->conv[3] + n->conv[0] * n->conv[1] * n->conv[2];
        } else if (isdigit((unsigned char) t->s[t->pos])) {
            n->axis = t->s[t->pos] - '0';
            t->pos += 2; /* digit and ',' */
        }
        for (;;) {
            int c = parse_node(t);
            if (c < 0) return -1;
            t->n[me].child[t->n[me].nchild++] = c;
            if (t->s[t->pos] == ',') {
                t->pos++;
                continue;
            }
            break;
        }
    }
    if (t->s[t->pos] != '>') {
        t->err = 1;
        return -1;
    }
    t->pos++;
    return me;
}



static int is(const onode* n, const char* s) { return !strcmp(n->name, s); }
/* flux-limiter node: Fl<Scheme>C2N / Fl<Scheme>N2C (not "Floor") */
static int is_fl(const onode* n, const char* dir) { return !strncmp(n->name, "Fl", 2) && strlen(n->name) > 5 && !strcmp(n->name + strlen(n->name) - 3, dir); }

static void inherit(onode* n, const onode* a) {
    n->scalar = a->scalar;
    memcpy(n->loc, a->loc, sizeof n->loc);
    n->acc = a->acc;
    n->local = a->local;
    n->logical = a->logical;
}

/* Expr::prepare(): post-order range/loc algebra; cites next to each case */
static int prepare(otree* t, int id, orc_field** fields) {
    onode* n = &t->n[id];
    int c, d = n->axis;
    for (c = 0; c < n->nchild; ++c)
        if (prepare(t, n->child[c], fields)) return 1;
    if (is(n, "F")) {
        const orc_field* f = fields[n->leaf];
        memcpy(n->loc, f->loc, sizeof n->loc);
        n->acc = f->accessible;
        n->local = f->local;
        n->logical = f->logical;
    } else if (is(n, "S")) {
        n->scalar = 1;
    } else if (is(n, "D2C")) {/* D2SecondOrderCentered.hpp:187-204 */
        inherit(n, &t->n[n->child[0]]);
        n->acc.start[d]++, n->acc.end[d]--, n->logical.start[d]++, n->logical.end[d]--, n->local.start[d]++, n->local.end[d]--;
    } else if (is(n, "D1C")) {/* D1FirstOrderCentered.hpp:38-63 */
        inherit(n, &t->n[n->child[0]]);
        if (t->n[n->child[0]].loc[d] == LOC_CENTER) {
            n->loc[d] = LOC_CORNER;
            n->acc.start[d]++, n->local.start[d]++, n->logical.start[d]++;
        } else {
            n->loc[d] = LOC_CENTER;
            n->acc.end[d]--, n->local.end[d]--, n->logical.end[d]--;
        }
    } else if (is(n, "D1Dn")) {/* D1FirstOrderBiasedDownwind.hpp:68-84 */
        inherit(n, &t->n[n->child[0]]);
        n->acc.start[d]++, n->logical.start[d]++, n->local.start[d]++;
    } else if (is(n, "D1Up")) {/* D1FirstOrderBiasedUpwind.hpp:68-84 */
        inherit(n, &t->n[n->child[0]]);
        n->acc.end[d]--, n->logical.end[d]--, n->local.end[d]--;
    } else if (is(n, "WenoDn") || is(n, "WenoUp")) {/* D1WENO53Downwind.hpp:98-109 */
        inherit(n, &t->n[n->child[0]]);
        n->acc.start[d] += 3, n->acc.end[d] -= 3, n->local.start[d] += 3, n->local.end[d] -= 3;
    } else if (is(n, "IntpC2N")) {/* D1Linear.hpp:50-59 */
        inherit(n, &t->n[n->child[0]]);
        n->loc[d] = LOC_CORNER;
        n->acc.start[d]++, n->local.start[d]++, n->logical.start[d]++;
    } else if (is(n, "IntpN2C")) {/* D1Linear.hpp:60-69 */
        inherit(n, &t->n[n->child[0]]);
        n->loc[d] = LOC_CENTER;
        n->acc.end[d]--, n->logical.end[d]--, n->local.end[d]--;
    } else if (is_fl(n, "C2N")) {/* D1FluxLimiter.hpp:155-171: props from arg2 */
        if (t->n[n->child[1]].loc[d] != LOC_CENTER) return 2;
        inherit(n, &t->n[n->child[1]]);
        n->loc[d] = LOC_CORNER;
        n->acc.start[d] += 2, n->acc.end[d] -= 1;
        n->local.start[d] += 2, n->local.end[d] -= 1;
        n->logical.start[d] += 2, n->logical.end[d] -= 1;
    } else if (is_fl(n, "N2C")) {/* D1FluxLimiter.hpp:188-203 */
        if (t->n[n->child[1]].loc[d] != LOC_CORNER) return 2;
        inherit(n, &t->n[n->child[1]]);
        n->loc[d] = LOC_CENTER;
        n->acc.start[d] += 1, n->acc.end[d] -= 2;
        n->local.start[d] += 1, n->local.end[d] -= 2;
        n->logical.start[d] += 1, n->logical.end[d] -= 2;
    } else if (is(n, "Conv")) {/* Convolution.hpp:66-82 */
        int a;
        inherit(n, &t->n[n->child[0]]);
        for (a = 0; a < 3; ++a) {
            const int h = n->conv[a] / 2;
            n->acc.start[a] += h, n->acc.end[a] -= h;
            n->local.start[a] += h, n->local.end[a] -= h;
            n->logical.start[a] += h, n->logical.end[a] -= h;
        }
    } else if (is(n, "Cond")) {/* Conditional.hpp:45-70 */
        onode *a = &t->n[n->child[0]], *b = &t->n[n->child[1]], *cc = &t->n[n->child[2]];
        inherit(n, b);
        n->acc = rcommon(rcommon(a->acc, b->acc), cc->acc);
        n->local = rcommon(rcommon(a->local, b->local), cc->local);
    } else if (n->nchild == 1) {/* UniOpDefMacros.hpp.in:16-26 */
        inherit(n, &t->n[n->child[0]]);
    } else {/* BinOpDefMacros.hpp.in:19-80 */
        onode *a = &t->n[n->child[0]], *b = &t->n[n->child[1]];
        if (a->scalar && b->scalar) n->scalar = 1;
        else if (a->scalar) inherit(n, b);
        else if (b->scalar) inherit(n, a);
        else {
            if (memcmp(a->loc, b->loc, sizeof a->loc)) return 2;
            inherit(n, a);
            n->acc = rcommon(a->acc, b->acc);
            n->local = rcommon(a->local, b->local);
        }
    }
    return 0;
}

typedef struct {
    otree* t;
    orc_field** fields;
    const double* scalars;
    const orc_mesh* m;
} ectx;

static double maxd(double a, double b) { return a < b ? b : a; } /* std::max */

/* D1WENO53Downwind::kernel src/Core/Operator/FDMOperators/D1WENO53Downwind.hpp:136-151 (Upwind :133-151 identical) */
static double weno_core(double d1, double d2, double d3, double d4, double d5) {
    double ddx1 = d1 / 3. - 7. * d2 / 6. + 11. * d3 / 6.;
    double ddx2 = -d2 / 6. + 5 * d3 / 6 + d4 / 3;
    double ddx3 = d3 / 3 + 5 * d4 / 6 - d5 / 6;
    double t1 = d1 - 2 * d2 + d3, t2 = d1 - 4 * d2 + 3 * d3, t3 = d2 - 2 * d3 + d4, t4 = d2 - d4, t5 = d3 - 2 * d4 + d5,
           t6 = 3 * d3 - 4 * d4 + d5;
    double s1 = 13. / 12. * (t1 * t1) + (t2 * t2) / 4;
    double s2 = 13. / 12. * (t3 * t3) + (t4 * t4) / 4.;
    double s3 = 13. / 12. * (t5 * t5) + (t6 * t6) / 4.;
    double eps = 1e-6 * maxd(maxd(maxd(maxd(d1 * d1, d2 * d2), d3 * d3), d4 * d4), d5 * d5) + 1e-99;
    double e1 = s1 + eps, e2 = s2 + eps, e3 = s3 + eps;
    double a1 = .1 / (e1 * e1), a2 = .6 / (e2 * e2), a3 = .3 / (e3 * e3);
    double w1 = a1 / (a1 + a2 + a3), w2 = a2 / (a1 + a2 + a3), w3 = a3 / (a1 + a2 + a3);
    return w1 * ddx1 + w2 * ddx2 + w3 * ddx3;
}

static double mind(double a, double b) { return b < a ? b : a; } /* std::min */

/* FluxLimiterKernels.hpp:32-80; kind = the scheme part of the node name */
static int fl_linear(const char* k) { return !strncmp(k, "Central", 7) || !strncmp(k, "Quick", 5) || !strncmp(k, "Cui", 3) || !strncmp(k, "Fromm", 5) || !strncmp(k, "Lui", 3); }
static double fl_kappa(const char* k, double su, double sf) {/* KappaKernel::eval(slop_u, slop_f) :35-51 */
    double kappa;
    if (!strncmp(k, "Central", 7)) kappa = 1;
    else if (!strncmp(k, "Quick", 5))
        kappa = 0.5;
    else if (!strncmp(k, "Cui", 3))
        kappa = 1. / 3.;
    else if (!strncmp(k, "Fromm", 5))
        kappa = 0.;
    else
        kappa = -1.;
    return (1 + kappa) / 2. * sf + (1 - kappa) / 2. * su;
}
static double fl_r(const char* k, double r) {
    if (!strncmp(k, "Minmod", 6)) return maxd(0., mind(r, 1.));                                    /* :56 */
    if (!strncmp(k, "Superbee", 8)) return maxd(maxd(0., mind(2. * r, 1.)), mind(r, 2.));          /* :61, initializer-list max = left fold */
    if (!strncmp(k, "Muscl", 5)) return maxd(0., mind(mind(2 * r, (r + 1) / 2.), 2.));             /* :66 */
    if (!strncmp(k, "Harmonic", 8)) return (r + fabs(r)) / (r + 1);                                /* :71 */
    return r * (r + 1) / (r * r + 1);                                                              /* vanAlbada :76 */
}
static double fl_side(const char* k, double su, double sf, double y2, double h, int plus) {
    double t;
    if (fl_linear(k)) t = h * 0.5 * fl_kappa(k, su, sf);
    else {
        double r = sf / (su + 1e-16);
        t = h * 0.5 * fl_r(k, r) * su;
    }
    return plus ? y2 + t : y2 - t;
}

static double eval(const ectx* c, int id, const int* g) {
    const onode* n = &c->t->n[id];
    const int d = n->axis;
    int gm[3] = {g[0], g[1], g[2]}, gp[3] = {g[0], g[1], g[2]};
    if (is(n, "F")) return f_get(c->fields[n->leaf], g);
    if (is(n, "S")) return c->scalars[n->leaf];
    if (is_fl(n, "C2N") || is_fl(n, "N2C")) {/* D1FluxLimiterImpl::eval (D1FluxLimiter.hpp:148-151,182-185) */
        const char* k = n->name + 2;
        const orc_mesh* m = c->m;
        const int q = g[d], e = n->child[1], c2n = is_fl(n, "C2N");
        const double uv = eval(c, n->child[0], g);
        double y[5];
        int o, gg[3] = {g[0], g[1], g[2]};
        for (o = -2; o <= 2; ++o) {
            if ((c2n && o == 2) || (!c2n && o == -2)) {
                y[o + 2] = 0.0;
                continue;
            }
            gg[d] = q + o;
            y[o + 2] = eval(c, e, gg);
        }
#define Y(o) y[(o) + 2]
        if (c2n) {
            if (uv > 0.) {/* :46-63 */
                double x1 = mesh_x(m, d, q - 2) + 0.5 * mesh_dx(m, d, q - 2), x2 = mesh_x(m, d, q - 1) + 0.5 * mesh_dx(m, d, q - 1),
                       x3 = mesh_x(m, d, q) + 0.5 * mesh_dx(m, d, q);
                return fl_side(k, (Y(-1) - Y(-2)) / (x2 - x1), (Y(0) - Y(-1)) / (x3 - x2), Y(-1), mesh_dx(m, d, q - 1), 1);
            } else {/* :95-112 */
                double x1 = mesh_x(m, d, q - 1) + mesh_dx(m, d, q - 1) * 0.5, x2 = mesh_x(m, d, q) + mesh_dx(m, d, q) * 0.5,
                       x3 = mesh_x(m, d, q + 1) + mesh_dx(m, d, q + 1) * 0.5;
                return fl_side(k, (Y(1) - Y(0)) / (x3 - x2), (Y(0) - Y(-1)) / (x2 - x1), Y(0), mesh_dx(m, d, q), 0);
            }
        } else {
            if (uv > 0.) {/* :67-84 */
                double x1 = mesh_x(m, d, q - 1), x2 = mesh_x(m, d, q), x3 = mesh_x(m, d, q + 1);
                return fl_side(k, (Y(0) - Y(-1)) / (x2 - x1), (Y(1) - Y(0)) / (x3 - x2), Y(0), mesh_dx(m, d, q), 1);
            } else {/* :116-133 */
                double x1 = mesh_x(m, d, q), x2 = mesh_x(m, d, q + 1), x3 = mesh_x(m, d, q + 2);
                return fl_side(k, (Y(2) - Y(1)) / (x3 - x2), (Y(1) - Y(0)) / (x2 - x1), Y(1), mesh_dx(m, d, q), 0);
            }
        }
#undef Y
    }
    if (is(n, "Conv")) {/* Convolution::eval (Convolution.hpp:45-63): rangeReduce_s, x fastest, identity 0 */
        double acc = 0.0;
        int a, b, z, gg[3];
        for (z = 0; z < n->conv[2]; ++z)
            for (b = 0; b < n->conv[1]; ++b)
                for (a = 0; a < n->conv[0]; ++a) {
                    gg[0] = g[0] + a - n->conv[0] / 2, gg[1] = g[1] + b - n->conv[1] / 2, gg[2] = g[2] + z - n->conv[2] / 2;
                    acc = acc + c->scalars[n->conv[3] + a + n->conv[0] * (b + n->conv[1] * z)] * eval(c, n->child[0], gg);
                }
        return acc;
    }
    if (d >= 0) {
        const int ch = n->child[0], center = c->t->n[ch].loc[d] == LOC_CENTER, q = g[d];
        const orc_mesh* m = c->m;
        gm[d] = q - 1;
        gp[d] = q + 1;
        if (is(n, "D2C")) {/* D2SecondOrderCentered.hpp:161-171 */
            double l = eval(c, ch, gm), cc = eval(c, ch, g), r = eval(c, ch, gp);
            double dxl = !center ? mesh_dx(m, d, q - 1) : (mesh_dx(m, d, q - 1) + mesh_dx(m, d, q)) * 0.5;
            double dxr = !center ? mesh_dx(m, d, q) : (mesh_dx(m, d, q) + mesh_dx(m, d, q + 1)) * 0.5;
            double dxc = (dxl + dxr) * 0.5;
            return ((r - cc) / dxr - (cc - l) / dxl) / dxc;
        }
        if (is(n, "D1C")) {/* D1FirstOrderCentered.hpp:31-36 */
            if (center) return (eval(c, ch, g) - eval(c, ch, gm)) / (mesh_dx(m, d, q - 1) + mesh_dx(m, d, q)) * 2;
            return (eval(c, ch, gp) - eval(c, ch, g)) / (mesh_dx(m, d, q));
        }
        if (is(n, "D1Dn")) /* D1FirstOrderBiasedDownwind.hpp:53-57 */
            return (eval(c, ch, g) - eval(c, ch, gm)) / (!center ? mesh_dx(m, d, q - 1) : (mesh_dx(m, d, q - 1) + mesh_dx(m, d, q)) * 0.5);
        if (is(n, "D1Up")) /* D1FirstOrderBiasedUpwind.hpp:54-58 */
            return (eval(c, ch, gp) - eval(c, ch, g)) / (!center ? mesh_dx(m, d, q) : (mesh_dx(m, d, q) + mesh_dx(m, d, q + 1)) * 0.5);
        if (is(n, "WenoDn") || is(n, "WenoUp")) {
            double p[7], h = mesh_dx(m, d, q);
            int k, gg[3] = {g[0], g[1], g[2]};
            for (k = -3; k <= 3; ++k) {
                gg[d] = q + k;
                p[k + 3] = (is(n, "WenoDn") && k == 3) || (is(n, "WenoUp") && k == -3) ? 0.0 : eval(c, ch, gg);
            }
            if (is(n, "WenoDn")) /* D1WENO53Downwind.hpp:77-86,137-138 */
                return weno_core((p[1] - p[0]) / h, (p[2] - p[1]) / h, (p[3] - p[2]) / h, (p[4] - p[3]) / h, (p[5] - p[4]) / h);
            /* D1WENO53Upwind.hpp:75-84,135-136 */
            return weno_core((p[6] - p[5]) / h, (p[5] - p[4]) / h, (p[4] - p[3]) / h, (p[3] - p[2]) / h, (p[2] - p[1]) / h);
        }
        if (is(n, "IntpC2N")) {/* D1Linear.hpp:36-42 */
            double x1 = mesh_x(m, d, q - 1) + 0.5 * mesh_dx(m, d, q - 1), x2 = mesh_x(m, d, q) + 0.5 * mesh_dx(m, d, q);
            return intp(x1, eval(c, ch, gm), x2, eval(c, ch, g), mesh_x(m, d, q));
        }
        if (is(n, "IntpN2C")) return (eval(c, ch, g) + eval(c, ch, gp)) * 0.5; /* D1Linear.hpp:44, Interpolator.hpp:83 */
    }
    if (is(n, "Cond")) return eval(c, n->child[0], g) != 0.0 ? eval(c, n->child[1], g) : eval(c, n->child[2], g);
    if (n->nchild == 1) {
        double x = eval(c, n->child[0], g);
        if (is(n, "Neg")) return -x;
        if (is(n, "Pos")) return x;
        if (is(n, "Not")) return x == 0.0;
        if (is(n, "Sqrt")) return sqrt(x);
        if (is(n, "Abs")) return fabs(x);
        if (is(n, "Exp")) return exp(x);
        if (is(n, "Log")) return log(x);
        if (is(n, "Sin")) return sin(x);
        if (is(n, "Cos")) return cos(x);
        if (is(n, "Tan")) return tan(x);
        if (is(n, "Tanh")) return tanh(x);
        if (is(n, "Pow2")) return x * x;
        /* AMDS.hpp:56-89: libm, integer-valued results stored as Real */
        if (is(n, "Exp2")) return exp2(x);
        if (is(n, "Expm1")) return expm1(x);
        if (is(n, "Log10")) return log10(x);
        if (is(n, "Log2")) return log2(x);
        if (is(n, "Log1p")) return log1p(x);
        if (is(n, "Cbrt")) return cbrt(x);
        if (is(n, "ASin")) return asin(x);
        if (is(n, "ACos")) return acos(x);
        if (is(n, "ATan")) return atan(x);
        if (is(n, "Sinh")) return sinh(x);
        if (is(n, "Cosh")) return cosh(x);
        if (is(n, "ASinh")) return asinh(x);
        if (is(n, "ACosh")) return acosh(x);
        if (is(n, "ATanh")) return atanh(x);
        if (is(n, "Erf")) return erf(x);
        if (is(n, "Erfc")) return erfc(x);
        if (is(n, "TGamma")) return tgamma(x);
        if (is(n, "LGamma")) return lgamma(x);
        if (is(n, "Ceil")) return ceil(x);
        if (is(n, "Floor")) return floor(x);
        if (is(n, "Trunc")) return trunc(x);
        if (is(n, "Round")) return round(x);
        if (is(n, "LRound")) return (double) lround(x);
        if (is(n, "LLRound")) return (double) llround(x);
        if (is(n, "NearbyInt")) return nearbyint(x);
        if (is(n, "Rint")) return rint(x);
        if (is(n, "LRint")) return (double) lrint(x);
        if (is(n, "LLRint")) return (double) llrint(x);
        if (is(n, "ILogb")) return (double) ilogb(x);
        if (is(n, "Logb")) return logb(x);
    } else {
        double x = eval(c, n->child[0], g), y = eval(c, n->child[1], g);
        if (is(n, "Add")) return x + y;
        if (is(n, "Sub")) return x - y;
        if (is(n, "Mul")) return x * y;
        if (is(n, "Div")) return x / y;
        if (is(n, "Min")) return y < x ? y : x;
        if (is(n, "Max")) return x < y ? y : x;
        if (is(n, "Pow")) return pow(x, y);
        /* AMDS.hpp:40-51 */
        if (is(n, "FMod")) return fmod(x, y);
        if (is(n, "Remainder")) return remainder(x, y);
        if (is(n, "FDim")) return fdim(x, y);
        if (is(n, "Hypot")) return hypot(x, y);
        if (is(n, "ATan2")) return atan2(x, y);
        if (is(n, "Ldexp")) return ldexp(x, (int) y);
        if (is(n, "Scalbn")) return scalbn(x, (int) y);
        if (is(n, "Scalbln")) return scalbn(x, (int) y);
        if (is(n, "Nextafter")) return nextafter(x, y);
        if (is(n, "Nexttoward")) return nextafter(x, y);
        if (is(n, "Copysing")) return copysign(x, y);
        if (is(n, "Lt")) return x < y;
        if (is(n, "Le")) return x <= y;
        if (is(n, "Gt")) return x > y;
        if (is(n, "Ge")) return x >= y;
        if (is(n, "Eq")) return x == y;
        if (is(n, "Ne")) return x != y;
        if (is(n, "And")) return x != 0.0 && y != 0.0;
        if (is(n, "Or")) return x != 0.0 || y != 0.0;
    }
    return NAN;
}

static int build_tree(otree* t, const char* sig, orc_field** fields) {
    char clean[2048];
    int i, j = 0;
    for (i = 0; sig[i] && j < 2047; ++i)
        if (sig[i] != ' ') clean[j++] = sig[i];
    clean[j] = 0;
    memset(t, 0, sizeof *t);
    t->s = clean;
    if (parse_node(t) != 0 || t->err || clean[t->pos] != 0) return 1;
    t->s = NULL;
    return prepare(t, 0, fields);
}

/* out[...] = expr.evalAt(i) for i in [lo,hi) (dense, axis 0 fastest): the body of rangeFor_s
 * (src/Core/Loops/RangeFor.hpp:40-47) applied to FieldAssigner's lambda (src/Core/Loops/FieldAssigner.hpp:48-50). */
int orc_eval(const char* sig, orc_field** fields, const double* scalars, const orc_mesh* m, const int* lo, const int* hi, double* out) {
    static otree t;
    ectx c;
    int g[3];
    long long k = 0;
    int rc = build_tree(&t, sig, fields);
    if (rc) return rc;
    c.t = &t, c.fields = fields, c.scalars = scalars, c.m = m;
    for (g[2] = lo[2]; g[2] < hi[2]; ++g[2])
        for (g[1] = lo[1]; g[1] < hi[1]; ++g[1])
            for (g[0] = lo[0]; g[0] < hi[0]; ++g[0]) out[k++] = eval(&c, 0, g);
    return 0;
}

/* ranges / loc of the prepared expression; which: 0 local 2 accessible 3 logical */
int orc_prepare(const char* sig, orc_field** fields, int which, orc_range* out, int* loc) {
    static otree t;
    int rc = build_tree(&t, sig, fields);
    if (rc) return rc;
    *out = which == 0 ? t.n[0].local : (which == 2 ? t.n[0].acc : t.n[0].logical);
    memcpy(loc, t.n[0].loc, sizeof t.n[0].loc);
    return 0;
}

/* dst (op)= expr over assignable ∩ local with alias-safe semantics (FieldAssigner::assign :27-36: RHS fully evaluated
 * first), then dst.updatePadding() (CartesianField.hpp:231).  op: 0 = 1 += 2 -= 3 *= 4 /= */
int orc_assign(orc_field* dst, int op, const char* sig, orc_field** fields, const double* scalars, const orc_mesh* m) {
    orc_range w = rcommon(dst->assignable, dst->local);
    long long n = 1, k = 0;
    int d, g[3], rc;
    double* tmp;
    for (d = 0; d < 3; ++d) n *= (w.end[d] - w.start[d] > 0 ? w.end[d] - w.start[d] : 0);
    if (n > 0) {
        tmp = (double*) malloc(sizeof(double) * n);
        rc = orc_eval(sig, fields, scalars, m, w.start, w.end, tmp);
        if (rc) {
            free(tmp);
            return rc;
        }
        for (g[2] = w.start[2]; g[2] < w.end[2]; ++g[2])
            for (g[1] = w.start[1]; g[1] < w.end[1]; ++g[1])
                for (g[0] = w.start[0]; g[0] < w.end[0]; ++g[0]) {
                    double* p = &dst->data[f_off(dst, g)];
                    double v = tmp[k++];
                    switch (op) {
                        case 0: *p = v; break;
                        case 1: *p += v; break;
                        case 2: *p -= v; break;
                        case 3: *p *= v; break;
                        default: *p /= v; break;
                    }
                }
        free(tmp);
    }
    orc_update_padding(dst, m);
    return 0;
}

/* ---------------------------------------------------------------------------------------------- decomposition
 * EvenSplitStrategy::gen_split_plan + splitMap_impl (src/Core/Parallel/EvenSplitStrategy.hpp:57-192) */
int orc_split_even(int dim, const int* mesh_start, const int* mesh_end, int nproc, orc_range* out) {
    int s[3] = {0, 0, 0}, e[3] = {1, 1, 1}, order[3] = {0, 1, 2}, splits[2][3] = {{1, 1, 1}, {1, 1, 1}}, idx[3] = {0, 0, 0};
    double cost[2] = {0, 0};
    int i, j, strat, rank;
    const int* sp;
    for (i = 0; i < dim; ++i) s[i] = mesh_start[i], e[i] = mesh_end[i] - 1;
    if (nproc == 1) {
        for (i = 0; i < 3; ++i) out[0].start[i] = s[i], out[0].end[i] = e[i];
        return 0;
    }
    for (i = 1; i < dim; ++i) /* insertion sort by extent (std::sort on <= 16 elements) */
        for (j = i; j > 0 && (e[order[j]] - s[order[j]]) < (e[order[j - 1]] - s[order[j - 1]]); --j) {
            int t = order[j];
            order[j] = order[j - 1];
            order[j - 1] = t;
        }
    for (strat = 0; strat < 2; ++strat) {
        int remain_vol = 1, remain_proc = nproc;
        for (i = 0; i < dim; ++i) remain_vol *= e[i] - s[i];
        for (i = 0; i < dim; ++i) {
            int factors[4096], nf = 0, p, n, ax = order[i];
            double target = pow(remain_proc * 1.0 / remain_vol, 1. / (dim - i)) * (e[ax] - s[ax]);
            for (j = 1; j <= remain_proc; ++j)
                if (remain_proc % j == 0) factors[nf++] = j;
            for (p = 0; p < nf && factors[p] < target; ++p) {} /* lower_bound */
            if (strat == 0) n = i == dim - 1 ? remain_proc : (p != nf ? factors[p] : nproc);
            else
                n = i == dim - 1 ? remain_proc : (p != 0 ? factors[p - 1] : 1);
            splits[strat][ax] = n;
            remain_proc /= n;
            remain_vol /= e[ax] - s[ax];
        }
        for (i = 0; i < dim; ++i) cost[strat] += 1.0 * splits[strat][i] / (e[i] - s[i]);
    }
    sp = cost[0] <= cost[1] ? splits[0] : splits[1];
    for (rank = 0; rank < nproc; ++rank) {
        for (i = 0; i < 3; ++i) out[rank].start[i] = 0, out[rank].end[i] = 1;
        for (i = 0; i < dim; ++i) {
            out[rank].start[i] = s[i] + (e[i] - s[i]) / sp[i] * idx[i];
            out[rank].end[i] = idx[i] < sp[i] - 1 ? out[rank].start[i] + (e[i] - s[i]) / sp[i] : e[i];
        }
        for (i = 0; i < dim; ++i) {
            if (++idx[i] < sp[i]) break;
            if (i < dim - 1) idx[i] = 0;
        }
    }
    return 0;
}
