/*
 * pgo_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the pose-graph-optimization hot path of
 * jgsimard/RustRobotics, written from the behaviour of
 *   src/mapping/g2o.rs                        (loader)
 *   src/mapping/pose_graph_optimization.rs    (linearise, assemble, update, chi2)
 * Every function cites the reference file:line it follows.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the CUDA product path never does.
 *
 * The sparse direct solve (reference: russell_sparse 0.7.1 -> SuiteSparse
 * UMFPACK, pose_graph_optimization.rs:130-141, neither vendored nor installed
 * here) is NOT in this file: oracle/oracle.py hands the CSC matrix built by
 * og_coo_to_csc() to SciPy's SuperLU, an exact sparse LU like UMFPACK.
 *
 * Parity pinning: oracle/oracle.py + tests/test_oracle_kat.py check this code
 * against every known-answer value in the reference's own tests
 * (g2o.rs:149-175, pose_graph_optimization.rs:580-739).
 *
 * SE(3): the reference has no SE(3) optimisation (todo!() at :241,:357,:570);
 * the SE(3) functions below restate THIS repo's documented semantics
 * (DESIGN.md "SE3 semantics") -- parity unpinned.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { OG_SE2 = 0, OG_XY = 1, OG_SE3 = 2 };           /* Node kinds, pose_graph_optimization.rs:149-154 */
enum { OG_E_SE2 = 0, OG_E_SE2_XY = 1, OG_E_SE3 = 2 }; /* Edge kinds, :21-26 */

static const int KIND_DIM[3] = {3, 2, 6};   /* tangent dim = lut stride, g2o.rs:61,68,77 */
static const int KIND_NVAL[3] = {3, 2, 7};  /* numbers on a VERTEX line */
static const int KIND_NSTATE[3] = {4, 2, 7}; /* stored state: SE2 = x,y,re,im ; SE3 = x,y,z,qw,qx,qy,qz */
static const int EKIND_NMEAS[3] = {3, 2, 7};
static const int EKIND_DIM[3] = {3, 2, 6};

typedef struct {
    int64_t n_vertices, n_edges, len;
    /* vertices in file (lut) order */
    uint32_t *vid;
    uint8_t *vkind;
    int64_t *voffset;   /* lut value: scalar offset, g2o.rs:60,67,76 */
    double *vstate;     /* stride 8 per vertex */
    /* edges in file order */
    uint8_t *ekind;
    int64_t *efrom, *eto; /* vertex INDEX (position in lut order), resolved from ids */
    double *emeas;      /* stride 8: SE2 = x,y,re,im ; XY = x,y ; SE3 = x,y,z,qw,qx,qy,qz */
    double *einfo;      /* stride 36: full symmetric dim x dim row-major, g2o.rs:88-93 */
    int64_t cap_v, cap_e;
} og_graph;

#define VS 8
#define IS 36

/* ------------------------------------------------------------------ small algebra */

/* Isometry2 as (tx, ty, re, im); nalgebra composition (R1,t1)*(R2,t2) = (R1R2, t1 + R1 t2) */
static void iso2_mul(const double *a, const double *b, double *o) {
    double tx = a[0] + (a[2] * b[0] - a[3] * b[1]);
    double ty = a[1] + (a[3] * b[0] + a[2] * b[1]);
    double re = a[2] * b[2] - a[3] * b[3];
    double im = a[2] * b[3] + a[3] * b[2];
    o[0] = tx; o[1] = ty; o[2] = re; o[3] = im;
}
static void iso2_inv(const double *a, double *o) {
    /* rotation^-1 = conjugate ; translation = -(R^-1 t) */
    double re = a[2], im = -a[3];
    o[0] = -(re * a[0] - im * a[1]);
    o[1] = -(im * a[0] + re * a[1]);
    o[2] = re; o[3] = im;
}

/* pose2D_pose2D_constraint + v3, pose_graph_optimization.rs:434-447 :
 *   E = z^-1 * x1^-1 * x2 ;  e = (E.t.x, E.t.y, atan2(E.R.im, E.R.re)) */
static void pose_pose_error(const double *x1, const double *x2, const double *z, double *e) {
    double zi[4], x1i[4], t[4], E[4];
    iso2_inv(z, zi);
    iso2_inv(x1, x1i);
    iso2_mul(zi, x1i, t);
    iso2_mul(t, x2, E);
    e[0] = E[0]; e[1] = E[1]; e[2] = atan2(E[3], E[2]);
}

/* linearize_pose2D_pose2D_constraint, :457-486.  A, B row-major 3x3. */
static void pose_pose_jac(const double *x1, const double *x2, const double *z, double *A, double *B) {
    double c1 = x1[2], s1 = x1[3], cz = z[2], sz = z[3];
    /* M = Rz^T R1^T  (z_rot.inverse() * x1_rot.inverse(), :466) */
    double m11 = cz * c1 - sz * s1, m12 = cz * s1 + sz * c1;
    double m21 = -m12, m22 = m11;
    /* xr1d = deriv * R1 with deriv = [[0,-1],[1,0]] (:462,:467) = [[-s1,-c1],[c1,-s1]] ; transpose: */
    double d11 = -s1, d12 = c1, d21 = -c1, d22 = -s1; /* xr1d^T */
    double dx = x2[0] - x1[0], dy = x2[1] - x1[1];
    double v0 = d11 * dx + d12 * dy, v1 = d21 * dx + d22 * dy;
    /* a_12 = Rz^T * xr1d^T * (t2 - t1)  (:468-469) */
    double a0 = cz * v0 + sz * v1, a1 = -sz * v0 + cz * v1;
    A[0] = -m11; A[1] = -m12; A[2] = a0;
    A[3] = -m21; A[4] = -m22; A[5] = a1;
    A[6] = 0.0;  A[7] = 0.0;  A[8] = -1.0;
    B[0] = m11; B[1] = m12; B[2] = 0.0;
    B[3] = m21; B[4] = m22; B[5] = 0.0;
    B[6] = 0.0; B[7] = 0.0; B[8] = 1.0;
}

/* pose2D_landmark2D_constraint, :449-455 : e = R1^T (l - t1) - z */
static void pose_landmark_error(const double *x, const double *l, const double *z, double *e) {
    double c = x[2], s = x[3], dx = l[0] - x[0], dy = l[1] - x[1];
    e[0] = (c * dx + s * dy) - z[0];
    e[1] = (-s * dx + c * dy) - z[1];
}

/* linearize_pose_landmark_constraint, :516-535.  A 2x3, B 2x2 row-major. */
static void pose_landmark_jac(const double *x, const double *l, double *A, double *B) {
    double c = x[2], s = x[3], dx = l[0] - x[0], dy = l[1] - x[1];
    /* a_1 = -R1^T ; a_2 = (deriv*R1)^T (l - t1) */
    A[0] = -c; A[1] = -s; A[2] = -s * dx + c * dy;
    A[3] = s;  A[4] = -c; A[5] = -c * dx - s * dy;
    B[0] = c;  B[1] = s;
    B[2] = -s; B[3] = c;
}

/* ------------------------------------------------------------------ SE(3), repo-defined semantics (parity unpinned) */
/* quaternion (w,x,y,z) helpers */
static void q_mul(const double *a, const double *b, double *o) {
    double w = a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3];
    double x = a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2];
    double y = a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1];
    double z = a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
static void q_to_R(const double *q, double *R) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2*(y*y + z*z); R[1] = 2*(x*y - w*z);     R[2] = 2*(x*z + w*y);
    R[3] = 2*(x*y + w*z);     R[4] = 1 - 2*(x*x + z*z); R[5] = 2*(y*z - w*x);
    R[6] = 2*(x*z - w*y);     R[7] = 2*(y*z + w*x);     R[8] = 1 - 2*(x*x + y*y);
}
static void q_log(const double *q, double *phi) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    if (w < 0) { w = -w; x = -x; y = -y; z = -z; }
    double n = sqrt(x*x + y*y + z*z);
    double k = (n < 1e-12) ? 2.0 / w : 2.0 * atan2(n, w) / n;
    phi[0] = k * x; phi[1] = k * y; phi[2] = k * z;
}
static void q_exp(const double *phi, double *q) {
    double th = sqrt(phi[0]*phi[0] + phi[1]*phi[1] + phi[2]*phi[2]);
    double k = (th < 1e-12) ? 0.5 - th*th/48.0 : sin(0.5*th) / th;
    q[0] = cos(0.5*th); q[1] = k*phi[0]; q[2] = k*phi[1]; q[3] = k*phi[2];
}
static void m3_mul(const double *a, const double *b, double *o) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += a[3*i+k] * b[3*k+j];
        o[3*i+j] = s;
    }
}
static void m3_tmul(const double *a, const double *b, double *o) { /* a^T b */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += a[3*k+i] * b[3*k+j];
        o[3*i+j] = s;
    }
}
static void skew3(const double *v, double *S) {
    S[0] = 0; S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2]; S[4] = 0; S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
/* inverse right Jacobian of SO(3): I + 1/2 [p]x + k [p]x^2 */
static void so3_jr_inv(const double *p, double *J) {
    double th2 = p[0]*p[0] + p[1]*p[1] + p[2]*p[2], th = sqrt(th2);
    double k = (th < 1e-5) ? (1.0/12.0 + th2/720.0) : (1.0/th2 - (1.0 + cos(th)) / (2.0 * th * sin(th)));
    double S[9], S2[9];
    skew3(p, S); m3_mul(S, S, S2);
    for (int i = 0; i < 9; i++) J[i] = 0.5 * S[i] + k * S2[i];
    J[0] += 1; J[4] += 1; J[8] += 1;
}
/* state: t(3), q(w,x,y,z).  e = [ Rz^T(R1^T(t2-t1) - tz) ; Log(qz^-1 q1^-1 q2) ] */
static void se3_error_jac(const double *x1, const double *x2, const double *z, double *e, double *A, double *B) {
    double R1[9], R2[9], Rz[9];
    q_to_R(x1 + 3, R1); q_to_R(x2 + 3, R2); q_to_R(z + 3, Rz);
    double d[3] = {x2[0]-x1[0], x2[1]-x1[1], x2[2]-x1[2]}, u[3], w[3];
    for (int i = 0; i < 3; i++) u[i] = R1[i]*d[0] + R1[3+i]*d[1] + R1[6+i]*d[2];     /* R1^T d */
    for (int i = 0; i < 3; i++) w[i] = u[i] - z[i];
    for (int i = 0; i < 3; i++) e[i] = Rz[i]*w[0] + Rz[3+i]*w[1] + Rz[6+i]*w[2];     /* Rz^T (.) */
    double qzi[4] = {z[3], -z[4], -z[5], -z[6]}, q1i[4] = {x1[3], -x1[4], -x1[5], -x1[6]}, t[4], qe[4];
    q_mul(qzi, q1i, t); q_mul(t, x2 + 3, qe);
    q_log(qe, e + 3);
    if (!A) return;
    double M[9], RzT_R1T[9], Su[9], T[9], Jri[9], R2T_R1[9];
    m3_mul(R1, Rz, M);                 /* R1 Rz ; (R1 Rz)^T = Rz^T R1^T */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) RzT_R1T[3*i+j] = M[3*j+i];
    skew3(u, Su); m3_tmul(Rz, Su, T);  /* Rz^T [R1^T d]x */
    so3_jr_inv(e + 3, Jri);
    m3_tmul(R2, R1, R2T_R1);
    double JR[9]; m3_mul(Jri, R2T_R1, JR);
    memset(A, 0, 36 * sizeof(double)); memset(B, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        A[6*i + j] = -RzT_R1T[3*i+j];
        A[6*i + 3 + j] = T[3*i+j];
        A[6*(3+i) + 3 + j] = -JR[3*i+j];
        B[6*i + j] = RzT_R1T[3*i+j];
        B[6*(3+i) + 3 + j] = Jri[3*i+j];
    }
}

/* ------------------------------------------------------------------ graph container */

static og_graph *og_alloc(void) { return (og_graph *)calloc(1, sizeof(og_graph)); }

void og_free(og_graph *g) {
    if (!g) return;
    free(g->vid); free(g->vkind); free(g->voffset); free(g->vstate);
    free(g->ekind); free(g->efrom); free(g->eto); free(g->emeas); free(g->einfo);
    free(g);
}

static void grow_v(og_graph *g) {
    if (g->n_vertices < g->cap_v) return;
    g->cap_v = g->cap_v ? 2 * g->cap_v : 1024;
    g->vid = realloc(g->vid, g->cap_v * sizeof(uint32_t));
    g->vkind = realloc(g->vkind, g->cap_v);
    g->voffset = realloc(g->voffset, g->cap_v * sizeof(int64_t));
    g->vstate = realloc(g->vstate, g->cap_v * VS * sizeof(double));
}
static void grow_e(og_graph *g) {
    if (g->n_edges < g->cap_e) return;
    g->cap_e = g->cap_e ? 2 * g->cap_e : 1024;
    g->ekind = realloc(g->ekind, g->cap_e);
    g->efrom = realloc(g->efrom, g->cap_e * sizeof(int64_t));
    g->eto = realloc(g->eto, g->cap_e * sizeof(int64_t));
    g->emeas = realloc(g->emeas, g->cap_e * VS * sizeof(double));
    g->einfo = realloc(g->einfo, g->cap_e * IS * sizeof(double));
}

/* vertex values as on a g2o line -> stored state.  iso2 (g2o.rs:14-16) stores the
 * angle as a unit complex (cos, sin).  SE3: g2o order is x y z qx qy qz qw; stored (w,x,y,z)
 * normalised (the reference's iso3 passes them to Quaternion::new in the wrong order,
 * g2o.rs:18-21 -- deliberately not reproduced). */
static void set_state(int kind, const double *val, double *st) {
    memset(st, 0, VS * sizeof(double));
    if (kind == OG_SE2) { st[0] = val[0]; st[1] = val[1]; st[2] = cos(val[2]); st[3] = sin(val[2]); }
    else if (kind == OG_XY) { st[0] = val[0]; st[1] = val[1]; }
    else {
        double n = sqrt(val[3]*val[3] + val[4]*val[4] + val[5]*val[5] + val[6]*val[6]);
        st[0] = val[0]; st[1] = val[1]; st[2] = val[2];
        st[3] = val[6] / n; st[4] = val[3] / n; st[5] = val[4] / n; st[6] = val[5] / n;
    }
}

static void add_vertex(og_graph *g, uint32_t id, int kind, const double *val) {
    grow_v(g);
    int64_t i = g->n_vertices++;
    g->vid[i] = id; g->vkind[i] = (uint8_t)kind; g->voffset[i] = g->len;
    set_state(kind, val, g->vstate + VS * i);
    g->len += KIND_DIM[kind];
}

/* information: upper triangle row-major -> full symmetric (g2o.rs:88-93, 106-110, 125-133) */
static void add_edge(og_graph *g, int kind, int64_t from, int64_t to, const double *meas, const double *upper) {
    grow_e(g);
    int64_t k = g->n_edges++;
    int d = EKIND_DIM[kind];
    g->ekind[k] = (uint8_t)kind; g->efrom[k] = from; g->eto[k] = to;
    set_state(kind == OG_E_SE2 ? OG_SE2 : kind == OG_E_SE2_XY ? OG_XY : OG_SE3, meas, g->emeas + VS * k);
    double *W = g->einfo + IS * k;
    memset(W, 0, IS * sizeof(double));
    int p = 0;
    for (int r = 0; r < d; r++) for (int c = r; c < d; c++) { W[r * d + c] = upper[p]; W[c * d + r] = upper[p]; p++; }
}

/* Build from flat arrays (ids resolved through an id->index map built here).
 * vertex_values / edge_meas / edge_info_upper are packed back to back with the per-kind
 * counts (3|2|7 ; 3|2|7 ; 6|3|21). Returns NULL if an edge names an unknown vertex id. */
static int cmp_idpair(const void *a, const void *b) {
    const int64_t *x = a, *y = b;
    return (x[0] > y[0]) - (x[0] < y[0]);
}
og_graph *og_create(int64_t nv, const uint32_t *vid, const uint8_t *vkind, const double *vval,
                    int64_t ne, const uint8_t *ekind, const uint32_t *efrom, const uint32_t *eto,
                    const double *emeas, const double *einfo_upper) {
    og_graph *g = og_alloc();
    const double *p = vval;
    for (int64_t i = 0; i < nv; i++) { add_vertex(g, vid[i], vkind[i], p); p += KIND_NVAL[vkind[i]]; }
    /* id -> index ; later duplicates of an id overwrite earlier ones like FxHashMap::insert (g2o.rs:59-60) */
    int64_t *map = malloc(2 * sizeof(int64_t) * (nv ? nv : 1));
    for (int64_t i = 0; i < nv; i++) { map[2*i] = vid[i]; map[2*i+1] = i; }
    qsort(map, nv, 2 * sizeof(int64_t), cmp_idpair);
    const double *pm = emeas, *pi = einfo_upper;
    for (int64_t k = 0; k < ne; k++) {
        int64_t idx[2];
        uint32_t want[2] = {efrom[k], eto[k]};
        for (int s = 0; s < 2; s++) {
            int64_t lo = 0, hi = nv - 1, f = -1;
            while (lo <= hi) { int64_t mid = (lo + hi) / 2;
                if (map[2*mid] < want[s]) lo = mid + 1; else if (map[2*mid] > want[s]) hi = mid - 1;
                else { f = mid; lo = mid + 1; } }   /* last among equal ids */
            if (f < 0) { free(map); og_free(g); return NULL; }
            idx[s] = map[2*f+1];
        }
        int d = EKIND_DIM[ekind[k]];
        add_edge(g, ekind[k], idx[0], idx[1], pm, pi);
        pm += EKIND_NMEAS[ekind[k]]; pi += d * (d + 1) / 2;
    }
    free(map);
    return g;
}

/* parse_g2o, g2o.rs:35-143.  Lines are split on ' ' with empty tokens dropped (:52); vertices
 * get lut offsets in line order; an unknown tag is an error (unimplemented!, :138), so is a blank
 * line (index panic at :53) and a wrong number of numeric fields (todo!() arms).  Edge endpoints
 * are kept as ids and resolved after the whole file is read, as the reference resolves them
 * lazily through the lut at use time (:312-313). */
#define MAXTOK 40
/* line[1].parse::<u32>()? (g2o.rs:55,80-81): digits only, optional '+', must fit 32 bits */
static int parse_id(const char *t, uint32_t *out) {
    if (*t == '+') t++;
    if (!*t) return 0;
    uint64_t a = 0;
    for (; *t; t++) { if (*t < '0' || *t > '9') return 0; a = a * 10 + (uint64_t)(*t - '0'); if (a > 0xffffffffull) return 0; }
    *out = (uint32_t)a;
    return 1;
}
og_graph *og_parse_g2o(const char *path, char *err, int errlen) {
    FILE *f = fopen(path, "r");
    if (!f) { snprintf(err, errlen, "cannot open %s", path); return NULL; }
    int64_t nv = 0, ne = 0, cv = 1024, ce = 1024, nvv = 0, cvv = 4096, nem = 0, cem = 4096, nei = 0, cei = 8192;
    uint32_t *vid = malloc(cv * 4), *efrom = malloc(ce * 4), *eto = malloc(ce * 4);
    uint8_t *vkind = malloc(cv), *ekind = malloc(ce);
    double *vval = malloc(cvv * 8), *emeas = malloc(cem * 8), *einfo = malloc(cei * 8);
    char *line = NULL; size_t cap = 0; ssize_t n; int64_t lineno = 0; int bad = 0;
    while (!bad && (n = getline(&line, &cap, f)) >= 0) {
        lineno++;
        while (n > 0 && (line[n-1] == '\n' || line[n-1] == '\r')) line[--n] = 0;
        char *tok[MAXTOK]; int nt = 0; char *sv = NULL;
        for (char *t = strtok_r(line, " ", &sv); t && nt < MAXTOK; t = strtok_r(NULL, " ", &sv)) tok[nt++] = t;
        if (nt == 0) { snprintf(err, errlen, "line %ld: blank line", (long)lineno); bad = 1; break; }
        int vk = -1, ek = -1;
        if (!strcmp(tok[0], "VERTEX_SE2")) vk = OG_SE2;
        else if (!strcmp(tok[0], "VERTEX_XY")) vk = OG_XY;
        else if (!strcmp(tok[0], "VERTEX_SE3:QUAT")) vk = OG_SE3;
        else if (!strcmp(tok[0], "EDGE_SE2")) ek = OG_E_SE2;
        else if (!strcmp(tok[0], "EDGE_SE2_XY")) ek = OG_E_SE2_XY;
        else if (!strcmp(tok[0], "EDGE_SE3:QUAT")) ek = OG_E_SE3;
        else { snprintf(err, errlen, "line %ld: not implemented: %s", (long)lineno, tok[0]); bad = 1; break; }
        double num[MAXTOK]; int nn = 0, first = vk >= 0 ? 2 : 3;
        for (int i = first; i < nt; i++) { char *end; num[nn++] = strtod(tok[i], &end);
            if (*end) { snprintf(err, errlen, "line %ld: bad number '%s'", (long)lineno, tok[i]); bad = 1; } }
        if (bad) break;
        if (vk >= 0) {
            if (nt < 2 || nn != KIND_NVAL[vk]) { snprintf(err, errlen, "line %ld: wrong field count", (long)lineno); bad = 1; break; }
            if (nv == cv) { cv *= 2; vid = realloc(vid, cv * 4); vkind = realloc(vkind, cv); }
            if (nvv + 8 > cvv) { cvv *= 2; vval = realloc(vval, cvv * 8); }
            if (!parse_id(tok[1], &vid[nv])) { snprintf(err, errlen, "line %ld: bad vertex id", (long)lineno); bad = 1; break; }
            vkind[nv] = (uint8_t)vk; nv++;
            memcpy(vval + nvv, num, nn * 8); nvv += nn;
        } else {
            int d = EKIND_DIM[ek], want = EKIND_NMEAS[ek] + d * (d + 1) / 2;
            if (nt < 3 || nn != want) { snprintf(err, errlen, "line %ld: wrong field count", (long)lineno); bad = 1; break; }
            if (ne == ce) { ce *= 2; efrom = realloc(efrom, ce * 4); eto = realloc(eto, ce * 4); ekind = realloc(ekind, ce); }
            if (nem + 8 > cem) { cem *= 2; emeas = realloc(emeas, cem * 8); }
            if (nei + 24 > cei) { cei *= 2; einfo = realloc(einfo, cei * 8); }
            if (!parse_id(tok[1], &efrom[ne]) || !parse_id(tok[2], &eto[ne])) { snprintf(err, errlen, "line %ld: bad edge endpoint id", (long)lineno); bad = 1; break; }
            ekind[ne] = (uint8_t)ek; ne++;
            memcpy(emeas + nem, num, EKIND_NMEAS[ek] * 8); nem += EKIND_NMEAS[ek];
            memcpy(einfo + nei, num + EKIND_NMEAS[ek], (want - EKIND_NMEAS[ek]) * 8); nei += want - EKIND_NMEAS[ek];
        }
    }
    free(line); fclose(f);
    og_graph *g = NULL;
    if (!bad) {
        g = og_create(nv, vid, vkind, vval, ne, ekind, efrom, eto, emeas, einfo);
        if (!g) snprintf(err, errlen, "edge references unknown vertex id");
    }
    free(vid); free(efrom); free(eto); free(vkind); free(ekind); free(vval); free(emeas); free(einfo);
    return g;
}

int64_t og_num_vertices(const og_graph *g) { return g->n_vertices; }
int64_t og_num_edges(const og_graph *g) { return g->n_edges; }
int64_t og_len(const og_graph *g) { return g->len; }

/* vertex table: ids, kinds, lut offsets, and values in g2o form (SE2: x,y,atan2(im,re); XY: x,y;
 * SE3: x,y,z,qx,qy,qz,qw), packed back to back */
void og_get_vertices(const og_graph *g, uint32_t *vid, uint8_t *vkind, int64_t *voffset, double *vval) {
    double *p = vval;
    for (int64_t i = 0; i < g->n_vertices; i++) {
        const double *s = g->vstate + VS * i;
        if (vid) vid[i] = g->vid[i];
        if (vkind) vkind[i] = g->vkind[i];
        if (voffset) voffset[i] = g->voffset[i];
        if (!vval) continue;
        if (g->vkind[i] == OG_SE2) { p[0] = s[0]; p[1] = s[1]; p[2] = atan2(s[3], s[2]); p += 3; }
        else if (g->vkind[i] == OG_XY) { p[0] = s[0]; p[1] = s[1]; p += 2; }
        else { p[0] = s[0]; p[1] = s[1]; p[2] = s[2]; p[3] = s[4]; p[4] = s[5]; p[5] = s[6]; p[6] = s[3]; p += 7; }
    }
}
int64_t og_num_vertex_values(const og_graph *g) {
    int64_t n = 0; for (int64_t i = 0; i < g->n_vertices; i++) n += KIND_NVAL[g->vkind[i]]; return n;
}
/* raw stored state (stride 8), for bit-level comparisons of the unit complex */
void og_get_state(const og_graph *g, double *out) { memcpy(out, g->vstate, g->n_vertices * VS * sizeof(double)); }
void og_set_state(og_graph *g, const double *in) { memcpy(g->vstate, in, g->n_vertices * VS * sizeof(double)); }

/* edge table in file order: kind, endpoint vertex indices, measurement/information in g2o form */
void og_get_edges(const og_graph *g, uint8_t *ekind, int64_t *from_idx, int64_t *to_idx) {
    for (int64_t k = 0; k < g->n_edges; k++) {
        if (ekind) ekind[k] = g->ekind[k];
        if (from_idx) from_idx[k] = g->efrom[k];
        if (to_idx) to_idx[k] = g->eto[k];
    }
}
void og_get_edge_data(const og_graph *g, uint32_t *from_id, uint32_t *to_id, double *meas, double *info_upper) {
    double *pm = meas, *pi = info_upper;
    for (int64_t k = 0; k < g->n_edges; k++) {
        int ek = g->ekind[k], d = EKIND_DIM[ek];
        const double *z = g->emeas + VS * k, *W = g->einfo + IS * k;
        from_id[k] = g->vid[g->efrom[k]]; to_id[k] = g->vid[g->eto[k]];
        if (ek == OG_E_SE2) { pm[0] = z[0]; pm[1] = z[1]; pm[2] = atan2(z[3], z[2]); pm += 3; }
        else if (ek == OG_E_SE2_XY) { pm[0] = z[0]; pm[1] = z[1]; pm += 2; }
        else { pm[0] = z[0]; pm[1] = z[1]; pm[2] = z[2]; pm[3] = z[4]; pm[4] = z[5]; pm[5] = z[6]; pm[6] = z[3]; pm += 7; }
        for (int r = 0; r < d; r++) for (int c = r; c < d; c++) *pi++ = W[r * d + c];
    }
}

/* e, A, B of one edge (row-major, A is d x dim(from), B is d x dim(to)) -- for the Jacobian KATs
 * (pose_graph_optimization.rs:633-722) */
void og_edge_linearize(const og_graph *g, int64_t k, double *e, double *A, double *B) {
    const double *x1 = g->vstate + VS * g->efrom[k], *x2 = g->vstate + VS * g->eto[k], *z = g->emeas + VS * k;
    if (g->ekind[k] == OG_E_SE2) { pose_pose_error(x1, x2, z, e); pose_pose_jac(x1, x2, z, A, B); }
    else if (g->ekind[k] == OG_E_SE2_XY) { pose_landmark_error(x1, x2, z, e); pose_landmark_jac(x1, x2, A, B); }
    else se3_error_jac(x1, x2, z, e, A, B);
}

/* global_error, :537-574 : sum over edges (file order, left to right) of e^T Omega e, no 1/2 */
double og_global_error(const og_graph *g) {
    double total = 0.0;
    for (int64_t k = 0; k < g->n_edges; k++) {
        const double *x1 = g->vstate + VS * g->efrom[k], *x2 = g->vstate + VS * g->eto[k];
        const double *z = g->emeas + VS * k, *W = g->einfo + IS * k;
        double e[6]; int d = EKIND_DIM[g->ekind[k]];
        if (g->ekind[k] == OG_E_SE2) pose_pose_error(x1, x2, z, e);
        else if (g->ekind[k] == OG_E_SE2_XY) pose_landmark_error(x1, x2, z, e);
        else se3_error_jac(x1, x2, z, e, NULL, NULL);
        double s = 0.0;
        for (int c = 0; c < d; c++) { double t = 0.0; for (int r = 0; r < d; r++) t += e[r] * W[r * d + c]; s += t * e[c]; }
        total += s;
    }
    return total;
}

/* number of SparseMatrix::put calls build_linear_system makes (:184-187, :332-334, :362-366) */
int64_t og_put_count(const og_graph *g, int lm) {
    int64_t n = 0; int prior = 0;
    for (int64_t k = 0; k < g->n_edges; k++) {
        int di = KIND_DIM[g->vkind[g->efrom[k]]], dj = KIND_DIM[g->vkind[g->eto[k]]];
        n += di * di + 2 * di * dj + dj * dj;
        if (!prior && (g->ekind[k] == OG_E_SE2 || g->ekind[k] == OG_E_SE3)) { n += di; prior = 1; }
    }
    if (lm) n += g->len;
    return n;
}

/* C = X^T W Y ; X is d x dx, W d x d, Y d x dy ; all row-major */
static void xtwy(int d, int dx, int dy, const double *X, const double *W, const double *Y, double *C) {
    double WY[36];
    for (int r = 0; r < d; r++) for (int c = 0; c < dy; c++) {
        double s = 0.0; for (int k = 0; k < d; k++) s += W[r * d + k] * Y[k * dy + c];
        WY[r * dy + c] = s;
    }
    for (int r = 0; r < dx; r++) for (int c = 0; c < dy; c++) {
        double s = 0.0; for (int k = 0; k < d; k++) s += X[k * dx + r] * WY[k * dy + c];
        C[r * dy + c] = s;
    }
}

/* build_linear_system, :305-369, with update_linear_system :165-192 and set_matrix/set_vector
 * :194-212.  Emits the COO triplets in the reference's exact put order:
 *   per edge: H_ii, H_ij, H_ji, H_jj, each row-major, zeros included (:200-204);
 *   after the FIRST pose-pose edge: +1e7 on the diagonal of its `from` (:330-336);
 *   LM only, at the end: lambda on every diagonal (:362-366).
 * b accumulates A^T W e / B^T W e (:189-190) and is negated at the end (:361).
 * (For SE3 graphs -- repo-defined -- the prior goes on the 6 diagonals of the first edge's from.) */
int64_t og_build_linear_system(const og_graph *g, double lambda, int lm,
                               int32_t *ci, int32_t *cj, double *cv, double *b) {
    int64_t n = 0; int need_prior = 1;
    memset(b, 0, g->len * sizeof(double));
    for (int64_t k = 0; k < g->n_edges; k++) {
        int64_t vi = g->efrom[k], vj = g->eto[k];
        const double *x1 = g->vstate + VS * vi, *x2 = g->vstate + VS * vj;
        const double *z = g->emeas + VS * k, *W = g->einfo + IS * k;
        int d = EKIND_DIM[g->ekind[k]], di = KIND_DIM[g->vkind[vi]], dj = KIND_DIM[g->vkind[vj]];
        int64_t oi = g->voffset[vi], oj = g->voffset[vj];
        double e[6], A[36], B[36];
        if (g->ekind[k] == OG_E_SE2) { pose_pose_error(x1, x2, z, e); pose_pose_jac(x1, x2, z, A, B); }
        else if (g->ekind[k] == OG_E_SE2_XY) { pose_landmark_error(x1, x2, z, e); pose_landmark_jac(x1, x2, A, B); }
        else se3_error_jac(x1, x2, z, e, A, B);
        double Hii[36], Hij[36], Hjj[36], bi[6], bj[6];
        xtwy(d, di, di, A, W, A, Hii);
        xtwy(d, di, dj, A, W, B, Hij);
        xtwy(d, dj, dj, B, W, B, Hjj);
        xtwy(d, di, 1, A, W, e, bi);
        xtwy(d, dj, 1, B, W, e, bj);
        for (int r = 0; r < di; r++) for (int c = 0; c < di; c++) { ci[n] = (int32_t)(oi + r); cj[n] = (int32_t)(oi + c); cv[n++] = Hii[r * di + c]; }
        for (int r = 0; r < di; r++) for (int c = 0; c < dj; c++) { ci[n] = (int32_t)(oi + r); cj[n] = (int32_t)(oj + c); cv[n++] = Hij[r * dj + c]; }
        for (int r = 0; r < dj; r++) for (int c = 0; c < di; c++) { ci[n] = (int32_t)(oj + r); cj[n] = (int32_t)(oi + c); cv[n++] = Hij[c * dj + r]; }
        for (int r = 0; r < dj; r++) for (int c = 0; c < dj; c++) { ci[n] = (int32_t)(oj + r); cj[n] = (int32_t)(oj + c); cv[n++] = Hjj[r * dj + c]; }
        for (int r = 0; r < di; r++) b[oi + r] += bi[r];
        for (int r = 0; r < dj; r++) b[oj + r] += bj[r];
        if (need_prior && (g->ekind[k] == OG_E_SE2 || g->ekind[k] == OG_E_SE3)) {
            for (int r = 0; r < di; r++) { ci[n] = (int32_t)(oi + r); cj[n] = (int32_t)(oi + r); cv[n++] = 10000000.0; }
            need_prior = 0;
        }
    }
    for (int64_t i = 0; i < g->len; i++) b[i] = -b[i];
    if (lm) for (int64_t i = 0; i < g->len; i++) { ci[n] = (int32_t)i; cj[n] = (int32_t)i; cv[n++] = lambda; }
    return n;
}

/* COO -> CSC with duplicate summation and sorted row indices inside each column: what
 * russell_sparse 0.7.1 does before handing the matrix to UMFPACK (third-party).
 * Duplicates are summed in COO (= put) order.  Returns nnz; col_ptr has n+1 entries;
 * row_idx/vals need capacity nnz_coo. */
int64_t og_coo_to_csc(int64_t n, int64_t nnz_coo, const int32_t *ci, const int32_t *cj, const double *cv,
                      int32_t *col_ptr, int32_t *row_idx, double *vals) {
    int64_t *cnt = calloc(n + 1, sizeof(int64_t));
    for (int64_t k = 0; k < nnz_coo; k++) cnt[cj[k] + 1]++;
    for (int64_t c = 0; c < n; c++) cnt[c + 1] += cnt[c];
    int64_t *pos = malloc((n + 1) * sizeof(int64_t));
    memcpy(pos, cnt, (n + 1) * sizeof(int64_t));
    int32_t *ri = malloc(nnz_coo * sizeof(int32_t)); double *rv = malloc(nnz_coo * sizeof(double));
    for (int64_t k = 0; k < nnz_coo; k++) { int64_t p = pos[cj[k]]++; ri[p] = ci[k]; rv[p] = cv[k]; }   /* stable */
    /* per column: stable insertion sort by row (columns are short), then merge duplicates */
    int64_t out = 0;
    col_ptr[0] = 0;
    for (int64_t c = 0; c < n; c++) {
        int64_t lo = cnt[c], hi = cnt[c + 1];
        for (int64_t a = lo + 1; a < hi; a++) {
            int32_t r = ri[a]; double v = rv[a]; int64_t q = a - 1;
            while (q >= lo && ri[q] > r) { ri[q + 1] = ri[q]; rv[q + 1] = rv[q]; q--; }
            ri[q + 1] = r; rv[q + 1] = v;
        }
        for (int64_t a = lo; a < hi; a++) {
            if (a > lo && ri[a] == ri[a - 1]) vals[out - 1] += rv[a];
            else { row_idx[out] = ri[a]; vals[out] = rv[a]; out++; }
        }
        col_ptr[c + 1] = (int32_t)out;
    }
    free(cnt); free(pos); free(ri); free(rv);
    return out;
}

/* update_nodes, :229-245 : SE2: t += dx.xy (global frame), R <- R * R(dx.z) as a unit-complex
 * product without renormalisation (:235-236) ; XY: l += dx (:239).
 * SE3 (repo-defined): t += dx[0..3], q <- q * Exp(dx[3..6]). */
void og_update_nodes(og_graph *g, const double *dx, double sign) {
    for (int64_t i = 0; i < g->n_vertices; i++) {
        double *s = g->vstate + VS * i; const double *d = dx + g->voffset[i];
        if (g->vkind[i] == OG_SE2) {
            s[0] += sign * d[0]; s[1] += sign * d[1];
            double c = cos(sign * d[2]), sn = sin(sign * d[2]);
            double re = s[2] * c - s[3] * sn, im = s[2] * sn + s[3] * c;
            s[2] = re; s[3] = im;
        } else if (g->vkind[i] == OG_XY) { s[0] += sign * d[0]; s[1] += sign * d[1]; }
        else {
            s[0] += sign * d[0]; s[1] += sign * d[1]; s[2] += sign * d[2];
            double w[3] = {sign * d[3], sign * d[4], sign * d[5]}, dq[4], q[4];
            q_exp(w, dq); q_mul(s + 3, dq, q);
            double n = sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]);
            for (int c = 0; c < 4; c++) s[3 + c] = q[c] / n;
        }
    }
}
