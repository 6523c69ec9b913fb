/*
 * Deterministic synthetic pose graphs for the benchmark configs of BASELINE.json
 * (the reference ships no generator -- these definitions are this repo's).
 *
 *  - pgo_synth_manhattan_se2: "Manhattan world" SE(2) graph (C3: 100k poses / 400k edges,
 *    C4: 1M poses / 4M edges).
 *  - pgo_synth_sphere_se3: SE(3) sphere-spiral graph (C5: 250k poses / ~1M edges).
 *
 * Plain C, no dependencies; counter-based splitmix64 + Box-Muller so that every number is a
 * pure function of (seed, stream, index).  Output arrays use the same packing as the C ABI
 * (include/pgo_b200.h): SE2 vertex = x y theta ; EDGE_SE2 = dx dy dtheta + 6 upper-triangular
 * information entries.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline double uni(uint64_t seed, uint64_t stream, uint64_t idx) { /* (0,1) */
    uint64_t h = mix64(mix64(seed ^ (stream * 0xD1B54A32D192ED03ull)) + idx);
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static inline double gauss(uint64_t seed, uint64_t stream, uint64_t idx) {
    double u1 = uni(seed, stream, 2 * idx), u2 = uni(seed, stream, 2 * idx + 1);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
}
static inline double wrap_pi(double a) {
    while (a > M_PI) a -= 2.0 * M_PI;
    while (a <= -M_PI) a += 2.0 * M_PI;
    return a;
}

/* open-addressing hash: grid cell -> most recent pose index at that cell; next[] chains older visits */
typedef struct { uint64_t *key; int64_t *head; uint64_t mask; } cellmap;
static inline uint64_t cell_key(int64_t x, int64_t y) { return ((uint64_t)(uint32_t)(int32_t)x << 32) | (uint32_t)(int32_t)y; }
static int64_t *cell_slot(cellmap *m, uint64_t key) {
    uint64_t h = mix64(key) & m->mask;
    while (m->key[h] != key && m->head[h] != -2) h = (h + 1) & m->mask;
    if (m->head[h] == -2) { m->key[h] = key; m->head[h] = -1; }
    return &m->head[h];
}
static int64_t cell_find(const cellmap *m, uint64_t key) {
    uint64_t h = mix64(key) & m->mask;
    while (m->head[h] != -2) { if (m->key[h] == key) return m->head[h]; h = (h + 1) & m->mask; }
    return -1;
}

/*
 * Ground truth: unit steps on the integer grid from (0,0,0); each step turns +90 / -90 degrees
 * with probability 0.15 each.  Edges, emitted pose by pose: odometry (i-1 -> i), then loop
 * closures (j -> i), j < i-10, to the most recent earlier visit of each grid cell in a
 * (2r+1)^2 window (r = 2, widened ring by ring up to r = 6 while the running quota of
 * target_edges * (i+1)/n is not met; at most 6 closures per pose; no duplicate pairs;
 * from < to always).  Measurement = exact relative pose + N(0, 0.05 m / 0.01 rad);
 * information = diag(400, 400, 10000).  Initial guess = ground truth + N(0, 0.1 m / 0.03 rad).
 * Returns the number of edges written (<= target_edges), or -1 on bad arguments.
 * vertex_values: 3n doubles; edge_*: capacity target_edges (from/to), 3x / 6x for meas / info.
 * ground_truth may be NULL.
 */
int64_t pgo_synth_manhattan_se2(int64_t n, int64_t target_edges, uint64_t seed,
                                double *vertex_values, double *ground_truth,
                                uint32_t *edge_from, uint32_t *edge_to,
                                double *edge_meas, double *edge_info_upper) {
    if (n < 2 || target_edges < n - 1 || n > 0x7fffffffll) return -1;
    int64_t *gx = malloc(n * sizeof(int64_t)), *gy = malloc(n * sizeof(int64_t)), *next = malloc(n * sizeof(int64_t));
    int8_t *gh = malloc(n);
    cellmap m; uint64_t cap = 16; while (cap < (uint64_t)(2 * n)) cap <<= 1;
    m.mask = cap - 1; m.key = malloc(cap * sizeof(uint64_t)); m.head = malloc(cap * sizeof(int64_t));
    for (uint64_t i = 0; i < cap; i++) m.head[i] = -2;
    static const int DX[4] = {1, 0, -1, 0}, DY[4] = {0, 1, 0, -1};
    int64_t x = 0, y = 0; int h = 0, ne = 0;
    const int64_t closure_target = target_edges - (n - 1);
    int64_t closures = 0;
    for (int64_t i = 0; i < n; i++) {
        if (i > 0) {
            double u = uni(seed, 1, (uint64_t)i);
            if (u < 0.15) h = (h + 1) & 3; else if (u < 0.30) h = (h + 3) & 3;
            x += DX[h]; y += DY[h];
        }
        gx[i] = x; gy[i] = y; gh[i] = (int8_t)h;
        int64_t cand[8]; int nc = 0;
        if (i > 0) cand[nc++] = i - 1;                              /* odometry first */
        /* closures */
        int64_t quota = (int64_t)((long double)closure_target * (long double)(i + 1) / (long double)n) - closures;
        if (quota > 6) quota = 6;
        int ncl = 0;
        for (int r = 0; r <= 6 && ncl < quota; r++) {
            if (r > 2 && ncl >= quota) break;
            for (int64_t cy = y - r; cy <= y + r && ncl < quota; cy++)
                for (int64_t cx = x - r; cx <= x + r && ncl < quota; cx++) {
                    int64_t ax = cx > x ? cx - x : x - cx, ay = cy > y ? cy - y : y - cy;
                    if ((ax > ay ? ax : ay) != r) continue;          /* ring r only */
                    int64_t j = cell_find(&m, cell_key(cx, cy));
                    while (j >= 0 && j >= i - 10) j = next[j];       /* need j < i - 10 */
                    if (j < 0) continue;
                    cand[nc++] = j; ncl++;
                }
        }
        closures += ncl;
        /* closures sorted by j ascending after the odometry edge */
        for (int a = (i > 0 ? 2 : 1); a < nc; a++) { int64_t v = cand[a]; int b = a - 1;
            while (b >= (i > 0 ? 1 : 0) && cand[b] > v) { cand[b + 1] = cand[b]; b--; } cand[b + 1] = v; }
        for (int c = 0; c < nc; c++) {
            int64_t j = cand[c];
            double thj = gh[j] * (M_PI / 2), cj = cos(thj), sj = sin(thj);
            double dx = (double)(x - gx[j]), dy = (double)(y - gy[j]);
            double *z = edge_meas + 3 * ne, *w = edge_info_upper + 6 * ne;
            z[0] = (cj * dx + sj * dy) + 0.05 * gauss(seed, 2, 3 * (uint64_t)ne + 0);
            z[1] = (-sj * dx + cj * dy) + 0.05 * gauss(seed, 2, 3 * (uint64_t)ne + 1);
            z[2] = wrap_pi((gh[i] - gh[j]) * (M_PI / 2) + 0.01 * gauss(seed, 2, 3 * (uint64_t)ne + 2));
            w[0] = 400.0; w[1] = 0.0; w[2] = 0.0; w[3] = 400.0; w[4] = 0.0; w[5] = 10000.0;
            edge_from[ne] = (uint32_t)j; edge_to[ne] = (uint32_t)i; ne++;
        }
        int64_t *slot = cell_slot(&m, cell_key(x, y));
        next[i] = *slot; *slot = i;
        double th = gh[i] * (M_PI / 2);
        if (ground_truth) { ground_truth[3 * i] = (double)x; ground_truth[3 * i + 1] = (double)y; ground_truth[3 * i + 2] = wrap_pi(th); }
        vertex_values[3 * i + 0] = (double)x + 0.1 * gauss(seed, 3, 3 * (uint64_t)i + 0);
        vertex_values[3 * i + 1] = (double)y + 0.1 * gauss(seed, 3, 3 * (uint64_t)i + 1);
        vertex_values[3 * i + 2] = wrap_pi(th + 0.03 * gauss(seed, 3, 3 * (uint64_t)i + 2));
    }
    free(gx); free(gy); free(gh); free(next); free(m.key); free(m.head);
    return ne;
}

/* ------------------------------------------------------------------ SE(3) sphere */
static void quat_mul(const double *a, const double *b, double *o) { /* (w,x,y,z) */
    o[0] = a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3];
    o[1] = a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2];
    o[2] = a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1];
    o[3] = a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0];
}
static void quat_rot(const double *q, const double *v, double *o) { /* R(q) v */
    double w = q[0], x = q[1], y = q[2], z = q[3];
    double tx = 2 * (y * v[2] - z * v[1]), ty = 2 * (z * v[0] - x * v[2]), tz = 2 * (x * v[1] - y * v[0]);
    o[0] = v[0] + w * tx + (y * tz - z * ty);
    o[1] = v[1] + w * ty + (z * tx - x * tz);
    o[2] = v[2] + w * tz + (x * ty - y * tx);
}
static void quat_exp(const double *p, double *q) {
    double th = sqrt(p[0]*p[0] + p[1]*p[1] + p[2]*p[2]);
    double k = th < 1e-12 ? 0.5 : sin(0.5 * th) / th;
    q[0] = cos(0.5 * th); q[1] = k * p[0]; q[2] = k * p[1]; q[3] = k * p[2];
}
static void mat_to_quat(const double R[9], double *q) {
    double tr = R[0] + R[4] + R[8];
    if (tr > 0) { double s = sqrt(tr + 1.0) * 2; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
    else if (R[0] > R[4] && R[0] > R[8]) { double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
    else if (R[4] > R[8]) { double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s; }
    else { double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s; }
}

/*
 * Sphere spiral (as sphere2500 = 50 levels x 50 poses): `levels` x `per_level` poses on a
 * radius-`radius` sphere; edges i -> i+1, and i -> i+per_level-1, i+per_level, i+per_level+1
 * (to the next level) where in range.  Information = diag(100,100,100,400,400,400)
 * (as torus3D.g2o); noise sigma_t = 0.1, sigma_r = 0.05 on measurements; initial guess = ground
 * truth + N(0, 0.2 m / 0.05 rad).  Vertex / measurement = x y z qx qy qz qw (g2o order).
 * Returns edges written; capacities: 7n vertex values, 4n edges (7x meas, 21x info).
 */
int64_t pgo_synth_sphere_se3(int64_t levels, int64_t per_level, double radius, uint64_t seed,
                             double *vertex_values, double *ground_truth,
                             uint32_t *edge_from, uint32_t *edge_to,
                             double *edge_meas, double *edge_info_upper) {
    int64_t n = levels * per_level;
    if (levels < 2 || per_level < 4 || n > 0x7fffffffll) return -1;
    double *gt = malloc(7 * n * sizeof(double)); /* x y z qw qx qy qz */
    for (int64_t i = 0; i < n; i++) {
        double frac = ((double)i + 0.5) / (double)n;
        double polar = M_PI * (0.05 + 0.9 * frac);                      /* avoid the poles */
        double az = 2.0 * M_PI * (double)i / (double)per_level;
        double sp = sin(polar), cp = cos(polar), sa = sin(az), ca = cos(az);
        double *g = gt + 7 * i;
        g[0] = radius * sp * ca; g[1] = radius * sp * sa; g[2] = radius * cp;
        /* body frame: x = direction of travel (azimuth tangent), z = outward normal */
        double ex[3] = {-sa, ca, 0}, ez[3] = {sp * ca, sp * sa, cp};
        double ey[3] = {ez[1]*ex[2] - ez[2]*ex[1], ez[2]*ex[0] - ez[0]*ex[2], ez[0]*ex[1] - ez[1]*ex[0]};
        double R[9] = {ex[0], ey[0], ez[0], ex[1], ey[1], ez[1], ex[2], ey[2], ez[2]};
        mat_to_quat(R, g + 3);
    }
    int64_t ne = 0;
    const int64_t off[4] = {1, per_level - 1, per_level, per_level + 1};
    for (int64_t i = 0; i < n; i++) {
        for (int c = 0; c < 4; c++) {
            int64_t j = i + off[c];
            if (j >= n) continue;
            const double *a = gt + 7 * i, *b = gt + 7 * j;
            double qai[4] = {a[3], -a[4], -a[5], -a[6]}, d[3] = {b[0]-a[0], b[1]-a[1], b[2]-a[2]}, t[3], q[4];
            quat_rot(qai, d, t);
            quat_mul(qai, b + 3, q);
            double w[3], dq[4], qn[4];
            for (int k = 0; k < 3; k++) { t[k] += 0.1 * gauss(seed, 5, 6 * (uint64_t)ne + k); w[k] = 0.05 * gauss(seed, 5, 6 * (uint64_t)ne + 3 + k); }
            quat_exp(w, dq); quat_mul(q, dq, qn);
            double *z = edge_meas + 7 * ne, *W = edge_info_upper + 21 * ne;
            z[0] = t[0]; z[1] = t[1]; z[2] = t[2]; z[3] = qn[1]; z[4] = qn[2]; z[5] = qn[3]; z[6] = qn[0];
            memset(W, 0, 21 * sizeof(double));
            W[0] = 100; W[6] = 100; W[11] = 100; W[15] = 400; W[18] = 400; W[20] = 400;
            edge_from[ne] = (uint32_t)i; edge_to[ne] = (uint32_t)j; ne++;
        }
    }
    for (int64_t i = 0; i < n; i++) {
        const double *g = gt + 7 * i; double w[3], dq[4], qn[4];
        double *v = vertex_values + 7 * i;
        for (int k = 0; k < 3; k++) { v[k] = g[k] + 0.2 * gauss(seed, 6, 6 * (uint64_t)i + k); w[k] = 0.05 * gauss(seed, 6, 6 * (uint64_t)i + 3 + k); }
        quat_exp(w, dq); quat_mul(g + 3, dq, qn);
        v[3] = qn[1]; v[4] = qn[2]; v[5] = qn[3]; v[6] = qn[0];
        if (ground_truth) { double *o = ground_truth + 7 * i; o[0] = g[0]; o[1] = g[1]; o[2] = g[2]; o[3] = g[4]; o[4] = g[5]; o[5] = g[6]; o[6] = g[3]; }
    }
    free(gt);
    return ne;
}
