/* oracle/cd_statebased.c -- scalar float64 C restatement of state-based conflict detection.
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline); PARITY UNPINNED (see oracle/__init__.py).
 *
 * [UPSTREAM-RECALL] follows bluesky/traffic/asas/statebased.py::StateBased.detect operation by
 * operation (the same order of float64 operations as oracle/statebased.py, one ordered pair at a
 * time instead of dense N x N temporaries), so it can check sampled rows at N = 100k where the
 * NumPy restatement would need ~1.2 TB.  The reference never enables ASAS (only `reso off`,
 * merge_env.py:157); BASELINE.json's north_star adds the detection.
 *
 * Build: make -C oracle   (gcc -O2 -pthread -shared -fPIC; no -ffast-math: keep IEEE semantics).
 * Rows are spread over `nthreads` POSIX threads (libgomp is not linkable in this image).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

#define NM 1852.0
#define RE 6371000.0
#define DEG2RAD 0.017453292519943295
#define RAD2DEG 57.29577951308232

typedef struct {
    int swconfl, swlos;
    double tcpa, dcpa2, tinconf, toutconf, dist, qdr;
} cd_pair;

static double pymod(double a, double m) {        /* numpy's % : result has the sign of m */
    double r = fmod(a, m);
    if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
    return r;
}

static void pair_eval(const double *lat, const double *lon, const double *u, const double *v,
                      const double *alt, const double *vs, const double *rpz, const double *hpz,
                      const double *dtl, long i, long j, cd_pair *o) {
    double I = (i == j) ? 1.0 : 0.0;
    /* geo.kwikqdrdist_matrix */
    double dlat = DEG2RAD * (lat[j] - lat[i]);
    double dlon = DEG2RAD * (pymod((lon[j] - lon[i]) + 180.0, 360.0) - 180.0);
    double cavelat = cos(DEG2RAD * (lat[i] + lat[j]) * 0.5);
    double dangle = sqrt(dlat * dlat + dlon * dlon * cavelat * cavelat);
    double distnm = RE * dangle / NM;
    double qdr = pymod(RAD2DEG * atan2(dlon * cavelat, dlat), 360.0);
    double dist = distnm * NM + 1e9 * I;
    double qr = DEG2RAD * qdr;
    double dx = dist * sin(qr), dy = dist * cos(qr);
    double du = u[j] - u[i], dv = v[j] - v[i];
    double dv2 = du * du + dv * dv;
    if (fabs(dv2) < 1e-6) dv2 = 1e-6;
    double vrel = sqrt(dv2);
    double tcpa = -(du * dx + dv * dy) / dv2 + 1e9 * I;
    double dcpa2 = fabs(dist * dist - tcpa * tcpa * dv2);
    double rp = fmax(rpz[i], rpz[j]);
    double R2 = rp * rp;
    int swhor = dcpa2 < R2;
    double dxinhor = sqrt(fmax(0.0, R2 - dcpa2));
    double dtinhor = dxinhor / vrel;
    double tinhor = swhor ? tcpa - dtinhor : 1e8;
    double touthor = swhor ? tcpa + dtinhor : -1e8;
    double dalt = alt[j] - alt[i] + 1e9 * I;
    double dvs = vs[j] - vs[i];
    if (fabs(dvs) < 1e-6) dvs = 1e-6;
    double hp = fmax(hpz[i], hpz[j]);
    double thi = (dalt + hp) / -dvs, tlo = (dalt - hp) / -dvs;
    double tinver = fmin(thi, tlo), toutver = fmax(thi, tlo);
    double tinconf = fmax(tinver, tinhor), toutconf = fmin(toutver, touthor);
    o->swconfl = swhor && (tinconf <= toutconf) && (toutconf > 0.0) && (tinconf < dtl[i]) && (i != j);
    o->swlos = (dist < rp) && (fabs(dalt) < hp);
    o->tcpa = tcpa; o->dcpa2 = dcpa2; o->tinconf = tinconf; o->toutconf = toutconf;
    o->dist = dist; o->qdr = qdr;
}

typedef struct {
    const double *lat, *lon, *u, *v, *alt, *vs, *rpz, *hpz, *dtl;
    long n, row0, nrows;
    int tid, nthreads;
    uint8_t *inconf; double *tcpamax; uint32_t *nconf_row, *nlos_row;
    uint64_t tc, tl;
} cd_job;

static void *cd_worker(void *arg) {
    cd_job *w = (cd_job *)arg;
    uint64_t tc = 0, tl = 0;
    for (long r = w->tid; r < w->nrows; r += w->nthreads) {
        long i = w->row0 + r;
        uint32_t nc = 0, nl = 0;
        double tmax = 0.0;
        for (long j = 0; j < w->n; ++j) {
            cd_pair p;
            pair_eval(w->lat, w->lon, w->u, w->v, w->alt, w->vs, w->rpz, w->hpz, w->dtl, i, j, &p);
            if (p.swconfl) { nc++; if (p.tcpa > tmax) tmax = p.tcpa; }
            if (p.swlos) nl++;
        }
        w->inconf[r] = nc > 0;
        w->tcpamax[r] = tmax;       /* max over the row of tcpa*swconfl: never below the zeros */
        w->nconf_row[r] = nc;
        w->nlos_row[r] = nl;
        tc += nc; tl += nl;
    }
    w->tc = tc; w->tl = tl;
    return 0;
}

/* Rows [row0, row0+nrows) against all n columns.  Per-row outputs are indexed by (i - row0).
 * pairs_conf / pairs_los (nullable) receive (i, j) int32 pairs in row-major order up to cap; the
 * true totals are returned in totals[0] (conflicts) and totals[1] (LoS).  Returns 0. */
int cd_detect_rows(const double *lat, const double *lon, const double *trk, const double *gs,
                   const double *alt, const double *vs, const double *rpz, const double *hpz,
                   const double *dtl, long n, long row0, long nrows,
                   uint8_t *inconf, double *tcpamax, uint32_t *nconf_row, uint32_t *nlos_row,
                   int32_t *pairs_conf, int32_t *pairs_los, long cap, uint64_t *totals, int nthreads) {
    double *u = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *v = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (!u || !v) { free(u); free(v); return -1; }
    for (long k = 0; k < n; ++k) {
        double tr = DEG2RAD * trk[k];
        u[k] = gs[k] * sin(tr);
        v[k] = gs[k] * cos(tr);
    }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    cd_job jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) {
        cd_job j = {lat, lon, u, v, alt, vs, rpz, hpz, dtl, n, row0, nrows, t, nthreads,
                    inconf, tcpamax, nconf_row, nlos_row, 0, 0};
        jobs[t] = j;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], 0, cd_worker, &jobs[t]);
    cd_worker(&jobs[0]);
    uint64_t tc = jobs[0].tc, tl = jobs[0].tl;
    for (int t = 1; t < nthreads; ++t) { pthread_join(th[t], 0); tc += jobs[t].tc; tl += jobs[t].tl; }
    totals[0] = tc; totals[1] = tl;
    if (pairs_conf || pairs_los) {  /* second, serial pass keeps row-major order */
        long kc = 0, kl = 0;
        for (long r = 0; r < nrows; ++r) {
            if (!nconf_row[r] && !nlos_row[r]) continue;
            long i = row0 + r;
            for (long j = 0; j < n; ++j) {
                cd_pair p;
                pair_eval(lat, lon, u, v, alt, vs, rpz, hpz, dtl, i, j, &p);
                if (p.swconfl && pairs_conf && kc < cap) { pairs_conf[2 * kc] = (int32_t)i; pairs_conf[2 * kc + 1] = (int32_t)j; kc++; }
                if (p.swlos && pairs_los && kl < cap) { pairs_los[2 * kl] = (int32_t)i; pairs_los[2 * kl + 1] = (int32_t)j; kl++; }
            }
        }
    }
    free(u); free(v);
    return 0;
}

/* One ordered pair, all intermediate quantities (for band-exemption logic in the tests). */
int cd_pair_eval(const double *lat, const double *lon, const double *trk, const double *gs,
                 const double *alt, const double *vs, const double *rpz, const double *hpz,
                 const double *dtl, long i, long j, double *out8) {
    double u[2], v[2], la[2] = {lat[i], lat[j]}, lo[2] = {lon[i], lon[j]};
    double al[2] = {alt[i], alt[j]}, vv[2] = {vs[i], vs[j]}, rp[2] = {rpz[i], rpz[j]};
    double hp[2] = {hpz[i], hpz[j]}, dl[2] = {dtl[i], dtl[j]};
    long idx[2] = {i, j};
    for (int k = 0; k < 2; ++k) {
        double tr = DEG2RAD * trk[idx[k]];
        u[k] = gs[idx[k]] * sin(tr); v[k] = gs[idx[k]] * cos(tr);
    }
    cd_pair p;
    pair_eval(la, lo, u, v, al, vv, rp, hp, dl, 0, (i == j) ? 0 : 1, &p);
    out8[0] = p.swconfl; out8[1] = p.swlos; out8[2] = p.tcpa; out8[3] = p.dcpa2;
    out8[4] = p.tinconf; out8[5] = p.toutconf; out8[6] = p.dist; out8[7] = p.qdr;
    return 0;
}
