/* oracle_mt.c — TEST/BENCH INFRASTRUCTURE, not a parity artefact: the Jacobi-PCG of the oracle port
 * (magnetite_oracle.c, orc_cg "port mode") on the same CSR matrix, run SPMD on T pthreads, so bench.py can
 * say what ALL host cores make of the CPU side.  The reference itself is single-threaded (no rayon/threads
 * anywhere in src/), so this is labelled "not the reference algorithm" wherever it is reported.
 * Every thread owns a contiguous row range; dot products are per-thread partials summed by every thread in
 * thread order after a barrier, so the run is deterministic for a given T (and differs from the sequential
 * port by rounding only).  No OpenMP: the image's compiler wrapper has no libgomp spec.
 *   gcc -O2 -ffp-contract=off -shared -fPIC -pthread -o _build/libmagnetite_oracle_mt.so oracle_mt.c -lm */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

int orc_mt_host_cores(void) {
    const long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

enum { SLOT_STRIDE = 8 };     /* doubles per thread in the partial table: one cache line */

typedef struct {
    uint64_t n;
    const int64_t *rowptr;
    const int32_t *col;
    const double *val, *b;
    double *x, *r, *p, *q, *dinv, *partial;
    double rel_tol;
    uint64_t max_iter;
    int T;
    pthread_barrier_t bar;
    pthread_mutex_t gate_lock;   /* start gate: workers wait here until every thread exists (go = 1) or the */
    pthread_cond_t gate_cond;    /* launch is abandoned (go = -1), so nobody ever waits on a short barrier  */
    int go;
    uint64_t iters;           /* written by thread 0 at the end */
    double rr;
} shared_t;

typedef struct { shared_t *s; int tid; } arg_t;

static double sum_slot(const shared_t *s, int slot) {
    double acc = 0.0;
    for (int t = 0; t < s->T; ++t) acc += s->partial[t * SLOT_STRIDE + slot];
    return acc;
}

static void *worker(void *varg) {
    const arg_t *a = (const arg_t *)varg;
    shared_t *s = a->s;
    const int tid = a->tid;
    pthread_mutex_lock(&s->gate_lock);
    while (s->go == 0) pthread_cond_wait(&s->gate_cond, &s->gate_lock);
    const int go = s->go;
    pthread_mutex_unlock(&s->gate_lock);
    if (go < 0) return NULL;
    const uint64_t lo = s->n * (uint64_t)tid / (uint64_t)s->T, hi = s->n * (uint64_t)(tid + 1) / (uint64_t)s->T;
    double *mine = s->partial + tid * SLOT_STRIDE;
    double bb = 0.0, rz = 0.0;
    for (uint64_t i = lo; i < hi; ++i) {
        double d = 0.0;
        for (int64_t k = s->rowptr[i]; k < s->rowptr[i + 1]; ++k) if ((uint64_t)s->col[k] == i) d = s->val[k];
        if (d == 0.0) d = 1.0;
        s->dinv[i] = 1.0 / d;
        s->x[i] = 0.0; s->r[i] = s->b[i]; s->p[i] = s->r[i] * s->dinv[i];
        bb += s->b[i] * s->b[i]; rz += s->r[i] * s->p[i];
    }
    mine[3] = bb; mine[4] = rz;
    pthread_barrier_wait(&s->bar);
    bb = sum_slot(s, 3); rz = sum_slot(s, 4);
    const double thr2 = s->rel_tol * s->rel_tol * bb;
    double rr = bb;
    uint64_t it = 0;
    while (it < s->max_iter && rr > thr2) {                 /* every thread holds the same rr: same decision */
        double pq = 0.0;
        for (uint64_t i = lo; i < hi; ++i) {
            double acc = 0.0;
            for (int64_t k = s->rowptr[i]; k < s->rowptr[i + 1]; ++k) acc += s->val[k] * s->p[s->col[k]];
            s->q[i] = acc;
            pq += s->p[i] * acc;
        }
        mine[0] = pq;
        pthread_barrier_wait(&s->bar);
        const double alpha = rz / sum_slot(s, 0);
        double rz_n = 0.0, rr_n = 0.0;
        for (uint64_t i = lo; i < hi; ++i) {
            s->x[i] += alpha * s->p[i];
            s->r[i] -= alpha * s->q[i];
            const double ri = s->r[i];
            rz_n += ri * ri * s->dinv[i];
            rr_n += ri * ri;
        }
        mine[1] = rz_n; mine[2] = rr_n;
        pthread_barrier_wait(&s->bar);
        rz_n = sum_slot(s, 1); rr = sum_slot(s, 2);
        const double beta = rz_n / rz;
        rz = rz_n;
        for (uint64_t i = lo; i < hi; ++i) s->p[i] = s->r[i] * s->dinv[i] + beta * s->p[i];
        ++it;
        pthread_barrier_wait(&s->bar);                      /* p is complete before anyone multiplies with it */
    }
    if (tid == 0) { s->iters = it; s->rr = rr; }
    return NULL;
}

/* x = A^-1 b by Jacobi-PCG until ||r||_2 <= rel_tol*||b||_2 or max_iter; CSR with int64 rowptr, int32 col.
 * threads <= 0: one per online core.  Returns 0, -1 out of memory, -2 thread creation failed. */
int orc_mt_pcg(uint64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const double *b,
               double *x, double rel_tol, uint64_t max_iter, int threads, uint64_t *iters_out, double *res_out) {
    int T = threads > 0 ? threads : orc_mt_host_cores();
    if (T > 256) T = 256;                                   /* barriers stop paying long before that */
    if ((uint64_t)T > n / 64) T = (n / 64) ? (int)(n / 64) : 1;    /* at least 64 rows per thread */
    shared_t s;
    s.n = n; s.rowptr = rowptr; s.col = col; s.val = val; s.b = b; s.x = x;
    s.rel_tol = rel_tol; s.max_iter = max_iter; s.T = T; s.iters = 0; s.rr = 0.0;
    const size_t m = n ? n : 1;
    s.r = malloc(m * sizeof(double)); s.p = malloc(m * sizeof(double)); s.q = malloc(m * sizeof(double));
    s.dinv = malloc(m * sizeof(double)); s.partial = calloc((size_t)T * SLOT_STRIDE, sizeof(double));
    pthread_t *th = malloc((size_t)T * sizeof(pthread_t));
    arg_t *args = malloc((size_t)T * sizeof(arg_t));
    int rc = 0;
    if (!s.r || !s.p || !s.q || !s.dinv || !s.partial || !th || !args) rc = -1;
    if (rc == 0 && pthread_barrier_init(&s.bar, NULL, (unsigned)T) != 0) rc = -2;
    if (rc == 0) {
        pthread_mutex_init(&s.gate_lock, NULL);
        pthread_cond_init(&s.gate_cond, NULL);
        s.go = 0;
        int started = 0;
        for (; started < T; ++started) {
            args[started].s = &s; args[started].tid = started;
            if (pthread_create(&th[started], NULL, worker, &args[started]) != 0) break;
        }
        pthread_mutex_lock(&s.gate_lock);
        s.go = (started == T) ? 1 : -1;       /* fewer threads than the barrier expects: nobody enters it */
        pthread_cond_broadcast(&s.gate_cond);
        pthread_mutex_unlock(&s.gate_lock);
        for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
        if (started == T) { *iters_out = s.iters; *res_out = sqrt(s.rr); }
        else rc = -2;
        pthread_cond_destroy(&s.gate_cond);
        pthread_mutex_destroy(&s.gate_lock);
        pthread_barrier_destroy(&s.bar);
    }
    free(s.r); free(s.p); free(s.q); free(s.dinv); free(s.partial); free(th); free(args);
    return rc;
}
