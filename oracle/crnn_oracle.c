/* CPU ORACLE (test infrastructure, NOT product code) -- C restatement of the reference's surrogate
 * ODE path, used by tests/ as the parity checker and by bench.py's cpu_baseline leg.  Nothing in the
 * product package links, loads or calls this file.
 *
 * Reference being restated (paths relative to /root/reference):
 *   CRNNFunc.forward + linear_interpolation  SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:106-155
 *                                             SURROGATE_MODEL/surrogate_model_Eon_single_model.py:76-151
 *   odeint(dopri5, atol, rtol) + clamp        ...Eoff_single_model.py:185-186, ...Eon_single_model.py:153-156
 * The dopri5 arithmetic itself is torchdiffeq's (un-vendored, version unpinned) -> PARITY UNPINNED to
 * the bit; see oracle/reference_path.py.  oracle_truth_* is not a restatement of anything: it is the
 * converged float64 solution of the same ODE (knot-to-knot, long double clock) used as ground truth.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -shared -fPIC -pthread)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* tiny pthread parallel-for (this image's gcc has no libgomp): body(n, arg) for n in [0,N), dynamic chunks */
typedef void (*pf_body)(int n, void *arg);
typedef struct { pf_body body; void *arg; int N; int next; pthread_mutex_t mu; } pf_state;
static void *pf_worker(void *vp) {
    pf_state *s = (pf_state *)vp;
    for (;;) {
        pthread_mutex_lock(&s->mu);
        int b = s->next; s->next += 4;
        pthread_mutex_unlock(&s->mu);
        if (b >= s->N) break;
        int e = b + 4 < s->N ? b + 4 : s->N;
        for (int n = b; n < e; n++) s->body(n, s->arg);
    }
    return NULL;
}
static void parallel_for(int N, int nthreads, pf_body body, void *arg) {
    if (nthreads <= 1 || N <= 1) { for (int n = 0; n < N; n++) body(n, arg); return; }
    if (nthreads > 256) nthreads = 256;
    pf_state s = {body, arg, N, 0, PTHREAD_MUTEX_INITIALIZER};
    pthread_t th[256];
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, pf_worker, &s);
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
}

#define NTOT 801
static double g_lb = 1.0e-6;   /* state clamp lower bound: 1e-6 inference / wide trainer, 1e-5 narrow trainers (Eon...:45) */
#define LB_ g_lb
void oracle_set_lb(double lb) { g_lb = lb; }
#define UB_ 6.0e1
/* the reference holds R_kcal in float32 (a float32 tensor in the Eoff script, a python scalar rounded to the
 * tensor dtype in the Eon script), so the model constant IS the float32-rounded value, in every precision */
#define RKCAL_ ((double)1.9872036e-3f)

/* ---- float32 state (the reference's dtype) ---- */
#define REAL float
#define SUF _f32
#define MLOG logf
#define MEXP expf
#define MSQRT sqrtf
#define MABS fabsf
#define MPOW powf
#define MNEXTAFTER nextafterf
#include "dopri5_impl.inc"
#undef REAL
#undef SUF
#undef MLOG
#undef MEXP
#undef MSQRT
#undef MABS
#undef MPOW
#undef MNEXTAFTER

/* ---- float64 state (what torchdiffeq does when handed double tensors) ---- */
#define REAL double
#define SUF _f64
#define MLOG log
#define MEXP exp
#define MSQRT sqrt
#define MABS fabs
#define MPOW pow
#define MNEXTAFTER nextafter
#include "dopri5_impl.inc"
#undef REAL
#undef SUF

/* Batched driver.  tgrid/Tprof [N][801] float32 row-major, u0 [N][9] float32, parameters float32.
 * precision: 32 | 64.  report[N]: knot index whose (clamped) state goes to y_out[N][9] (double).
 * sol (optional) [N][801][9] double, unclamped dense output.  stats [N][4] = nfe, accepted, rejected, status. */
typedef struct {
    const float *tgrid, *Tprof, *u0, *w_in, *w_b, *w_out;
    double rtol, atol, inter_lo, inter_hi;
    int precision; const int *report; double *y_out, *sol; int *stats; int *bad;
} d5_args;

static void d5_body(int n, void *vp) {
    d5_args *a = (d5_args *)vp;
    int st[4];
    int rep = a->report ? a->report[n] : NTOT - 1;
    if (a->precision == 32) {
        rhs_ctx_f32 c;
        c.ts = a->tgrid + (size_t)n * NTOT; c.Ts = a->Tprof + (size_t)n * NTOT;
        memcpy(c.w_in, a->w_in, sizeof(c.w_in)); memcpy(c.w_b, a->w_b, sizeof(c.w_b)); memcpy(c.w_out, a->w_out, sizeof(c.w_out));
        c.inter_lo = (float)a->inter_lo; c.inter_hi = (float)a->inter_hi;
        float y[9], *s = a->sol ? (float *)malloc(sizeof(float) * NTOT * 9) : NULL;
        dopri5_one_f32(&c, a->u0 + (size_t)n * 9, a->rtol, a->atol, NTOT, s, rep, y, st);
        for (int i = 0; i < 9; i++) a->y_out[(size_t)n * 9 + i] = y[i];
        if (s) { for (int i = 0; i < NTOT * 9; i++) a->sol[(size_t)n * NTOT * 9 + i] = s[i]; free(s); }
    } else {
        rhs_ctx_f64 c;
        double ts[NTOT], Ts[NTOT], u[9];
        for (int i = 0; i < NTOT; i++) { ts[i] = a->tgrid[(size_t)n * NTOT + i]; Ts[i] = a->Tprof[(size_t)n * NTOT + i]; }
        for (int i = 0; i < 9; i++) u[i] = a->u0[(size_t)n * 9 + i];
        c.ts = ts; c.Ts = Ts;
        for (int i = 0; i < 99; i++) c.w_in[i] = a->w_in[i];
        for (int i = 0; i < 9; i++) c.w_b[i] = a->w_b[i];
        for (int i = 0; i < 81; i++) c.w_out[i] = a->w_out[i];
        c.inter_lo = a->inter_lo; c.inter_hi = a->inter_hi;
        dopri5_one_f64(&c, u, a->rtol, a->atol, NTOT, a->sol ? a->sol + (size_t)n * NTOT * 9 : NULL, rep, a->y_out + (size_t)n * 9, st);
    }
    if (a->stats) memcpy(a->stats + (size_t)n * 4, st, sizeof(st));
    if (st[3] != 0) __sync_fetch_and_add(a->bad, 1);
}

int oracle_dopri5_batch(int N, const float *tgrid, const float *Tprof, const float *u0, const float *w_in,
                        const float *w_b, const float *w_out, double rtol, double atol, double inter_lo,
                        double inter_hi, int precision, const int *report, double *y_out, double *sol, int *stats,
                        int nthreads) {
    int bad = 0;
    d5_args a = {tgrid, Tprof, u0, w_in, w_b, w_out, rtol, atol, inter_lo, inter_hi, precision, report, y_out, sol, stats, &bad};
    parallel_for(N, nthreads, d5_body, &a);
    return bad;
}

/* ------------------------------------------------------------------------------------------------
 * Converged truth: integrate knot to knot (T(t) is linear, hence the RHS smooth, inside each interval)
 * with an error-controlled Dormand-Prince 5(4) in double at a tolerance far below any parity bar.
 * ------------------------------------------------------------------------------------------------ */
typedef struct { double w_in[99], w_b[9], w_out[81], lo, hi; } truth_par;

static void truth_rhs(const truth_par *p, double Tq, const double *u, double *du) {
    double wv[11], r[9];
    for (int k = 0; k < 9; k++) { double Y = u[k] < LB_ ? LB_ : (u[k] > UB_ ? UB_ : u[k]); wv[k] = log(Y); }
    wv[9] = -1.0 / (RKCAL_ * Tq); wv[10] = log(Tq);
    for (int j = 0; j < 9; j++) {
        double z = p->w_b[j];
        for (int k = 0; k < 11; k++) z += p->w_in[k * 9 + j] * wv[k];
        z = z < p->lo ? p->lo : (z > p->hi ? p->hi : z);
        r[j] = exp(z);
    }
    for (int i = 0; i < 9; i++) {
        double s = 0;
        for (int j = 0; j < 9; j++) s += p->w_out[i * 9 + j] * r[j];
        du[i] = s < -1e5 ? -1e5 : (s > 1e5 ? 1e5 : s);
    }
}

static int truth_interval(const truth_par *p, double ta, double tb, double Ta, double slope, double *y, double rtol, double atol) {
    static const double A[6] = {1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
    static const double B[6][6] = {
        {1.0 / 5}, {3.0 / 40, 9.0 / 40}, {44.0 / 45, -56.0 / 15, 32.0 / 9},
        {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729},
        {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656},
        {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
    static const double CE[7] = {35.0 / 384 - 1951.0 / 21600, 0, 500.0 / 1113 - 22642.0 / 50085, 125.0 / 192 - 451.0 / 720,
                                 -2187.0 / 6784 - -12231.0 / 42400, 11.0 / 84 - 649.0 / 6300, -1.0 / 60.0};
    double k[7][9], yi[9];
    double t = ta, h = tb - ta;
    int guard = 0;
    truth_rhs(p, Ta + slope * (t - ta), y, k[0]);
    while (t < tb) {
        if (++guard > 2000000) return 1;
        if (t + h > tb) h = tb - t;
        for (int j = 0; j < 6; j++) {
            for (int i = 0; i < 9; i++) {
                double s = 0;
                for (int m = 0; m <= j; m++) s += B[j][m] * k[m][i];
                yi[i] = y[i] + h * s;
            }
            truth_rhs(p, Ta + slope * (t + A[j] * h - ta), yi, k[j + 1]);
        }
        double e2 = 0;
        for (int i = 0; i < 9; i++) {
            double s = 0;
            for (int m = 0; m < 7; m++) s += CE[m] * k[m][i];
            double tol = atol + rtol * fmax(fabs(y[i]), fabs(yi[i]));
            e2 += (h * s / tol) * (h * s / tol);
        }
        double ratio = sqrt(e2 / 9);
        if (ratio <= 1.0) {
            t = (t + h >= tb) ? tb : t + h;
            for (int i = 0; i < 9; i++) { y[i] = yi[i]; k[0][i] = k[6][i]; }
        }
        double fac = ratio > 0 ? 0.9 / pow(ratio, 0.2) : 5.0;
        fac = fac < 0.2 ? 0.2 : (fac > 5.0 ? 5.0 : fac);
        h *= fac;
        if (h < 1e-18) return 2;
    }
    return 0;
}

typedef struct {
    const float *tgrid, *Tprof, *u0; const truth_par *p; const int *upto; double rtol, atol; double *y_out, *y_knots; int *bad;
} tr_args;

static void tr_body(int n, void *vp) {
    tr_args *a = (tr_args *)vp;
    const float *ts = a->tgrid + (size_t)n * NTOT, *Ts = a->Tprof + (size_t)n * NTOT;
    double y[9];
    for (int i = 0; i < 9; i++) y[i] = a->u0[(size_t)n * 9 + i];
    int last = a->upto ? a->upto[n] : NTOT - 1;
    if (a->y_knots) for (int i = 0; i < 9; i++) a->y_knots[((size_t)n * NTOT) * 9 + i] = y[i];
    for (int kk = 0; kk < last; kk++) {
        double ta = ts[kk], tb = ts[kk + 1], Ta = Ts[kk], Tb = Ts[kk + 1];
        double slope = (Tb - Ta) / (tb - ta);
        if (truth_interval(a->p, ta, tb, Ta, slope, y, a->rtol, a->atol) != 0) __sync_fetch_and_add(a->bad, 1);
        if (a->y_knots) for (int i = 0; i < 9; i++) a->y_knots[((size_t)n * NTOT + kk + 1) * 9 + i] = y[i];
    }
    for (int i = 0; i < 9; i++) a->y_out[(size_t)n * 9 + i] = y[i];
}

/* y_knots (optional) [N][801][9]: state at every knot up to upto[n]; y_out [N][9]: state at knot upto[n] (unclamped) */
int oracle_truth_batch(int N, const float *tgrid, const float *Tprof, const float *u0, const float *w_in,
                       const float *w_b, const float *w_out, double inter_lo, double inter_hi, const int *upto,
                       double rtol, double atol, double *y_out, double *y_knots, int nthreads) {
    truth_par p;
    for (int i = 0; i < 99; i++) p.w_in[i] = w_in[i];
    for (int i = 0; i < 9; i++) p.w_b[i] = w_b[i];
    for (int i = 0; i < 81; i++) p.w_out[i] = w_out[i];
    p.lo = inter_lo; p.hi = inter_hi;
    int bad = 0;
    tr_args a = {tgrid, Tprof, u0, &p, upto, rtol, atol, y_out, y_knots, &bad};
    parallel_for(N, nthreads, tr_body, &a);
    return bad;
}

/* Same as oracle_truth_batch with DOUBLE parameters (finite-difference checks of the training gradient) */
int oracle_truth_batch_dp(int N, const float *tgrid, const float *Tprof, const float *u0, const double *w_in,
                          const double *w_b, const double *w_out, double inter_lo, double inter_hi, const int *upto,
                          double rtol, double atol, double *y_out, double *y_knots, int nthreads) {
    truth_par p;
    for (int i = 0; i < 99; i++) p.w_in[i] = w_in[i];
    for (int i = 0; i < 9; i++) p.w_b[i] = w_b[i];
    for (int i = 0; i < 81; i++) p.w_out[i] = w_out[i];
    p.lo = inter_lo; p.hi = inter_hi;
    int bad = 0;
    tr_args a = {tgrid, Tprof, u0, &p, upto, rtol, atol, y_out, y_knots, &bad};
    parallel_for(N, nthreads, tr_body, &a);
    return bad;
}

/* Batched RHS in double at given temperatures (a7 without the knot lookup): du[N][9] */
void oracle_rhs_batch(int N, const double *T, const double *u, const float *w_in, const float *w_b, const float *w_out,
                      double inter_lo, double inter_hi, double *du) {
    truth_par p;
    for (int i = 0; i < 99; i++) p.w_in[i] = w_in[i];
    for (int i = 0; i < 9; i++) p.w_b[i] = w_b[i];
    for (int i = 0; i < 81; i++) p.w_out[i] = w_out[i];
    p.lo = inter_lo; p.hi = inter_hi;
    for (int n = 0; n < N; n++) truth_rhs(&p, T[n], u + (size_t)n * 9, du + (size_t)n * 9);
}
