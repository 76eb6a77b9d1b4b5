/* C restatement of the remap decode.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Follows sloika/transducer.py:14-73 (`map_to_sequence`) and sloika/viterbi_helpers.pyx:12-35 (`slip_update`)
 * operation for operation in float32.  PINNED: tests/test_oracle.py checks it bit for bit against
 * tests/golden/remap_cases.npz (outputs of the reference's own code, tools/make_golden_remap.py) and
 * against oracle/remap_ref.py.  Exists so that GPU parity can be checked on batches too large for the
 * NumPy restatement, and as the CPU arm of the remap measurement.
 *
 * Input `lt` is the LOG transducer [nev][nstate] (np.log already applied by the caller for log=False, so that
 * libm-vs-device logf differences stay outside the comparison); column 0 = stay; `seq` holds state columns.
 * `slip` may be NaN: that is what the reference computes with for slip=None (see oracle/remap_ref.py).
 * Priors are float64 like util.geometric_prior's; `pscore += prior` adds in double and rounds to float.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* viterbi_helpers.pyx:12-35; n >= 3 */
void sloika_oracle_slip_update(const float *x, long n, float slip, float *from_score, int64_t *from_pos)
{
    for (long j = 0; j < n; j++) { from_score[j] = 0.0f; from_pos[j] = 0; }
    from_score[0] = from_score[1] = -1e38f;
    from_score[2] = x[0] - slip;
    for (long j = 3; j < n; j++) {
        if (from_score[j - 1] >= x[j - 2]) {
            from_pos[j] = from_pos[j - 1];
            from_score[j] = from_score[j - 1];
        } else {
            from_pos[j] = j - 2;
            from_score[j] = x[j - 2];
        }
        from_score[j] -= slip;
    }
}

/* One read.  back: caller scratch nev*npos int32.  Returns 0, or -1 on bad arguments. */
int sloika_oracle_remap(const float *lt, long nev, long nstate, const int32_t *seq, long npos, float slip,
                        const double *prior0, const double *prior1, int32_t *back, int32_t *path_out,
                        float *score_out)
{
    if (nev < 1 || npos < 3 || nstate < 2) return -1;
    float *prev = (float *)malloc(sizeof(float) * npos), *cur = (float *)malloc(sizeof(float) * npos);
    float *fs = (float *)malloc(sizeof(float) * npos);
    int64_t *fp = (int64_t *)malloc(sizeof(int64_t) * npos);
    if (!prev || !cur || !fs || !fp) { free(prev); free(cur); free(fs); free(fp); return -1; }
    for (long j = 0; j < npos; j++) {
        float p = prior0 ? (float)(0.0 + prior0[j]) : 0.0f;               /* transducer.py:42-43 */
        prev[j] = p + fmaxf(lt[seq[j]], lt[0]);                           /* :44 (np.fmax ignores NaN like fmaxf) */
    }
    for (long i = 1; i < nev; i++) {
        const float *row = lt + i * nstate;
        int32_t *bi = back + i * npos;
        for (long j = 0; j < npos; j++) {                                 /* stay :50-51, step :53-56 */
            cur[j] = prev[j] + row[0];
            bi[j] = (int32_t)j;
            if (j >= 1) {
                float step = prev[j - 1] + row[seq[j]];
                if (step > cur[j]) { cur[j] = step; bi[j] = (int32_t)(j - 1); }
            }
        }
        sloika_oracle_slip_update(prev, npos, slip, fs, fp);              /* slip :58-62 */
        for (long j = 0; j < npos; j++) {
            float from = fs[j] + row[seq[j]];
            if (!(from <= cur[j])) { bi[j] = (int32_t)fp[j]; cur[j] = from; }
        }
        float *t = prev; prev = cur; cur = t;
    }
    if (prior1)
        for (long j = 0; j < npos; j++) prev[j] = (float)((double)prev[j] + prior1[j]);   /* :66-67 */
    long best = 0;                                                         /* np.argmax: first maximum, NaN maximal */
    for (long j = 1; j < npos; j++) {
        if (isnan(prev[best])) break;
        if (isnan(prev[j]) || prev[j] > prev[best]) best = j;
    }
    *score_out = prev[best];
    long p = best;
    path_out[nev - 1] = (int32_t)p;
    for (long i = nev - 1; i >= 1; i--) { p = back[i * npos + p]; path_out[i - 1] = (int32_t)p; }
    free(prev); free(cur); free(fs); free(fp);
    return 0;
}

/* Batch: lt [T][B][nstate] (time-major like the posteriors), nev[B] <= T, seq [B][P], npos[B] <= P,
 * priors [B][P] or NULL, paths_out [B][T]. */
typedef struct {
    const float *lt; long T, B, nstate, P; const int *nev; const int32_t *seq; const int *npos; float slip;
    const double *prior0, *prior1; int32_t *paths_out; float *score_out; long next; int bad; pthread_mutex_t mu;
} remap_job;

static void *remap_worker(void *arg)
{
    remap_job *job = (remap_job *)arg;
    for (;;) {
        pthread_mutex_lock(&job->mu);
        long b = job->next++;
        pthread_mutex_unlock(&job->mu);
        if (b >= job->B) break;
        const long n = job->nev ? job->nev[b] : job->T, np_ = job->npos ? job->npos[b] : job->P;
        int rc = -1;
        if (n >= 1 && np_ >= 3) {
            float *lt = (float *)malloc(sizeof(float) * n * job->nstate);
            int32_t *back = (int32_t *)malloc(sizeof(int32_t) * n * np_);
            if (lt && back) {
                for (long t = 0; t < n; t++)
                    memcpy(lt + t * job->nstate, job->lt + (t * job->B + b) * job->nstate, sizeof(float) * job->nstate);
                rc = sloika_oracle_remap(lt, n, job->nstate, job->seq + b * job->P, np_, job->slip,
                                         job->prior0 ? job->prior0 + b * job->P : 0,
                                         job->prior1 ? job->prior1 + b * job->P : 0, back,
                                         job->paths_out + b * job->T, job->score_out + b);
            }
            free(lt); free(back);
        }
        if (rc != 0) job->bad = 1;
    }
    return 0;
}

int sloika_oracle_remap_batch(const float *lt, long T, long B, long nstate, const int *nev, const int32_t *seq, long P,
                              const int *npos, float slip, const double *prior0, const double *prior1, int nthreads,
                              int32_t *paths_out, float *score_out)
{
    remap_job job;
    job.lt = lt; job.T = T; job.B = B; job.nstate = nstate; job.P = P; job.nev = nev; job.seq = seq; job.npos = npos;
    job.slip = slip; job.prior0 = prior0; job.prior1 = prior1; job.paths_out = paths_out; job.score_out = score_out;
    job.next = 0; job.bad = 0;
    pthread_mutex_init(&job.mu, 0);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads - 1; i++)
        if (pthread_create(&th[started], 0, remap_worker, &job) == 0) started++;
    remap_worker(&job);
    for (int i = 0; i < started; i++) pthread_join(th[i], 0);
    pthread_mutex_destroy(&job.mu);
    return job.bad ? -1 : 0;
}
