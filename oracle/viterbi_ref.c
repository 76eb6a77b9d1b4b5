/* C restatement of the transducer Viterbi decode.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Follows sloika/decode.py:39-93 (`viterbi`) step for step in float32 -- the arithmetic type of the
 * real path (posteriors are float32; python-float `skip_pen` does not upcast).  PINNED: checked bit
 * for bit against oracle/decode_ref.py (itself pinned to the reference's decode.py) by
 * tests/test_oracle.py.  Exists so that full-size batches (1024 reads x 800 events) can be checked
 * exactly in seconds; reads are independent, so they are spread over OpenMP threads the same way the
 * reference spreads them over processes (here: pthreads) (sloika/iterators.py:343-351).
 *
 * Input is the log-posterior `lpost[nev][K+1]` (decode.py:56 already applied by the caller, so that
 * libm-vs-device `logf` differences stay outside the comparison), column 0 = stay.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* One read.  path_out must hold nev ints; returns path length (>= 1) or -1 on bad arguments. */
int sloika_oracle_viterbi(const float *lpost, long nev, int nbase, int klen, float skip_pen,
                          int32_t *tb /* nev*K, caller scratch */, int *path_out, float *score_out)
{
    if (nev < 1 || klen < 3 || nbase < 2) return -1;
    long K = 1;
    for (int i = 0; i < klen; i++) K *= nbase;
    const long S = K + 1;
    const int nstep = nbase, nskip = nbase * nbase;
    const long rstep = K / nstep, rskip = K / nskip;
    float *v = (float *)malloc(sizeof(float) * K);
    float *p = (float *)malloc(sizeof(float) * K);
    if (!v || !p) { free(v); free(p); return -1; }
    memcpy(v, lpost + 1, sizeof(float) * K);                       /* decode.py:57 */
    for (long i = 1; i < nev; i++) {
        const float *lp = lpost + i * S;
        float *tmp = p; p = v; v = tmp;                            /* decode.py:62 */
        int32_t *tbi = tb + i * K;
        for (long j = 0; j < K; j++) {
            /* step, decode.py:65-68: first maximum over a */
            long r = j / nstep;
            float ss = p[r]; long fs = r;
            for (int a = 1; a < nstep; a++) {
                float c = p[a * rstep + r];
                if (c > ss) { ss = c; fs = a * rstep + r; }
            }
            /* skip, decode.py:70-73 */
            long q = j / nskip;
            float sk = p[q]; long fk = q;
            for (int a = 1; a < nskip; a++) {
                float c = p[a * rskip + q];
                if (c > sk) { sk = c; fk = a * rskip + q; }
            }
            sk = sk - skip_pen;
            float best = ss > sk ? ss : sk;                        /* np.maximum */
            long from = ss > sk ? fs : fk;                         /* decode.py:76, tie -> skip */
            float move = lp[1 + j] + best;                         /* decode.py:75 */
            float stay = p[j] + lp[0];                             /* decode.py:80 */
            tbi[j] = move > stay ? (int32_t)from : -1;             /* decode.py:81, tie -> stay */
            v[j] = move > stay ? move : stay;                      /* decode.py:82 */
        }
    }
    long cur = 0;
    for (long j = 1; j < K; j++) if (v[j] > v[cur]) cur = j;       /* np.argmax: first maximum */
    *score_out = v[cur];
    int n = 0;
    path_out[n++] = (int)cur;
    for (long i = nev - 1; i > 0; i--) {                           /* decode.py:86-91 */
        int32_t t = tb[i * K + cur];
        if (t >= 0) { path_out[n++] = t; cur = t; }
    }
    for (int a = 0, b = n - 1; a < b; a++, b--) { int t = path_out[a]; path_out[a] = path_out[b]; path_out[b] = t; }
    free(v); free(p);
    return n;
}

/* Batch of reads laid out [T, B, S] (time-major, like the network output); lengths[b] events each.
 * paths_out is [B, T]; path_len[b] receives the length.  Reads are dealt to `nthreads` pthreads
 * (libgomp is not in the image).  Returns 0 or -1. */
#include <pthread.h>

typedef struct {
    const float *lpost; long T, B, S, K; const int *lengths; int nbase, klen; float skip_pen;
    int *paths_out, *path_len; float *score_out;
    long next; int bad; pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *job = (batch_job *)arg;
    for (;;) {
        pthread_mutex_lock(&job->mu);
        long b = job->next++;
        pthread_mutex_unlock(&job->mu);
        if (b >= job->B) break;
        long n = job->lengths ? job->lengths[b] : job->T;
        int len = -1;
        if (n >= 1) {
            float *lp = (float *)malloc(sizeof(float) * n * job->S);
            int32_t *tb = (int32_t *)malloc(sizeof(int32_t) * n * job->K);
            if (lp && tb) {
                for (long t = 0; t < n; t++)
                    memcpy(lp + t * job->S, job->lpost + (t * job->B + b) * job->S, sizeof(float) * job->S);
                len = sloika_oracle_viterbi(lp, n, job->nbase, job->klen, job->skip_pen, tb,
                                            job->paths_out + b * job->T, job->score_out + b);
            }
            free(lp); free(tb);
        }
        job->path_len[b] = len;
        if (len < 0) job->bad = 1;
    }
    return 0;
}

int sloika_oracle_viterbi_batch(const float *lpost, long T, long B, const int *lengths, int nbase, int klen,
                                float skip_pen, int nthreads, int *paths_out, int *path_len, float *score_out)
{
    batch_job job;
    job.lpost = lpost; job.T = T; job.B = B; job.lengths = lengths; job.nbase = nbase; job.klen = klen;
    job.skip_pen = skip_pen; job.paths_out = paths_out; job.path_len = path_len; job.score_out = score_out;
    job.K = 1;
    for (int i = 0; i < klen; i++) job.K *= nbase;
    job.S = job.K + 1; job.next = 0; job.bad = 0;
    pthread_mutex_init(&job.mu, 0);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads - 1; i++)
        if (pthread_create(&th[started], 0, batch_worker, &job) == 0) started++;
    batch_worker(&job);
    for (int i = 0; i < started; i++) pthread_join(th[i], 0);
    pthread_mutex_destroy(&job.mu);
    return job.bad ? -1 : 0;
}
