/*
 * oracle/ans_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's entropy-coder arithmetic
 * (fhkingma/bitswap @ dfe0bf7d).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * shipped path (bitswap_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 * against tests/golden/*.json|npz, which tests/golden/make_golden.py generated
 * by running the reference's own `ANS` class / `logistic_cdf` in the build
 * container (the reference holds no golden vectors of its own).
 *
 * State convention (reference: Python list, cifar_compress.py:157-159):
 *   words[0..n)  = the 32-bit stack, bottom first   (list[:-1])
 *   head         = the 64-bit top element            (list[-1])
 * Invariant after every symbol-op: 2^32 <= head < 2^64.
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>

#define ORC_OK          0
#define ORC_E_UNDERFLOW 1   /* reference: IndexError from x.pop(-2), cifar_compress.py:65 */
#define ORC_E_OVERFLOW  2   /* caller buffer too small (no reference equivalent: lists grow) */
#define ORC_E_BADTABLE  3   /* reference: AssertionError, cifar_compress.py:45-46 */

/* cifar_compress.py:182-184 + utils/torch/rand.py:67-68.
 * endpoints[L][S-1], mu/scale indexed [i*ms] (ms = 0 broadcasts a single
 * value, as the prior does at cifar_compress.py:245).  pmfs[L][S].
 * sigmoid restated as 1/(1+exp(-t)) on the C library's exp; torch-CPU's own
 * vectorised sigmoid can differ from this by 1 ulp, which is
 * why oracle.py keeps a torch path for pinning and this one for speed. */
void orc_logistic_pmfs(const double *endpoints, const double *mu, const double *scale,
                       int64_t ms, int64_t L, int64_t S, double *pmfs)
{
    for (int64_t i = 0; i < L; ++i) {
        const double *e = endpoints + i * (S - 1);
        double *p = pmfs + i * S;
        double m = mu[i * ms], s = scale[i * ms];
        double prev = 0.0;
        for (int64_t k = 0; k < S - 1; ++k) {
            double t = (e[k] - m) / s;
            double c = 1.0 / (1.0 + exp(-t));
            p[k] = (k == 0) ? c : c - prev;          /* :183 adjacent differences, :184 first = cdf_0 */
            prev = c;
        }
        p[S - 1] = 1.0 - prev;                       /* :184 last = 1 - cdf_last */
    }
}

/* ANS.__init__, cifar_compress.py:13-46.  P[L][S], C[L][S+1] as int64 like
 * the reference's torch.long tables. */
int orc_tables(const double *pmfs, int64_t L, int64_t S, int bits, int quantbits,
               int64_t *P, int64_t *C)
{
    const int64_t multiplier = ((int64_t)1 << bits) - ((int64_t)1 << quantbits);   /* :28 */
    for (int64_t i = 0; i < L; ++i) {
        const double *p = pmfs + i * S;
        int64_t *Pi = P + i * S, *Ci = C + i * (S + 1);
        int64_t sum = 0, best = 0, bestv = INT64_MIN;
        for (int64_t k = 0; k < S; ++k) {
            int64_t v = (int64_t)(p[k] * (double)multiplier) + 1;   /* :29 .long() truncates; :32 +1 */
            Pi[k] = v;
            sum += v;
            if (v > bestv) { bestv = v; best = k; }                 /* :35 argmax, first index wins ties */
        }
        Pi[best] += ((int64_t)1 << bits) - sum;                     /* :35 remnant */
        Ci[0] = 0;                                                  /* :39 */
        for (int64_t k = 0; k < S; ++k) Ci[k + 1] = Ci[k] + Pi[k];  /* :38 */
        if (Ci[S] != ((int64_t)1 << bits)) return ORC_E_BADTABLE;   /* :46 */
        if (Pi[best] <= 0) return ORC_E_BADTABLE;
    }
    return ORC_OK;
}

/* ANS.encode, cifar_compress.py:48-55.  Rows ascending. */
int orc_push(uint32_t *words, int64_t *nwords, int64_t cap, uint64_t *head,
             const int64_t *P, const int64_t *C, const int64_t *sym,
             int64_t L, int64_t S, int bits)
{
    uint64_t x = *head;
    int64_t n = *nwords;
    for (int64_t i = 0; i < L; ++i) {
        int64_t s = sym[i];
        uint64_t pmf = (uint64_t)P[i * S + s];
        /* :51  ((lbound >> bits) << 32) * pmf  with lbound = 2^32 */
        uint64_t lim = ((((uint64_t)1 << 32) >> bits) << 32) * pmf;
        if (x >= lim) {
            if (n >= cap) return ORC_E_OVERFLOW;
            words[n++] = (uint32_t)(x & 0xffffffffu);   /* :52-53 */
            x >>= 32;
        }
        x = ((x / pmf) << bits) + (x % pmf) + (uint64_t)C[i * (S + 1) + s];   /* :54 */
    }
    *head = x; *nwords = n;
    return ORC_OK;
}

/* ANS.decode, cifar_compress.py:57-67.  Rows descending. */
int orc_pop(uint32_t *words, int64_t *nwords, uint64_t *head,
            const int64_t *P, const int64_t *C, int64_t *sym_out,
            int64_t L, int64_t S, int bits)
{
    uint64_t x = *head;
    int64_t n = *nwords;
    const uint64_t mask = ((uint64_t)1 << bits) - 1;
    for (int64_t i = L - 1; i >= 0; --i) {
        const int64_t *Ci = C + i * (S + 1);
        uint64_t m = x & mask;                               /* :60 */
        /* :61 searchsorted(C[i,:-1], m, 'right') - 1  == max{k : C[k] <= m} */
        int64_t lo = 0, hi = S;                              /* answer in [lo, hi) */
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if ((uint64_t)Ci[mid] <= m) lo = mid; else hi = mid;
        }
        int64_t s = lo;
        sym_out[i] = s;                                      /* :62 */
        x = (uint64_t)P[i * S + s] * (x >> bits) + m - (uint64_t)Ci[s];   /* :63 */
        if (x < ((uint64_t)1 << 32)) {                       /* :64 */
            if (n <= 0) { *head = x; *nwords = n; return ORC_E_UNDERFLOW; }
            x = (x << 32) | words[--n];                      /* :65 */
        }
    }
    *head = x; *nwords = n;
    return ORC_OK;
}
