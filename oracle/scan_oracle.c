/*
 * scan_oracle.c — CPU restatement of the reference's PWM scan, threshold filter and occurrence
 * counts.  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs; never by the product path.
 *
 * PARITY UNPINNED: the reference (kchu25/MOTIFs.jl) ships no tests, fixtures or golden vectors
 * (test/runtests.jl:4-6 is empty) and cannot be executed here (Julia absent; the code is typed to
 * CuArray).  This file follows the reference source line by line instead; every function cites it.
 *
 * Conventions: bases are codes A=0,C=1,G=2,T=3 (loadfasta/helpers.jl:125-128 row order); all
 * indices are 0-based (reference l = pos + 1).  pwms are Float16 bits in Julia column-major
 * (K,4,maxlen): element (k,a,ind) at k + K*(a + 4*ind)  (inference/_h3_1_alignment.jl:66-69).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef _Float16 f16;

typedef struct {
    uint32_t seq, pos;
    uint16_t motif, score_f16;
    uint8_t comp, pad[3];
} oracle_hit;   /* same layout as mb200_hit */

static inline f16 bits2h(uint16_t b) { f16 h; memcpy(&h, &b, 2); return h; }
static inline uint16_t h2bits(f16 h) { uint16_t b; memcpy(&b, &h, 2); return b; }

/* Float16 `a + b` / `a * b` exactly as Julia evaluates them: one correctly rounded operation.
 * (float has 24 >= 2*11+2 significand bits, so going through float does not double-round.) */
static inline f16 hadd(f16 a, f16 b) { return (f16)((float)a + (float)b); }
static inline f16 hmul(f16 a, f16 b) { return (f16)((float)a * (float)b); }

/* reverse(pwm) of a (4 x len) matrix = reverse both dims (inference/_h3_1_alignment.jl:68-69):
 * rc[a][ind] = pwm[3-a][len-1-ind]. */
static inline uint16_t pwm_entry(const uint16_t* pwms, int K, int k, int a, int ind, int len, int rc) {
    if (rc) { a = 3 - a; ind = len - 1 - ind; }
    return pwms[(size_t)k + (size_t)K * ((size_t)a + 4 * (size_t)ind)];
}

/* greedy_search! (inference/_h3_1_alignment.jl:18-36), literal: for every column the four products
 * pwm[k,a,ind] * x[a] with x the one-hot Float16 column are added one after the other into a Float16
 * accumulator that starts at 0 (CUDA.zeros, :75); the result is kept when > 0, else 0 (:33). */
static f16 score_literal(const uint16_t* pwms, int K, int k, int len, int rc, const uint8_t* bases, int64_t l) {
    f16 s = (f16)0.0f;
    for (int ind = 0; ind < len; ++ind) {
        const int b = bases[l + ind];
        for (int a = 0; a < 4; ++a) {
            const f16 x = (a == b) ? (f16)1.0f : (f16)0.0f;
            s = hadd(s, hmul(bits2h(pwm_entry(pwms, K, k, a, ind, len, rc)), x));
        }
    }
    return s > (f16)0.0f ? s : (f16)0.0f;
}

/* Dense pos_scores for one strand, laid out [k][n][l] with l < Lb (the reference allocates 4*Lb in
 * the last dimension, :75, but only l <= Lb - len_k + 1 is ever written). */
void oracle_pos_scores(const uint16_t* pwms, const int64_t* lens, int K, int maxlen, const uint8_t* bases,
                       int64_t N, int64_t Lb, int rc, uint16_t* out) {
    (void)maxlen;
    #pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const int len = (int)lens[k];
            for (int64_t l = 0; l < Lb; ++l) {
                f16 s = (f16)0.0f;
                if (l + len <= Lb) s = score_literal(pwms, K, k, len, rc, bases + n * Lb, l);
                out[((size_t)k * N + n) * Lb + l] = h2bits(s);
            }
        }
}

/* Equivalent table form used for the larger cases and the CPU baseline: because x is one-hot the
 * sum only ever adds pwm[k, base, ind] (finite entries times 0 are signed zeros, which never change
 * a Float16 sum's value); a non-finite entry in a NON-selected row contributes Inf*0 = NaN.
 * tests/test_scan_oracle.py checks this form against score_literal. */
static inline int h_nonfinite(uint16_t b) { return (b & 0x7C00u) == 0x7C00u; }

typedef struct { int len; f16 (*e[2])[4]; } motif_tab;      /* e[rc][ind][base], len columns each */

static void build_tab(const uint16_t* pwms, int K, int k, int len, motif_tab* t) {
    t->len = len;
    t->e[0] = (f16 (*)[4])malloc(sizeof(f16) * 4 * (size_t)len);
    t->e[1] = (f16 (*)[4])malloc(sizeof(f16) * 4 * (size_t)len);
    for (int rc = 0; rc < 2; ++rc)
        for (int ind = 0; ind < len; ++ind) {
            int nf = 0;
            for (int a = 0; a < 4; ++a) nf += h_nonfinite(pwm_entry(pwms, K, k, a, ind, len, rc));
            for (int b = 0; b < 4; ++b) {
                const uint16_t v = pwm_entry(pwms, K, k, b, ind, len, rc);
                t->e[rc][ind][b] = (nf - h_nonfinite(v)) > 0 ? bits2h(0x7E00u) : bits2h(v);
            }
        }
}

static inline f16 score_tab(const motif_tab* t, int rc, const uint8_t* bases, int64_t l) {
    f16 s = (f16)0.0f;
    for (int ind = 0; ind < t->len; ++ind) s = hadd(s, t->e[rc][ind][bases[l + ind]]);
    return s;
}

/* hit test: scan keeps score > 0 (:33,:82); filter_position_by_best_thresh! keeps score .> thresh
 * (inference/_s2_filter_pos_w_scores.jl:116-125).  thresh == NULL = scan only. */
static inline int is_hit(f16 s, const uint16_t* thresh, int k) {
    if (!(s > (f16)0.0f)) return 0;
    if (thresh && !(s > bits2h(thresh[k]))) return 0;
    return 1;
}

/* union_ranges / push_ranges! (inference/_h4_overlap_ratio.jl:40-56) restated literally on
 * [start, stop] pairs.  `for i in eachindex(@view ranges[2:end])` runs i = 1..n-1 but indexes the
 * UNSLICED sorted array, so ranges[1] is merged with itself and ranges[n] is never visited.     */
static int cmp_i64(const void* a, const void* b) { const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b; return (x > y) - (x < y); }
static int64_t union_ranges_total(int64_t* starts, int64_t n, int len) {
    if (n == 0) return 0;
    /* sort(ranges, by = x->x[1]) — stable; all ranges have the same length so stability is moot */
    if (n <= 64) { for (int64_t i = 1; i < n; ++i) { int64_t v = starts[i]; int64_t j = i - 1; while (j >= 0 && starts[j] > v) { starts[j + 1] = starts[j]; --j; } starts[j + 1] = v; } }
    else qsort(starts, (size_t)n, sizeof(int64_t), cmp_i64);
    int64_t total = 0;
    int64_t cur_s = starts[0], cur_e = starts[0] + len - 1;        /* _ranges_ = [ranges[1]] */
    for (int64_t i = 1; i <= n - 1; ++i) {                          /* i in eachindex(view) = 1..n-1 */
        const int64_t rs = starts[i - 1], re = starts[i - 1] + len - 1;   /* ranges[i] (1-based) */
        if (cur_e >= rs) { cur_e = re; }                            /* _ranges_[end] = _ranges_[end][1]:ranges[i][end] */
        else { total += cur_e - cur_s + 1; cur_s = rs; cur_e = re; }       /* push!(_ranges_, ranges[i]) */
    }
    total += cur_e - cur_s + 1;
    return total;                                                   /* get_total_occupied_positions (:71-79) */
}

static int64_t true_union_total(const int64_t* sorted_starts, int64_t n, int len) {
    int64_t total = 0, cu = -1;   /* cu = last covered position */
    for (int64_t i = 0; i < n; ++i) {
        int64_t s = sorted_starts[i], e = s + len - 1;
        if (s > cu) total += len; else if (e > cu) total += e - cu;
        if (e > cu) cu = e;
    }
    return total;
}

/* gpu_scan + filter + counts for all sequences.
 *   strands: bit0 = forward pass (rc=false), bit1 = reverse pass (rc=true)  (gpu_scan :89-99)
 *   hits (optional, capacity hits_cap) are written sorted by (seq, motif, comp, pos): per (motif,seq)
 *   that is the order modify_w_found! builds (:38-52): forward hits ascending, then rc hits ascending.
 *   counts (optional) K*4: n_hits, unique start positions (get_uniq_pos, _h4:5-11), coverage as the
 *   reference computes it (union_ranges), true union coverage.
 *   returns the total number of hits (even when it exceeds hits_cap).                             */
int64_t oracle_scan(const uint16_t* pwms, const int64_t* lens, int K, int maxlen, const uint16_t* thresh,
                    const uint8_t* bases, int64_t N, int64_t Lb, int strands,
                    oracle_hit* hits, int64_t hits_cap, int64_t* counts) {
    (void)maxlen;
    motif_tab* tabs = (motif_tab*)malloc(sizeof(motif_tab) * (size_t)K);
    for (int k = 0; k < K; ++k) build_tab(pwms, K, k, (int)lens[k], &tabs[k]);
    int64_t* per_seq = (int64_t*)calloc((size_t)N + 1, sizeof(int64_t));
    int64_t* cnt = counts ? (int64_t*)calloc((size_t)K * 4, sizeof(int64_t)) : NULL;

    /* pass 1: count hits per sequence (for placement) and accumulate counts.  Work items are (sequence, motif) pairs so that one
     * chromosome-length sequence still uses every thread; all sums are integers, hence independent of the schedule. */
    const int64_t items = N * (int64_t)K;
    #pragma omp parallel
    {
        int64_t* local = cnt ? (int64_t*)calloc((size_t)K * 4, sizeof(int64_t)) : NULL;
        size_t cap = (size_t)(2 * (Lb + 1) < (1 << 20) ? 2 * (Lb + 1) : (1 << 20));      /* grows on demand (thresholded scans of long sequences are sparse) */
        int64_t* pos_buf = (int64_t*)malloc(sizeof(int64_t) * cap);
        int64_t* tmp = (int64_t*)malloc(sizeof(int64_t) * cap);
        #pragma omp for schedule(dynamic, 16)
        for (int64_t it = 0; it < items; ++it) {
            const int64_t n = it / K;
            const int k = (int)(it - n * K);
            const uint8_t* b = bases + n * Lb;
            const int len = tabs[k].len;
            int64_t np = 0;
            for (int rc = 0; rc < 2; ++rc) {
                if (!((strands >> rc) & 1)) continue;
                for (int64_t l = 0; l + len <= Lb; ++l)
                    if (is_hit(score_tab(&tabs[k], rc, b, l), thresh, k)) {
                        if ((size_t)np == cap) { cap *= 2; pos_buf = (int64_t*)realloc(pos_buf, sizeof(int64_t) * cap); tmp = (int64_t*)realloc(tmp, sizeof(int64_t) * cap); }
                        pos_buf[np++] = l;
                    }
            }
            if (np) {
                #pragma omp atomic
                per_seq[n + 1] += np;
            }
            if (local && np) {
                /* positions[m][n] = forward hits then rc hits */
                memcpy(tmp, pos_buf, sizeof(int64_t) * (size_t)np);
                local[k * 4 + 0] += np;
                local[k * 4 + 2] += union_ranges_total(tmp, np, len);      /* sorts tmp */
                int64_t uq = 0; for (int64_t i = 0; i < np; ++i) if (i == 0 || tmp[i] != tmp[i - 1]) ++uq;
                local[k * 4 + 1] += uq;                                    /* unique(positions) */
                local[k * 4 + 3] += true_union_total(tmp, np, len);
            }
        }
        if (local) {
            #pragma omp critical
            for (int i = 0; i < K * 4; ++i) cnt[i] += local[i];
            free(local);
        }
        free(pos_buf); free(tmp);
    }
    for (int64_t n = 0; n < N; ++n) per_seq[n + 1] += per_seq[n];
    const int64_t total = per_seq[N];
    if (counts) { memcpy(counts, cnt, sizeof(int64_t) * (size_t)K * 4); free(cnt); }

    /* pass 2: emit */
    if (hits && total <= hits_cap) {
        #pragma omp parallel for schedule(dynamic, 64)
        for (int64_t n = 0; n < N; ++n) {
            const uint8_t* b = bases + n * Lb;
            int64_t o = per_seq[n];
            for (int k = 0; k < K; ++k) {
                const int len = tabs[k].len;
                for (int rc = 0; rc < 2; ++rc) {
                    if (!((strands >> rc) & 1)) continue;
                    for (int64_t l = 0; l + len <= Lb; ++l) {
                        const f16 s = score_tab(&tabs[k], rc, b, l);
                        if (is_hit(s, thresh, k)) {
                            oracle_hit h; memset(&h, 0, sizeof h);
                            h.seq = (uint32_t)n; h.pos = (uint32_t)l; h.motif = (uint16_t)k; h.score_f16 = h2bits(s); h.comp = (uint8_t)rc;
                            hits[o++] = h;
                        }
                    }
                }
            }
        }
    }
    for (int k = 0; k < K; ++k) { free(tabs[k].e[0]); free(tabs[k].e[1]); }
    free(per_seq); free(tabs);
    return total;
}

/* table-form score of one (motif, strand, sequence, position): exposed so tests can compare the table
 * form with the literal form entry by entry. */
uint16_t oracle_score_tab(const uint16_t* pwms, const int64_t* lens, int K, int k, int rc, const uint8_t* bases, int64_t l) {
    motif_tab t; build_tab(pwms, K, k, (int)lens[k], &t);
    f16 s = score_tab(&t, rc, bases, l);
    free(t.e[0]); free(t.e[1]);
    return h2bits(s > (f16)0.0f ? s : (f16)0.0f);
}
uint16_t oracle_score_literal(const uint16_t* pwms, const int64_t* lens, int K, int k, int rc, const uint8_t* bases, int64_t l) {
    return h2bits(score_literal(pwms, K, k, (int)lens[k], rc, bases, l));
}

/* thread control for the timed CPU baseline: torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently make the
 * "all host cores" reference arm single-threaded */
#include <omp.h>
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_max_threads(void) { return omp_get_max_threads(); }
