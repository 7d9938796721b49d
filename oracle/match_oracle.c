/*
 * oracle/match_oracle.c — CPU restatement of XRSfM's SIFT descriptor matcher.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under xrsfm_b200/ (the product) may link, import
 * or execute this file; it is the checker used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY STATUS: openxrlab/xrsfm ships no tests, golden vectors or CPU SIFT matcher
 * (SURVEY.md §4, §8c), so nothing of the reference's own pins this path.  This oracle is
 * PINNED AGAINST OUTPUTS OF THE REFERENCE ITSELF RUN HERE: the reference's CUDA kernels,
 * compiled verbatim from /root/reference/3rdparty/SiftGPU/ProgramCU.cu into oracle/_ref/,
 *   (a) produced the committed golden vectors tests/golden/match_ref_golden.npz on a B200
 *       (generator: tests/golden/make_match_golden.py; 8 cases: ragged sizes, loose and
 *       tight thresholds, no mutual filter, truncation at max_match) — checked without a
 *       GPU by tests/test_match_oracle.py::test_oracle_equals_reference_kernel_golden;
 *   (b) are run live next to the oracle and the product on the GPU box
 *       (tests/test_match_gpu.py::test_reference_kernels_agree);
 * plus an independent numpy int32-matmul restatement (tests/test_match_oracle.py).
 *
 * What is restated (paths relative to the reference tree):
 *   - MultiplyDescriptor_Kernel   3rdparty/SiftGPU/ProgramCU.cu:1491-1578
 *       dot[i1][i2] = sum_d a[i1][d]*b[i2][d] (exact int32) and, per 8-row block and
 *       column, the partial (best, idx, second) with init (0,-1,0), strict '>' for best
 *       and max() for second.
 *   - RowMatch_Kernel             ProgramCU.cu:1780-1837
 *       32 lanes stride the row; per lane strict '>' scan; 16/8/4/2/1 tree merge that
 *       prefers the lower lane on ties; thresholds in float after a DOUBLE acos
 *       (min(float,double) promotes).
 *   - ColMatch_Kernel             ProgramCU.cu:1852-1872
 *       merge of the 8-row partials in block order, strict '<'.
 *   - SiftMatchCU::GetBestMatch   3rdparty/SiftGPU/SiftMatchCU.cpp:186-215
 *       ascending i, keep (i, m12[i]) iff m12[i] >= 0 and (!mbm or m21[m12[i]] == i),
 *       stop at max_match.
 *   - SiftMatchCU::SetDescriptors SiftMatchCU.cpp:100-118 (num clamped to max_sift).
 * Constants used by the caller: feature_processing.cc:118-154 (0.7 / 0.8 / 16384 / mbm).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define XRO_DIM 128
#define XRO_SCALE 0.000003814697265625f /* 2^-18, ProgramCU.cu:1830 */

/* float(acos(double(min(float(dot)*2^-18, 1.0)))) — ProgramCU.cu:1830-1831,1865-1866 */
float xro_dist_of_dot(int dot) {
    float x = (float)dot * XRO_SCALE;
    double xm = fmin((double)x, 1.0);
    return (float)acos(xm);
}

/* (dist < distmax) && (dist < distn * ratiomax) — ProgramCU.cu:1833,1868 */
int xro_accept(int best, int second, float distmax, float ratiomax) {
    float dist = xro_dist_of_dot(best);
    float distn = xro_dist_of_dot(second);
    return (dist < distmax) && (dist < distn * ratiomax);
}

/* dot[i][j] = sum_d a[i][d] * b[j][d], exact in int32.  B is repacked as Bt[d/2][j][2] (int16
 * pairs) so that one vpmaddwd yields 16 consecutive j of a row without horizontal sums; the
 * AVX-512BW path is chosen at run time, the portable loop computes the same integers. */
#include <immintrin.h>

static void dot_rows_generic(const int16_t *a, const int16_t *bt, int n2, int i0, int i1, int32_t *dot) {
    for (int i = i0; i < i1; ++i) {
        int32_t *out = dot + (size_t)i * n2;
        for (int j = 0; j < n2; ++j) out[j] = 0;
        for (int dp = 0; dp < XRO_DIM / 2; ++dp) {
            const int32_t a0 = a[(size_t)i * XRO_DIM + 2 * dp], a1 = a[(size_t)i * XRO_DIM + 2 * dp + 1];
            const int16_t *bp = bt + (size_t)dp * n2 * 2;
            for (int j = 0; j < n2; ++j) out[j] += a0 * bp[2 * j] + a1 * bp[2 * j + 1];
        }
    }
}

__attribute__((target("avx512f,avx512bw"))) static void
dot_rows_avx512(const int16_t *a, const int16_t *bt, int n2, int i0, int i1, int32_t *dot) {
    int i = i0;
    for (; i + 4 <= i1; i += 4) {
        int j = 0;
        for (; j + 32 <= n2; j += 32) {
            __m512i acc[4][2];
            for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = _mm512_setzero_si512();
            for (int dp = 0; dp < XRO_DIM / 2; ++dp) {
                const int16_t *bp = bt + ((size_t)dp * n2 + j) * 2;
                const __m512i b0 = _mm512_loadu_si512((const void *)bp);
                const __m512i b1 = _mm512_loadu_si512((const void *)(bp + 32));
                for (int r = 0; r < 4; ++r) {
                    int32_t pair;
                    memcpy(&pair, a + (size_t)(i + r) * XRO_DIM + 2 * dp, 4);
                    const __m512i av = _mm512_set1_epi32(pair);
                    acc[r][0] = _mm512_add_epi32(acc[r][0], _mm512_madd_epi16(av, b0));
                    acc[r][1] = _mm512_add_epi32(acc[r][1], _mm512_madd_epi16(av, b1));
                }
            }
            for (int r = 0; r < 4; ++r) {
                _mm512_storeu_si512((void *)(dot + (size_t)(i + r) * n2 + j), acc[r][0]);
                _mm512_storeu_si512((void *)(dot + (size_t)(i + r) * n2 + j + 16), acc[r][1]);
            }
        }
        if (j < n2) { /* ragged tail of the 4 rows */
            for (int r = 0; r < 4; ++r)
                for (int jj = j; jj < n2; ++jj) {
                    int32_t sum = 0;
                    for (int dp = 0; dp < XRO_DIM / 2; ++dp)
                        sum += (int32_t)a[(size_t)(i + r) * XRO_DIM + 2 * dp] * bt[((size_t)dp * n2 + jj) * 2] +
                               (int32_t)a[(size_t)(i + r) * XRO_DIM + 2 * dp + 1] * bt[((size_t)dp * n2 + jj) * 2 + 1];
                    dot[(size_t)(i + r) * n2 + jj] = sum;
                }
        }
    }
    if (i < i1) dot_rows_generic(a, bt, n2, i, i1, dot);
}

/* Full int32 dot matrix, row-major [n1][n2] (ProgramCU.cu:1536-1554). */
void xro_dot_matrix(int n1, const uint8_t *d1, int n2, const uint8_t *d2, int32_t *dot) {
    int16_t *a = (int16_t *)malloc((size_t)n1 * XRO_DIM * sizeof(int16_t));
    int16_t *bt = (int16_t *)malloc((size_t)n2 * XRO_DIM * sizeof(int16_t));
    for (size_t i = 0; i < (size_t)n1 * XRO_DIM; ++i) a[i] = d1[i];
    for (int j = 0; j < n2; ++j)
        for (int dp = 0; dp < XRO_DIM / 2; ++dp) {
            bt[((size_t)dp * n2 + j) * 2] = d2[(size_t)j * XRO_DIM + 2 * dp];
            bt[((size_t)dp * n2 + j) * 2 + 1] = d2[(size_t)j * XRO_DIM + 2 * dp + 1];
        }
    const int fast = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f");
#pragma omp parallel for schedule(dynamic, 1)
    for (int i0 = 0; i0 < n1; i0 += 32) {
        const int i1 = i0 + 32 < n1 ? i0 + 32 : n1;
        if (fast)
            dot_rows_avx512(a, bt, n2, i0, i1, dot);
        else
            dot_rows_generic(a, bt, n2, i0, i1, dot);
    }
    free(a);
    free(bt);
}

/* RowMatch_Kernel for one row (ProgramCU.cu:1796-1835). */
static int row_match(const int32_t *row, int n2, float distmax, float ratiomax) {
    int mx[32], nx[32], ix[32];
    for (int t = 0; t < 32; ++t) mx[t] = 0, nx[t] = 0, ix[t] = -1;
    /* lane t of the kernel visits j = t, t+32, ... in order; walking j in natural order and
     * updating lane j % 32 performs exactly the same per-lane sequence (and vectorises). */
    for (int j0 = 0; j0 < n2; j0 += 32) {
        const int lim = n2 - j0 < 32 ? n2 - j0 : 32;
        for (int t = 0; t < lim; ++t) {
            const int v = row[j0 + t];
            const int test = v > mx[t];
            const int nz = nx[t] > v ? nx[t] : v;
            nx[t] = test ? mx[t] : nz;
            ix[t] = test ? j0 + t : ix[t];
            mx[t] = test ? v : mx[t];
        }
    }
    for (int step = 16; step > 0; step /= 2) {
        for (int t = 0; t < step; ++t) {
            int v1 = mx[t], v2 = mx[t + step];
            int test = v2 > v1;
            int a = test ? v1 : nx[t];
            int b = test ? nx[t + step] : v2;
            nx[t] = a > b ? a : b;
            ix[t] = test ? ix[t + step] : ix[t];
            mx[t] = test ? v2 : v1;
        }
    }
    return xro_accept(mx[0], nx[0], distmax, ratiomax) ? ix[0] : -1;
}

/* Column partials of MultiplyDescriptor_Kernel (ProgramCU.cu:1556-1570) merged as
 * ColMatch_Kernel does (ProgramCU.cu:1858-1870), for the column stripe [c0, c1).  Rows are
 * streamed in order (cache-friendly); per column the arithmetic and its order are exactly the
 * kernels': 8-row partial (best, idx, second) with init (0,-1,0), then the block merge. */
static void col_match_stripe(const int32_t *dot, int n1, int n2, int c0, int c1, float distmax,
                             float ratiomax, int32_t *m21) {
    const int w = c1 - c0;
    int32_t *buf = (int32_t *)malloc((size_t)w * 6 * sizeof(int32_t));
    int32_t *cx = buf, *cy = buf + w, *cz = buf + 2 * w, *rx = buf + 3 * w, *ry = buf + 4 * w, *rz = buf + 5 * w;
    const int nblk = (n1 + 7) / 8;
    for (int blk = 0; blk < nblk; ++blk) {
        for (int j = 0; j < w; ++j) cx[j] = 0, cy[j] = -1, cz[j] = 0; /* make_int3(0,-1,0) */
        for (int i = 0; i < 8; ++i) {
            const int r = blk * 8 + i;
            if (r >= n1) break;
            const int32_t *row = dot + (size_t)r * n2 + c0;
            for (int j = 0; j < w; ++j) {
                const int v = row[j];
                const int gt = v > cx[j];
                const int nz = cz[j] > v ? cz[j] : v;
                cz[j] = gt ? cx[j] : nz;
                cy[j] = gt ? r : cy[j];
                cx[j] = gt ? v : cx[j];
            }
        }
        if (blk == 0) {
            memcpy(rx, cx, (size_t)w * 4), memcpy(ry, cy, (size_t)w * 4), memcpy(rz, cz, (size_t)w * 4);
        } else {
            for (int j = 0; j < w; ++j) {
                if (rx[j] < cx[j]) {
                    rz[j] = rx[j] > cz[j] ? rx[j] : cz[j], rx[j] = cx[j], ry[j] = cy[j];
                } else {
                    rz[j] = rz[j] > cx[j] ? rz[j] : cx[j];
                }
            }
        }
    }
    for (int j = 0; j < w; ++j) m21[c0 + j] = xro_accept(rx[j], rz[j], distmax, ratiomax) ? ry[j] : -1;
    free(buf);
}

/*
 * One image pair, the whole of SiftMatchCU::GetSiftMatch (SiftMatchCU.cpp:175-215).
 * m12 (n1 ints) and m21 (n2 ints) may be NULL.  Returns the number of matches written.
 */
int xro_match_pair(int n1, const uint8_t *d1, int n2, const uint8_t *d2, float distmax,
                   float ratiomax, int mbm, int max_match, uint32_t (*out)[2],
                   int32_t *m12_out, int32_t *m21_out) {
    if (n1 <= 0 || n2 <= 0) return 0; /* SiftMatchCU.cpp:179-180 */
    /* grow-only scratch for the n1 x n2 matrix: re-faulting 64 MB of fresh pages per pair
     * would dominate the baseline */
    static int32_t *dot_buf = NULL;
    static size_t dot_cap = 0;
    if ((size_t)n1 * n2 > dot_cap) {
        free(dot_buf);
        dot_cap = (size_t)n1 * n2;
        dot_buf = (int32_t *)malloc(dot_cap * sizeof(int32_t));
    }
    int32_t *dot = dot_buf;
    int32_t *m12 = (int32_t *)malloc((size_t)n1 * sizeof(int32_t));
    int32_t *m21 = (int32_t *)malloc((size_t)n2 * sizeof(int32_t));
    xro_dot_matrix(n1, d1, n2, d2, dot);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n1; ++i)
        m12[i] = row_match(dot + (size_t)i * n2, n2, distmax, ratiomax);
    if (mbm) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int c0 = 0; c0 < n2; c0 += 128)
            col_match_stripe(dot, n1, n2, c0, c0 + 128 < n2 ? c0 + 128 : n2, distmax, ratiomax, m21);
    } else {
        for (int j = 0; j < n2; ++j) m21[j] = -1;
    }
    int nmatch = 0;
    for (int i = 0; i < n1 && nmatch < max_match; ++i) { /* SiftMatchCU.cpp:199-207 */
        int j = m12[i];
        if (j >= 0 && (!mbm || m21[j] == i)) {
            out[nmatch][0] = (uint32_t)i;
            out[nmatch][1] = (uint32_t)j;
            nmatch++;
        }
    }
    if (m12_out) memcpy(m12_out, m12, (size_t)n1 * sizeof(int32_t));
    if (m21_out) memcpy(m21_out, m21, (size_t)n2 * sizeof(int32_t));
    free(m12);
    free(m21);
    return nmatch;
}

/*
 * Batched form mirroring xrb_match_pairs: images packed in one block, image i at
 * row_offsets[i]; counts clamped to max_features (SiftMatchCU.cpp:113-114).
 */
int xro_match_pairs(int n_pairs, const int32_t (*pairs)[2], const int64_t *row_offsets,
                    const uint8_t *block, int max_features, float distmax, float ratiomax,
                    int mbm, int max_match, int64_t *out_offsets, uint32_t (*out)[2],
                    int64_t out_capacity) {
    int64_t off = 0;
    int overflow = 0;
    out_offsets[0] = 0;
    for (int p = 0; p < n_pairs; ++p) {
        int a = pairs[p][0], b = pairs[p][1];
        int n1 = (int)(row_offsets[a + 1] - row_offsets[a]);
        int n2 = (int)(row_offsets[b + 1] - row_offsets[b]);
        if (n1 > max_features) n1 = max_features;
        if (n2 > max_features) n2 = max_features;
        int cap = n1 < max_match ? n1 : max_match;
        uint32_t(*tmp)[2] = (uint32_t(*)[2])malloc((size_t)(cap > 0 ? cap : 1) * 8);
        int n = xro_match_pair(n1, block + row_offsets[a] * XRO_DIM, n2,
                               block + row_offsets[b] * XRO_DIM, distmax, ratiomax, mbm,
                               max_match, tmp, NULL, NULL);
        if (off + n <= out_capacity)
            memcpy(out + off, tmp, (size_t)n * 8);
        else
            overflow = 1;
        off += n;
        out_offsets[p + 1] = off;
        free(tmp);
    }
    return overflow ? -4 : 0;
}
