/*
 * oracle/match_oracle.c — CPU restatement of XRSfM's SIFT descriptor matcher.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under xrsfm_b200/ (the product) may link, import
 * or execute this file; it is the checker used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY STATUS: "parity unpinned" by the reference itself — openxrlab/xrsfm ships no
 * tests, golden vectors or CPU SIFT matcher.  This oracle is pinned
 * instead against the reference's own CUDA kernels, compiled verbatim from
 * /root/reference/3rdparty/SiftGPU/ProgramCU.cu into oracle/_ref/ and run on the GPU box
 * (tests/test_match_gpu.py::test_reference_kernels_agree), and against an independent
 * numpy int32-matmul restatement (tests/test_match_oracle.py).
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

__attribute__((target_clones("arch=skylake-avx512", "avx2", "default"))) static void
dot_row(const int16_t *a, const int16_t *b, int n2, int32_t *out) {
    for (int j = 0; j < n2; ++j) {
        const int16_t *bj = b + (size_t)j * XRO_DIM;
        int32_t s = 0;
        for (int d = 0; d < XRO_DIM; ++d) s += (int32_t)a[d] * (int32_t)bj[d];
        out[j] = s;
    }
}

/* Full int32 dot matrix, row-major [n1][n2] (ProgramCU.cu:1536-1554). */
void xro_dot_matrix(int n1, const uint8_t *d1, int n2, const uint8_t *d2, int32_t *dot) {
    int16_t *a = (int16_t *)malloc((size_t)n1 * XRO_DIM * sizeof(int16_t));
    int16_t *b = (int16_t *)malloc((size_t)n2 * XRO_DIM * sizeof(int16_t));
    for (size_t i = 0; i < (size_t)n1 * XRO_DIM; ++i) a[i] = d1[i];
    for (size_t i = 0; i < (size_t)n2 * XRO_DIM; ++i) b[i] = d2[i];
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n1; ++i)
        dot_row(a + (size_t)i * XRO_DIM, b, n2, dot + (size_t)i * n2);
    free(a);
    free(b);
}

/* RowMatch_Kernel for one row (ProgramCU.cu:1796-1835). */
static int row_match(const int32_t *row, int n2, float distmax, float ratiomax) {
    int mx[32], nx[32], ix[32];
    for (int t = 0; t < 32; ++t) {
        int t_max = 0, t_nxt = 0, t_idx = -1;
        for (int j = t; j < n2; j += 32) {
            int v = row[j];
            int test = v > t_max;
            t_nxt = test ? t_max : (t_nxt > v ? t_nxt : v);
            t_idx = test ? j : t_idx;
            t_max = test ? v : t_max;
        }
        mx[t] = t_max, nx[t] = t_nxt, ix[t] = t_idx;
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
 * ColMatch_Kernel does (ProgramCU.cu:1858-1870). */
static int col_match(const int32_t *dot, int n1, int n2, int col, float distmax,
                     float ratiomax) {
    int rx = 0, ry = -1, rz = 0; /* merged (best, idx, second) */
    int nblk = (n1 + 7) / 8;
    for (int blk = 0; blk < nblk; ++blk) {
        int cx = 0, cy = -1, cz = 0; /* make_int3(0,-1,0) */
        for (int i = 0; i < 8; ++i) {
            int r = blk * 8 + i;
            if (r < n1) {
                int v = dot[(size_t)r * n2 + col];
                if (v > cx) {
                    cz = cx, cx = v, cy = r;
                } else {
                    cz = cz > v ? cz : v;
                }
            }
        }
        if (blk == 0) {
            rx = cx, ry = cy, rz = cz;
        } else if (rx < cx) {
            rz = rx > cz ? rx : cz, rx = cx, ry = cy;
        } else {
            rz = rz > cx ? rz : cx;
        }
    }
    return xro_accept(rx, rz, distmax, ratiomax) ? ry : -1;
}

/*
 * One image pair, the whole of SiftMatchCU::GetSiftMatch (SiftMatchCU.cpp:175-215).
 * m12 (n1 ints) and m21 (n2 ints) may be NULL.  Returns the number of matches written.
 */
int xro_match_pair(int n1, const uint8_t *d1, int n2, const uint8_t *d2, float distmax,
                   float ratiomax, int mbm, int max_match, uint32_t (*out)[2],
                   int32_t *m12_out, int32_t *m21_out) {
    if (n1 <= 0 || n2 <= 0) return 0; /* SiftMatchCU.cpp:179-180 */
    int32_t *dot = (int32_t *)malloc((size_t)n1 * n2 * sizeof(int32_t));
    int32_t *m12 = (int32_t *)malloc((size_t)n1 * sizeof(int32_t));
    int32_t *m21 = (int32_t *)malloc((size_t)n2 * sizeof(int32_t));
    xro_dot_matrix(n1, d1, n2, d2, dot);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n1; ++i)
        m12[i] = row_match(dot + (size_t)i * n2, n2, distmax, ratiomax);
    if (mbm) {
#pragma omp parallel for schedule(static)
        for (int j = 0; j < n2; ++j) m21[j] = col_match(dot, n1, n2, j, distmax, ratiomax);
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
    free(dot);
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
