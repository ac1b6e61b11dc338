/* TEST INFRASTRUCTURE — a restatement of the few GSL entry points the reference calls, so that the UNMODIFIED reference can
 * be compiled here (oracle/Makefile) and run on instances WITH equal-flow sets.  Never part of the product.
 *
 * Third-party dependency: GNU Scientific Library, linked by the reference as `-lgsl -lgslcblas` (/root/reference/Makefile:4),
 * version not pinned there; GSL itself is absent from this image (no headers, no libgsl).  The reference (dated 2010) was
 * written against GSL 1.x; this file follows the algorithms GSL 1.x .. 2.5 publish, in their operation order:
 *   gsl_linalg_LU_decomp   linalg/lu.c: Gaussian elimination with partial pivoting, column by column; pivot = FIRST row of
 *                          maximal |a_ij| (strict >), row swap, then for every row below: a_ij = a_ij / a_jj and
 *                          a_ik = a_ik - a_ij * a_jk for k > j, in that order;
 *   gsl_linalg_LU_solve    copy b to x, gsl_permute_vector (x'[i] = x[p[i]]), then cblas_dtrsv(Lower, NoTrans, Unit) and
 *                          cblas_dtrsv(Upper, NoTrans, NonUnit) as gslcblas' reference loops run them (cblas/source_trsv_r.h:
 *                          row-major forward / back substitution, tmp -= A_ij * x_j for rising j, then one division).
 * (GSL >= 2.6 factors with a recursive Level-3 algorithm and scales pivot columns by a reciprocal: results may differ from
 * this file in the last bit.  Parity of the equal-flow path is therefore pinned to "the reference + this restatement", and
 * DESIGN.md says so.)
 * Call sites: /root/reference/gnePotential.cpp:145-212, gneFlow.cpp:204-237, 242-289.                                    */
#ifndef GNE_ORACLE_GSL_MINI_H
#define GNE_ORACLE_GSL_MINI_H
#include <cstddef>
#include <cstdlib>
#include <cmath>

struct gsl_vector { size_t size; double *data; };
struct gsl_matrix { size_t size1, size2; double *data; };      /* row-major, tda == size2 */
struct gsl_permutation { size_t size; size_t *data; };

static inline gsl_vector *gsl_vector_alloc(size_t n)
{
    gsl_vector *v = (gsl_vector *) malloc(sizeof(gsl_vector));
    v->size = n; v->data = (double *) malloc(sizeof(double) * (n ? n : 1));
    return v;
}
static inline gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2)
{
    gsl_matrix *m = (gsl_matrix *) malloc(sizeof(gsl_matrix));
    m->size1 = n1; m->size2 = n2; m->data = (double *) malloc(sizeof(double) * (n1 * n2 ? n1 * n2 : 1));
    return m;
}
static inline gsl_permutation *gsl_permutation_alloc(size_t n)
{
    gsl_permutation *p = (gsl_permutation *) malloc(sizeof(gsl_permutation));
    p->size = n; p->data = (size_t *) malloc(sizeof(size_t) * (n ? n : 1));
    for (size_t i = 0; i < n; ++i) p->data[i] = i;              /* gsl_permutation_alloc initialises to the identity */
    return p;
}
static inline void gsl_vector_free(gsl_vector *v) { if (v) { free(v->data); free(v); } }
static inline void gsl_matrix_free(gsl_matrix *m) { if (m) { free(m->data); free(m); } }
static inline void gsl_permutation_free(gsl_permutation *p) { if (p) { free(p->data); free(p); } }
static inline double gsl_vector_get(const gsl_vector *v, size_t i) { return v->data[i]; }
static inline void gsl_vector_set(gsl_vector *v, size_t i, double x) { v->data[i] = x; }
static inline double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j) { return m->data[i * m->size2 + j]; }
static inline void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x) { m->data[i * m->size2 + j] = x; }

static inline int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum)
{
    const size_t N = A->size1;
    *signum = 1;
    for (size_t i = 0; i < N; ++i) p->data[i] = i;              /* gsl_permutation_init */
    for (size_t j = 0; j + 1 < N; ++j) {
        double max = fabs(gsl_matrix_get(A, j, j));
        size_t i_pivot = j;
        for (size_t i = j + 1; i < N; ++i) {
            const double aij = fabs(gsl_matrix_get(A, i, j));
            if (aij > max) { max = aij; i_pivot = i; }
        }
        if (i_pivot != j) {
            for (size_t k = 0; k < N; ++k) {                    /* gsl_matrix_swap_rows */
                const double t = gsl_matrix_get(A, j, k);
                gsl_matrix_set(A, j, k, gsl_matrix_get(A, i_pivot, k));
                gsl_matrix_set(A, i_pivot, k, t);
            }
            const size_t t = p->data[j]; p->data[j] = p->data[i_pivot]; p->data[i_pivot] = t;   /* gsl_permutation_swap */
            *signum = -(*signum);
        }
        const double ajj = gsl_matrix_get(A, j, j);
        if (ajj != 0.0) {
            for (size_t i = j + 1; i < N; ++i) {
                const double aij = gsl_matrix_get(A, i, j) / ajj;
                gsl_matrix_set(A, i, j, aij);
                for (size_t k = j + 1; k < N; ++k) {
                    const double aik = gsl_matrix_get(A, i, k);
                    const double ajk = gsl_matrix_get(A, j, k);
                    gsl_matrix_set(A, i, k, aik - aij * ajk);
                }
            }
        }
    }
    return 0;
}

static inline int gsl_linalg_LU_solve(const gsl_matrix *LU, const gsl_permutation *p, const gsl_vector *b, gsl_vector *x)
{
    const size_t N = LU->size1;
    /* gsl_vector_memcpy(x, b); gsl_permute_vector(p, x): x'[i] = b[p[i]] */
    for (size_t i = 0; i < N; ++i) x->data[i] = b->data[p->data[i]];
    /* cblas_dtrsv(RowMajor, Lower, NoTrans, Unit): forward substitution */
    for (size_t i = 1; i < N; ++i) {
        double tmp = x->data[i];
        for (size_t j = 0; j < i; ++j) tmp -= gsl_matrix_get(LU, i, j) * x->data[j];
        x->data[i] = tmp;
    }
    /* cblas_dtrsv(RowMajor, Upper, NoTrans, NonUnit): back substitution */
    if (N > 0) {
        x->data[N - 1] = x->data[N - 1] / gsl_matrix_get(LU, N - 1, N - 1);
        for (size_t i = N - 1; i > 0 && i--;) {
            double tmp = x->data[i];
            for (size_t j = i + 1; j < N; ++j) tmp -= gsl_matrix_get(LU, i, j) * x->data[j];
            x->data[i] = tmp / gsl_matrix_get(LU, i, i);
        }
    }
    return 0;
}
#endif
