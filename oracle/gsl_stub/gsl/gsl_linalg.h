/* TEST INFRASTRUCTURE — abort-on-call stand-in for GSL (absent from this image).
 *
 * The reference reaches GSL only when a basic equal-flow set exists
 * (/root/reference/gnePotential.cpp:39, gneFlow.cpp:33,192).  Every instance this
 * repo handles has zero equal-flow sets, so none of these functions can run; if one
 * ever does, the oracle aborts instead of returning a wrong answer.
 */
#ifndef GNE_ORACLE_GSL_STUB_H
#define GNE_ORACLE_GSL_STUB_H
#include <cstdio>
#include <cstdlib>
#include <cstddef>

struct gsl_vector { int unused; };
struct gsl_matrix { int unused; };
struct gsl_permutation { int unused; };

static inline void gne_gsl_stub_abort(const char *fn)
{
    fprintf(stderr, "oracle GSL stub: %s called (equal-flow sets are out of scope)\n", fn);
    abort();
}
static inline gsl_vector *gsl_vector_alloc(size_t) { gne_gsl_stub_abort("gsl_vector_alloc"); return 0; }
static inline gsl_matrix *gsl_matrix_alloc(size_t, size_t) { gne_gsl_stub_abort("gsl_matrix_alloc"); return 0; }
static inline gsl_permutation *gsl_permutation_alloc(size_t) { gne_gsl_stub_abort("gsl_permutation_alloc"); return 0; }
static inline void gsl_vector_free(gsl_vector *) { gne_gsl_stub_abort("gsl_vector_free"); }
static inline void gsl_matrix_free(gsl_matrix *) { gne_gsl_stub_abort("gsl_matrix_free"); }
static inline void gsl_permutation_free(gsl_permutation *) { gne_gsl_stub_abort("gsl_permutation_free"); }
static inline double gsl_vector_get(const gsl_vector *, size_t) { gne_gsl_stub_abort("gsl_vector_get"); return 0.0; }
static inline void gsl_vector_set(gsl_vector *, size_t, double) { gne_gsl_stub_abort("gsl_vector_set"); }
static inline void gsl_matrix_set(gsl_matrix *, size_t, size_t, double) { gne_gsl_stub_abort("gsl_matrix_set"); }
static inline int gsl_linalg_LU_decomp(gsl_matrix *, gsl_permutation *, int *) { gne_gsl_stub_abort("gsl_linalg_LU_decomp"); return 0; }
static inline int gsl_linalg_LU_solve(const gsl_matrix *, const gsl_permutation *, const gsl_vector *, gsl_vector *) { gne_gsl_stub_abort("gsl_linalg_LU_solve"); return 0; }
#endif
