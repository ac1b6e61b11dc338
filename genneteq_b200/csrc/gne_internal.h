// genneteq_b200 — declarations shared between the translation units of the library (not part
// of the C ABI).
#pragma once
#include <stdint.h>
#include <vector>

struct gne_dev;

void gne_set_error(const char *fmt, ...);
int gne_dev_flush(gne_dev *d);
void gne_dev_traffic(gne_dev *d, int64_t *h2d, int64_t *d2h);
int gne_dev_price_lv_sharded(gne_dev *d, double *largest, std::vector<int64_t> &list, int *any);

#ifdef GNE_TUNING
// tuning harness (not part of the C ABI; `make -C genneteq_b200/csrc tuning` builds lib/libgenneteq_b200_tuning.so with it):
// times the LV kernels and their experimental variants (msOut[2v] = mean ms, msOut[2v+1] = best ms over reps)
extern "C" int gne_bench_variants(gne_dev *d, int reps, float *msOut, int maxVariants, int *nVariants);
extern "C" const char *gne_bench_variant_name(int i);
#endif
