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
