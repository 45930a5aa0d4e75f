// qgmap_band.cu -- row-band decomposition of one frame pair over several GPUs (NCCL over NVLink).
#include "qgmap_internal.h"
#include <cstring>

struct QgBand { int dummy; };

void qgmap_band_release(qgmap_handle *h) { delete h->band; h->band = nullptr; }
int qgmap_band_refresh(qgmap_handle *h) { (void)h; return QGMAP_OK; }
int qgmap_band_iteration(qgmap_handle *h, long long *launches) { (void)launches; h->err = "band mode not connected"; return QGMAP_ERR_COMM; }
extern "C" int qgmap_band_unique_id(void *id128) { (void)id128; return QGMAP_ERR_COMM; }
extern "C" int qgmap_band_connect(qgmap_handle *h, int rank, int nranks, const void *id) { (void)h; (void)rank; (void)nranks; (void)id; return QGMAP_ERR_COMM; }
