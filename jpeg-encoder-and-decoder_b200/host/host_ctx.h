/* process-wide libjpgb200 context of the host shim (one GPU: $JPGB200_DEVICE, default 0) */
#ifndef HOST_CTX_H
#define HOST_CTX_H
#include "jpgb200.h"
jpgb200_ctx *host_ctx(void);
const char *host_ctx_error(void);
#endif
