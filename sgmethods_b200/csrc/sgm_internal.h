// sgm_internal.h -- shared between the host builder (sgm_host.cpp) and the CUDA side.
#ifndef SGM_INTERNAL_H
#define SGM_INTERNAL_H

#include "../../include/sgm_b200.h"

// Records `msg` as the calling thread's last error and returns `code`.
int sgm_fail(int code, const char* msg);
int sgm_failf(int code, const char* fmt, ...);

#endif
