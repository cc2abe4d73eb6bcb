// bsg_internal.h -- error plumbing shared by the translation units of libbsg_b200.so.
#pragma once
#include <cuda_runtime.h>

#include "../../include/bsg.h"

// records a thread-local message and returns `code`
int bsg_fail(int code, const char* msg);
// maps a cudaError_t to BSG_OK / BSG_ECUDA (recording the CUDA error string with `what`)
int bsg_cuda_check(cudaError_t e, const char* what);

#define BSG_STR2(x) #x
#define BSG_STR(x) BSG_STR2(x)
#define BSG_CUDA(call)                                                                  \
    do {                                                                                \
        int _rc = bsg_cuda_check((call), #call " @ " __FILE__ ":" BSG_STR(__LINE__));   \
        if (_rc != BSG_OK) return _rc;                                                  \
    } while (0)
