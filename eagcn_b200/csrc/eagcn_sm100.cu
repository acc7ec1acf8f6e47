// Unity translation unit of libeagcn_sm100.so (one nvcc invocation, no relocatable device code).
#include "gemm_simt.cu"
#include "pack.cu"
#include "rows.cu"
#include "layer_fwd.cu"
#include "layer_bwd.cu"
#include "attention.cu"

extern "C" int eagcn_version(void) { return EAGCN_ABI_VERSION; }
