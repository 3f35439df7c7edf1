// Unity build of libbpm_b200.so (one translation unit: kernels inline across files,
// no relocatable device code needed).
#include "filter.cu"
#include "select.cu"
#include "peaks.cu"
#include "floor.cu"
#include "metrics.cu"
#include "pipeline.cu"
