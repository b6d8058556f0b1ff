// The two-level K2 kernels of nnls4.cu once more, with the per-phase cycle counters compiled in
// (selected by the dispatcher when PLS_K2_PHASES is set; the default build carries no clock64 reads).
#define PLS_K4_PROF 1
#define K4_NAME(x) x##_prof
#include "nnls4.cu"
