// fp64 instantiation of the step kernels (parity build; compile this unit with -fmad=false).
#define B2B_REAL double
#define B2B_SUFFIX f64
#include "b2b_kernels.inl"
