// fp32 instantiation of the step kernels (throughput build; FMA contraction allowed).
#define B2B_REAL float
#define B2B_SUFFIX f32
#include "b2b_kernels.inl"
