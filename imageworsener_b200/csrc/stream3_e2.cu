// stream3_e2.cu -- fused vertical + horizontal kernel, epilogue variant 2 (see stream3.inc).
#define IWB3_EPI 2
#include "stream3.inc"
