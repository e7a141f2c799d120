// stream3_e3.cu -- fused vertical + horizontal kernel, epilogue variant 3 (see stream3.inc).
#define IWB3_EPI 3
#include "stream3.inc"
