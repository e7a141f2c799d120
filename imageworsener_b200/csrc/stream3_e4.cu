// stream3_e4.cu -- fused vertical + horizontal kernel, epilogue variant 4 (see stream3.inc).
#define IWB3_EPI 4
#include "stream3.inc"
