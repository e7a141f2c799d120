// stream2_h1.cu -- horizontal streaming kernel, epilogue variant 1 (see stream2_h.inc).
#define IWB2_EPI 1
#include "stream2_h.inc"
