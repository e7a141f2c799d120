// stream2_h3.cu -- horizontal streaming kernel, epilogue variant 3 (see stream2_h.inc).
#define IWB2_EPI 3
#include "stream2_h.inc"
