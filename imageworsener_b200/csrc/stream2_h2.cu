// stream2_h2.cu -- horizontal streaming kernel, epilogue variant 2 (see stream2_h.inc).
#define IWB2_EPI 2
#include "stream2_h.inc"
