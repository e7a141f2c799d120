// stream2_h4.cu -- horizontal streaming kernel, epilogue variant 4 (see stream2_h.inc).
#define IWB2_EPI 4
#include "stream2_h.inc"
