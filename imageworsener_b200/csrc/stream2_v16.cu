// stream2_v16.cu -- vertical streaming kernel, 16-bit big-endian input (see stream2_v.inc).
#define IWB2_D16 1
#include "stream2_v.inc"
