// gf_eval_kernel instantiations for DOUBLE precision (same dispatch code as the MIXED unit).
#define GFB_GENERAL_F64 1
#include "gf_launch_general_f32.cu"
