/* Empty stand-in for <mkl.h> so that /root/reference/mv/mv.c compiles unmodified
 * against OpenBLAS (see mkl_cblas.h beside this file). TEST INFRASTRUCTURE ONLY. */
