// TEST INFRASTRUCTURE (oracle).  The reference includes <gmp.h> only as a
// Vivado-HLS workaround (GIN/src/dcl.h:4-6); nothing from it is used.
#ifndef FLOWGNN_ORACLE_SHIM_GMP_H
#define FLOWGNN_ORACLE_SHIM_GMP_H
#endif
