/* tests/r_stub -- TEST INFRASTRUCTURE (see Rinternals.h) */
#ifndef R_STUB_R_H
#define R_STUB_R_H
#include <stddef.h>
#endif
