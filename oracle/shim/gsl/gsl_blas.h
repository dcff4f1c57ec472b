// stand-in for <gsl/gsl_blas.h>: see gsl_stub.h (test infrastructure, oracle/ only)
#include "gsl_stub.h"
