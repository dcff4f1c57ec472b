// stand-in for <gsl/gsl_sf_gamma.h>: see gsl_stub.h (test infrastructure, oracle/ only)
#include "gsl_stub.h"
