// stand-in for <gsl/gsl_multimin.h>: see gsl_stub.h (test infrastructure, oracle/ only)
#include "gsl_stub.h"
