/* TEST INFRASTRUCTURE — see gsl_linalg.h in this directory. */
#include "gsl_linalg.h"
