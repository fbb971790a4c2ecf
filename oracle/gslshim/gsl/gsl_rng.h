#include "gsl_shim_all.h"
