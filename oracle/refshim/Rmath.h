/* oracle/refshim stand-in (TEST INFRASTRUCTURE ONLY): everything lives in RcppArmadillo.h */
#include "RcppArmadillo.h"
