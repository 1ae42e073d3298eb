/* TEST INFRASTRUCTURE ONLY: see Rinternals.h */
#include "Rinternals.h"
