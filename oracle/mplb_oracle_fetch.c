/* ORACLE -- TEST INFRASTRUCTURE ONLY (see mplb_oracle.h).  Fetch path: filled in below. */
#include "mplb_oracle.h"
