/* COMPILE-CHECK ONLY: see gts.h in this directory */
#include "gts.h"
