/* COMPILE-CHECK ONLY: see gts.h in this directory */
#include "gts.h"
typedef struct _GModule GModule;
