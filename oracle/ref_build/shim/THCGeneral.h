/* see THC.h in this directory (oracle test infrastructure) */
#include "THC.h"
