// placeholder (FlowNet2 ops land next)
#include "common.h"
