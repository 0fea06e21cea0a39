/* forwards to the stand-ins in qt_shim_core.h (test infrastructure, see that header) */
#include "../qt_shim_core.h"
