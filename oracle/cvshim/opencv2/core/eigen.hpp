// ORACLE cv shim: forwards to cvshim.h (see there).
#include "../../cvshim.h"
