/* Shim so that unmodified reference callers, which include "aecm/echo_control_mobile.h" (main.cc:16),
 * pick up the drop-in header when compiled with -I <repo>/include. */
#include "../echo_control_mobile.h"
