/* TEST INFRASTRUCTURE — see cl.h */
#include "cl.h"
