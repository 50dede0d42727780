/* TEST INFRASTRUCTURE: stand-in for <rlglue/utils/C/RLStruct_util.h>; the two helpers
 * are declared in include/rlglue_types.h and defined in oracle/ref_wrapper.cpp. */
#include "../../../../../include/rlglue_types.h"
