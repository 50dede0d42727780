/* TEST INFRASTRUCTURE: stand-in for <rlglue/utils/C/RLStruct_util.h>.  The product header
 * include/rlglue_types.h deliberately does not declare these two helpers (a real RL-Glue build gets
 * them from librlutils); they are declared here and defined in oracle/ref_wrapper.cpp. */
#include "../../../../../include/rlglue_types.h"
#ifdef __cplusplus
extern "C" {
#endif
void allocateRLStruct(rl_abstract_type_t* dst, unsigned int numInts, unsigned int numDoubles, unsigned int numChars);
void clearRLStruct(rl_abstract_type_t* dst);
#ifdef __cplusplus
}
#endif
