/* TEST INFRASTRUCTURE: stand-in for RL-Glue's <rlglue/Environment_common.h> so the
 * reference SwimmerEnvironment.cpp compiles as an oracle.  Types come from the
 * repo's own restatement of the RL-Glue 3.04 layout. */
#include "../../../include/rlglue_types.h"
