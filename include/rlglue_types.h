/*
 * RL-Glue 3.04 plain-C data types used by the environment plugin ABI.
 *
 * The reference environment (rlglue/environment/SwimmerEnvironment.h:16-18) includes
 * <rlglue/Environment_common.h> and <rlglue/utils/C/RLStruct_util.h> from an RL-Glue
 * installation that is not vendored in the reference tree.  This header restates the
 * public layout of those types so that (a) the env_* shim in this repo
 * (csrc/rlglue_env_shim.cpp) exports the same symbols with the same ownership rules and
 * (b) the reference C++ file can be compiled as an oracle (oracle/Makefile) without an
 * RL-Glue installation.
 */
#ifndef SWM_RLGLUE_TYPES_H
#define SWM_RLGLUE_TYPES_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  unsigned int numInts;
  unsigned int numDoubles;
  unsigned int numChars;
  int* intArray;
  double* doubleArray;
  char* charArray;
} rl_abstract_type_t;

typedef rl_abstract_type_t observation_t;
typedef rl_abstract_type_t action_t;

typedef struct {
  double reward;
  const observation_t* observation;
  int terminal;
} reward_observation_terminal_t;

/* rlglue/utils/C/RLStruct_util.h (allocateRLStruct / clearRLStruct) is deliberately NOT declared or
 * exported here: a real RL-Glue build links -lrlutils, which defines those symbols, and a second default-
 * visibility definition in libswimmer_ars.so would interpose on it.  The shim allocates its two observation
 * structs with private helpers (csrc/rlglue_env_shim.cu). */

#ifdef __cplusplus
}
#endif
#endif
