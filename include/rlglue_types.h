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

/* RLStruct_util.h equivalents: the callee owns the arrays (SwimmerEnvironment.cpp:20-21,72-73). */
#if defined(__GNUC__)
__attribute__((visibility("default")))
#endif
void allocateRLStruct(rl_abstract_type_t* dst, unsigned int numInts, unsigned int numDoubles,
                      unsigned int numChars);
#if defined(__GNUC__)
__attribute__((visibility("default")))
#endif
void clearRLStruct(rl_abstract_type_t* dst);

#ifdef __cplusplus
}
#endif
#endif
