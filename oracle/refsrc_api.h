/* oracle/refsrc_api.h -- TEST INFRASTRUCTURE: flat views the reference-source harness (refsrc_harness.cpp) takes in
 * addition to the C-ABI structs of include/b200orb.h. */
#ifndef REFSRC_API_H_
#define REFSRC_API_H_
#include <stdint.h>
typedef struct {              /* a list of MapPoints as the searches read them */
  int n;
  const uint8_t* valid;       /* 0 = NULL pointer in the list */
  const uint8_t* bad;         /* isBad(); NULL = none */
  const float* xw;            /* n x 3, GetWorldPos() */
  const float* normal;        /* n x 3, GetNormal(); NULL = derived from the proto frame */
  const float* min_dist;      /* mfMinDistance (GetMinDistanceInvariance() = 0.8f * it); NULL = 0 */
  const float* max_dist;      /* mfMaxDistance (GetMaxDistanceInvariance() = 1.2f * it); NULL = 1e9 */
  const uint8_t* desc;        /* n x 32, GetDescriptor() */
  const int32_t* obs;         /* Observations(); NULL = 1 */
} RefMapPoints;
#endif
