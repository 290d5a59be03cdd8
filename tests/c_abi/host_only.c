/* TEST: include/art_b200.h is plain C, and a C program can link libart_b200.so and call the entry
 * points that need no GPU (version, struct sizes, pose -> rotation, detector construction). */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "art_b200.h"

int main(void) {
  int32_t sizes[6];
  double rot[9];
  const double n[3] = {0.0, 0.0, 1.0}, m[3] = {1.0, 0.0, 0.0};
  const double centre[3] = {0.0, 0.0, 10.0}, normal[3] = {0.0, 0.0, -2.0}, ref[3] = {0.0, 0.0, 0.0};
  ArtDetector det;
  int i, bad = 0;
  if (art_version() != ART_B200_VERSION) return 1;
  if (art_abi_sizes(sizes) != ART_OK) return 2;
  if (sizes[0] != (int32_t)sizeof(ArtElementDesc) || sizes[1] != (int32_t)sizeof(ArtZernikeDesc) ||
      sizes[2] != (int32_t)sizeof(ArtBundleView) || sizes[3] != (int32_t)sizeof(ArtDetector) ||
      sizes[4] != (int32_t)sizeof(ArtGridMapDesc) || sizes[5] != (int32_t)sizeof(ArtSourceDesc))
    return 3;
  if (art_element_rotation(n, m, rot) != ART_OK) return 4;
  for (i = 0; i < 9; ++i) bad += fabs(rot[i] - ((i % 4 == 0) ? 1.0 : 0.0)) > 1e-15;
  if (bad) return 5;
  if (art_detector_make(centre, normal, ref, 0.0, &det) != ART_OK) return 6;
  if (fabs(det.normal[2] + 1.0) > 1e-15 || fabs(det.cvec[2] - 1.0) > 1e-15) return 7;
  if (art_element_rotation(NULL, m, rot) != ART_E_INVALID || strlen(art_last_error()) == 0) return 8;
  printf("c-abi ok %d %d %d %d %d\n", sizes[0], sizes[1], sizes[2], sizes[3], sizes[4]);
  return 0;
}
