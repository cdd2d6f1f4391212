/*
 * me_io.c -- 8-bit luma .yuv I/O + wall clock of the drop-in host layer (plain C).
 * Behaviour of reference src/common/utils.c:23-27 (getTimeStamp), :61-73
 * (yuvReadFrame: first numElems bytes widened uint8 -> int, 1 ok / 0 fail) and
 * :75-92 (yuvWriteFrame: int narrowed by C cast to uint8).
 */
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "me_common.h"

double getTimeStamp(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (double)tv.tv_sec + (double)tv.tv_usec / 1000000;
}

int yuvReadFrameU8(const char *file_name, unsigned char *target_buffer, int numElems) {
  FILE *f = fopen(file_name, "rb");
  if (!f) {
    printf("yuvOpenInputFile: Could not open the file %s\n", file_name);
    return 0;
  }
  size_t got = fread(target_buffer, 1, (size_t)numElems, f);
  fclose(f);
  if (got != (size_t)numElems) {
    printf("yuvReadFrame: The read was failed!\n");
    return 0;
  }
  return 1;
}

int yuvReadFrame(const char *file_name, int *const target_buffer, int numElems) {
  uint8_t *bytes = (uint8_t *)malloc((size_t)numElems);
  if (!bytes) return 0;
  int ok = yuvReadFrameU8(file_name, bytes, numElems);
  if (ok)
    for (int i = 0; i < numElems; i++) target_buffer[i] = bytes[i];
  free(bytes);
  return ok;
}

int yuvWriteFrame(const char *file_name, const int *const data_buffer, int numElems) {
  uint8_t *bytes = (uint8_t *)malloc((size_t)numElems);
  if (!bytes) return 0;
  for (int i = 0; i < numElems; i++) bytes[i] = (uint8_t)data_buffer[i];
  FILE *f = fopen(file_name, "wb");
  if (!f) {
    printf("yuvWriteToFile: Could not open the file %s (%s)\n", file_name, strerror(errno));
    free(bytes);
    return 0;
  }
  size_t put = fwrite(bytes, 1, (size_t)numElems, f);
  fclose(f);
  free(bytes);
  return put == (size_t)numElems;
}
