// internal kernel parameter block of the tcgen05 conv (host fills it from TpzTcConvArgs)
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "../../include/topaz_b200.h"

struct TcConvParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  int nsrc;
  int org[2][3];
  int N, Do, Ho, Wo, Co;
  int TW, TH, tiles_x, tiles_y, num_tiles;
  int nkb, stages;
  const float* bias;
  float neg_slope;
  const __half* res;
  const float* res_scale;
  int res_ld, res_D, res_H, res_W, res_org[3];
  __half* out;
  int out_ld, out_coff;
  const float* dot_w;
  float dot_b;
  float* dot_out;
  const float* dot_affine;
  const float* oscale;
  const float* range;
  int out_lo;
  TcKBlock kb[TPZ_TC_MAX_KB];
};
