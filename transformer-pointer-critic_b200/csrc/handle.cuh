// Library handle: per-device state shared by all entry points (no global mutable state besides the one-time
// driver entry-point lookup). Not thread-safe; different handles are independent.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

struct tpc_handle {
  int device = 0;
  int num_sms = 0;
  tpc::TmapCache* tmaps = nullptr;
};
