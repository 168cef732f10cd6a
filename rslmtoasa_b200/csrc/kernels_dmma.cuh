// kernels_dmma.cuh -- FP64 tensor-core (DMMA m8n8k4) pipeline (kernel family 1).  STUB: filled in next.
#pragma once
#include "common.cuh"
#include <vector>
struct DmmaTiles { int ntiles = 0; };
static int dmma_configure() { return 0; }
static int dmma_build_tiles(DmmaTiles &, const std::vector<int32_t> &, const std::vector<int32_t> &, int, int, int) { return 0; }
static void dmma_free_tiles(DmmaTiles &) {}
static bool dmma_supported(const ApplyParams &) { return false; }
static int dmma_max_ctas(int sms) { return sms; }
static int dmma_parts_for(const DmmaTiles &, int sms, int) { return sms; }
static int dmma_launch_apply(DmmaTiles &, ApplyParams &, int, int, cudaStream_t, long long *) { return -1; }
