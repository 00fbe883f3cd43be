// Shared host-side helpers of the C-ABI library (error string, launch counter).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/allrank_b200.h"

void arb_set_error(const char* msg);
void arb_count_launch(int n = 1);

// loss = sum(val)/sum(cnt), grad *= 1/sum(cnt); an all-zero count gives loss 0 and zero grad (slate_kernels.cu)
int arb_finalize_mean_over_count(const float* val, const float* cnt, int B, float* loss, float* grad, size_t n_grad,
                                 cudaStream_t st);
