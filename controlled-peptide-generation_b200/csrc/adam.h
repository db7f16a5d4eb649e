#pragma once
#include "cpg_common.cuh"

namespace cpg {

struct AdamStep { float step_size; float bc2_sqrt; };     // lr / (1 - b1^t),  sqrt(1 - b2^t)
struct AdamHyper {
    float beta1, beta2, eps;
    AdamStep single;        // step t of the non-duplicated tensors
    AdamStep dup_first;     // steps 2t-1 and 2t of the duplicated embedding
    AdamStep dup_second;
};

int adam_norm_parts(int64_t n, int sm_count);
// partial sums of squares, then one kernel that finishes the norm, clips and applies Adam (norm_out / coef_out nullable)
void launch_clip_adam_fused(cudaStream_t s, float* p, float* g, float* m, float* v, int64_t n, int64_t dup_off,
                            int64_t dup_n, float max_norm, float* part, float* norm_out, float* coef_out,
                            const AdamHyper& h, int sm_count, unsigned* bar = nullptr);

}  // namespace cpg
