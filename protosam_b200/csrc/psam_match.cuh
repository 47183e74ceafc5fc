// Shared declarations of the two variants of kernel 2 (fused query x prototype match).
#pragma once

#include "psam_common.cuh"

namespace psam {

struct MatchParams {
    const float* qry;
    int64_t slice_stride, row_stride;
    int Q, HW, C;
    const float* protos;
    int cap_rows;
    const int32_t* counts;
    const int32_t* eff_modes;
    int nsets;
    float* scores;
    float* assign;
    float* sims;
    int32_t* status;
};

// exact-fp32 CUDA-core variant (psam_match_simt.cu)
int launch_match_simt(const MatchParams& p, cudaStream_t stream);

// tcgen05 tensor-core variant (psam_match_tc.cu)
bool match_tc_supported(int Q, int HW, int C, int nsets, int cap_rows, bool want_sims);
// fused = the GEMM converts the fp32 query rows itself (no packed query image, no separate pass over the query)
// the fused variant additionally needs dense slices (one 2-D tensor map over all query rows)
bool match_ts_supported(const MatchParams& p);
int match_reserve_sms(int n);
// auto (algo 0): fused for narrow prototype tables, packed operands for wide ones (see the definition)
bool match_ts_preferred(const MatchParams& p);
size_t match_tc_workspace(int Q, int HW, int C, int nsets, int cap_rows, bool fused);
int launch_match_tc(const MatchParams& p, void* workspace, size_t workspace_bytes, bool fused, cudaStream_t stream);

}  // namespace psam
