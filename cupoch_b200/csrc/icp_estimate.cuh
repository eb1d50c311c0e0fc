// icp_estimate.cuh -- standalone ComputeTransformation / ComputeRMSE kernels on an explicit correspondence list
// Part of the icp.cu translation unit (included there); split out for readability only.
#pragma once

// ---------------------------------------------------------------------------
// TransformationEstimation*::ComputeTransformation / ComputeRMSE on an explicit
// correspondence list (transformation_estimation.cu:92-350, generalized_icp.cu:112-183,
// colored_icp.cu:218-327, kabsch.cu:42-120): same rows and reduction as the fused kernel,
// one thread per correspondence, packed (original-order) source attributes.
// ---------------------------------------------------------------------------
struct EstArgs {
    TargetAttrs ta;
    const float *src_xyz, *src_nrm, *src_col, *src_cov;
    int src_cov_col_major;
    const int32_t *corr;
    unsigned n_corr;
    double *partials;
    double *total;     // [32]
    unsigned *ticket;
};
template <int KIND>
__global__ void __launch_bounds__(ICP_BLOCK) estimate_kernel(const __grid_constant__ EstArgs a) {
    __shared__ double s_rows[ICP_WARPS][32 * ROW_STRIDE];
    __shared__ double s_acc[ICP_WARPS][32];
    __shared__ unsigned s_last;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const unsigned c = blockIdx.x * ICP_BLOCK + threadIdx.x;
    const bool valid = c < a.n_corr;
    constexpr int NROWS = (KIND == CPHB_EST_COLORED_ICP) ? 2 : (KIND == CPHB_EST_GENERALIZED_ICP) ? 3 : 1;
    float J[NROWS][6], r[NROWS];
#pragma unroll
    for (int q = 0; q < NROWS; ++q) {
        r[q] = 0.f;
#pragma unroll
        for (int k = 0; k < 6; ++k) J[q][k] = 0.f;
    }
    float term = 0.f;  // per-kind ComputeRMSE summand
    if (valid) {
        const size_t i = (size_t)a.corr[2 * (size_t)c];
        const unsigned j = (unsigned)a.corr[2 * (size_t)c + 1];
        const float vs[3] = {a.src_xyz[3 * i], a.src_xyz[3 * i + 1], a.src_xyz[3 * i + 2]};
        float sn[3] = {0.f, 0.f, 0.f}, Cs[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        float4 cs4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (KIND == CPHB_EST_SYMMETRIC && a.src_nrm) { sn[0] = a.src_nrm[3 * i]; sn[1] = a.src_nrm[3 * i + 1]; sn[2] = a.src_nrm[3 * i + 2]; }
        if (KIND == CPHB_EST_COLORED_ICP && a.src_col) cs4 = make_float4(a.src_col[3 * i], a.src_col[3 * i + 1], a.src_col[3 * i + 2], 0.f);
        if (KIND == CPHB_EST_GENERALIZED_ICP && a.src_cov)
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int q = 0; q < 3; ++q) Cs[3 * p + q] = a.src_cov[9 * i + (a.src_cov_col_major ? 3 * q + p : 3 * p + q)];
        build_rows<KIND, NROWS>(a.ta, vs[0], vs[1], vs[2], sn, cs4, Cs, j, J, r);
        drop_nonfinite_rows<NROWS>(J, r);
        const float vt[3] = {a.ta.tgt_xyz[3 * (size_t)j], a.ta.tgt_xyz[3 * (size_t)j + 1], a.ta.tgt_xyz[3 * (size_t)j + 2]};
        if (KIND == CPHB_EST_POINT_TO_POINT) {
            term = dist2(vs[0], vs[1], vs[2], vt[0], vt[1], vt[2]);  // (lhs - rhs).squaredNorm()
        } else if (KIND == CPHB_EST_SYMMETRIC) {
            const float e = r[0] * r[0];  // ComputeErrorUsingNormals returns the squared residual ...
            term = e * e;                 // ... which the caller squares again (transformation_estimation.cu:283-286)
        } else if (KIND == CPHB_EST_GENERALIZED_ICP) {
            // d^T W d with W = sqrt((Ct+Cs)^-1)  (generalized_icp.cu:121-130): rows hold W_i and r_i = W_i . d
            const float d[3] = {vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2]};
            term = dot3(d[0], d[1], d[2], r[0], r[1 % NROWS], r[2 % NROWS]);
        }
    }
    double *rows = s_rows[warp];
    const unsigned char(*pair)[2] = (KIND == CPHB_EST_POINT_TO_POINT) ? c_pair_p2p : c_pair_jtj;
    const int ca = pair[lane][0], cb = pair[lane][1];
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < NROWS; ++q) {
        double *my = rows + lane * ROW_STRIDE;
#pragma unroll
        for (int k = 0; k < 6; ++k) my[k] = (double)J[q][k];
        my[6] = (double)r[q];
        my[7] = (q == 0 && valid) ? (double)term : 0.0;
        my[8] = (q == 0 && valid) ? 1.0 : 0.0;
        __syncwarp();
#pragma unroll 8
        for (int t = 0; t < 32; ++t) acc = fma(rows[t * ROW_STRIDE + ca], rows[t * ROW_STRIDE + cb], acc);
        __syncwarp();
    }
    {
        const unsigned live = (KIND == CPHB_EST_POINT_TO_POINT) ? c_live_p2p : c_live_jtj;
        if (!((live >> lane) & 1u)) acc = 0.0;
    }
    s_acc[warp][lane] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < ICP_WARPS; ++k) t += s_acc[k][threadIdx.x];
        a.partials[(size_t)blockIdx.x * 32 + threadIdx.x] = t;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(a.ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        const int col = threadIdx.x & 31, g = threadIdx.x >> 5;
        double t = 0.0;
        for (unsigned b = g; b < gridDim.x; b += ICP_WARPS) t += __ldcg(&a.partials[(size_t)b * 32 + col]);
        s_acc[g][col] = t;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < ICP_WARPS; ++k) t += s_acc[k][threadIdx.x];
        a.total[threadIdx.x] = t;
    }
    if (threadIdx.x == 0) *a.ticket = 0;
}
// sums -> 4x4 (one thread)
template <int KIND>
__global__ void estimate_solve_kernel(const double *S, unsigned long long n_model, float det_thresh, int have, float *T_out) {
    if (threadIdx.x != 0) return;
    float T[16];
    identity4(T);
    if (S[29] > 0 && have) {
        if (KIND == CPHB_EST_POINT_TO_POINT) kabsch_from_sums(S, n_model, T);
        else {
            bool ok = solve_jtj(S, (KIND == CPHB_EST_GENERALIZED_ICP) ? -1.f : det_thresh, T);
            if (ok && KIND == CPHB_EST_SYMMETRIC) {
                double R[9], R2[9];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) R[3 * i + j] = (double)T[4 * i + j];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j)
                        R2[3 * i + j] = R[3 * i] * R[j] + R[3 * i + 1] * R[3 + j] + R[3 * i + 2] * R[6 + j];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) T[4 * i + j] = (float)R2[3 * i + j];
            }
        }
    }
    for (int i = 0; i < 16; ++i) T_out[i] = T[i];
}

