// icp_rows.cuh -- estimator rows (J, r) of one correspondence for the five estimation methods
// Part of the icp.cu translation unit (included there, in this order: icp_types, icp_solve, icp_rows); split out
// for readability only -- the arithmetic contract and the reference citations are stated in icp.cu.
#pragma once

// ===========================================================================
// estimator rows for one correspondence (source point vs, target index j)
// ===========================================================================
struct TargetAttrs {
    const float *tgt_xyz, *tgt_nrm, *tgt_col, *tgt_grad, *tgt_cov;
    int tgt_cov_col_major;
    float sg, sp;
    bool src_nrm, src_col, src_cov;  // source attribute present
};
// The target-side values one row needs, loaded either from the caller's arrays by original index (standalone
// ComputeTransformation / ComputeRMSE, icp_estimate.cuh) or from the ICP context's private copies in INDEX order
// (icp_kernels.cuh).  `ok`: every attribute the estimator needs is present on both clouds (else the row stays zero).
struct TgtVals {
    float vt[3];  // matched target point
    float nt[3];  // its normal
    float gt[3];  // its colour gradient (Colored)
    float it;     // its intensity (Colored): intensity(colour), colored_icp.cu:176-181
    float ct[9];  // its covariance, row-major (GICP)
    bool ok;
};
template <int KIND>
__device__ __forceinline__ void load_tgt_orig(const TargetAttrs &a, unsigned j, TgtVals &t) {
    const size_t j3 = 3 * (size_t)j;
    t.vt[0] = a.tgt_xyz[j3]; t.vt[1] = a.tgt_xyz[j3 + 1]; t.vt[2] = a.tgt_xyz[j3 + 2];
    t.ok = true;
    if (KIND == CPHB_EST_POINT_TO_PLANE) t.ok = a.tgt_nrm != nullptr;
    if (KIND == CPHB_EST_SYMMETRIC) t.ok = a.tgt_nrm && a.src_nrm;
    if (KIND == CPHB_EST_COLORED_ICP) t.ok = a.tgt_nrm && a.tgt_col && a.src_col && a.tgt_grad;
    if (KIND == CPHB_EST_GENERALIZED_ICP) t.ok = a.tgt_cov && a.src_cov;
    if (!t.ok) return;
    if (KIND == CPHB_EST_POINT_TO_PLANE || KIND == CPHB_EST_SYMMETRIC || KIND == CPHB_EST_COLORED_ICP) {
        t.nt[0] = a.tgt_nrm[j3]; t.nt[1] = a.tgt_nrm[j3 + 1]; t.nt[2] = a.tgt_nrm[j3 + 2];
    }
    if (KIND == CPHB_EST_COLORED_ICP) {
        t.gt[0] = a.tgt_grad[j3]; t.gt[1] = a.tgt_grad[j3 + 1]; t.gt[2] = a.tgt_grad[j3 + 2];
        t.it = intensity(a.tgt_col[j3], a.tgt_col[j3 + 1], a.tgt_col[j3 + 2]);
    }
    if (KIND == CPHB_EST_GENERALIZED_ICP) {
        const float *ct = a.tgt_cov + 9 * (size_t)j;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) t.ct[3 * p + q] = ct[a.tgt_cov_col_major ? (3 * q + p) : (3 * p + q)];
    }
}

// rows (J, r) of one correspondence from the source-side values and the loaded target-side values
template <int KIND, int NROWS>
__device__ __forceinline__ void build_rows_vals(const float sg, const float sp, const float s_x, const float s_y,
                                                const float s_z, const float *sn, const float4 cs_in, const float *Cs,
                                                const TgtVals &t, float (&J)[NROWS][6], float (&r)[NROWS]) {
        const float vs[3] = {s_x, s_y, s_z};
        const float *vt = t.vt;
        if (KIND == CPHB_EST_POINT_TO_POINT) {
            J[0][0] = vs[0]; J[0][1] = vs[1]; J[0][2] = vs[2];
            J[0][3] = vt[0]; J[0][4] = vt[1]; J[0][5] = vt[2];
        } else if (KIND == CPHB_EST_POINT_TO_PLANE) {  // transformation_estimation.cu:34-56
            if (t.ok) {
                const float *nt = t.nt;
                r[0] = dot3(vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2], nt[0], nt[1], nt[2]);
                cross3(vs, nt, J[0]);
                J[0][3] = nt[0]; J[0][4] = nt[1]; J[0][5] = nt[2];
            }
        } else if (KIND == CPHB_EST_SYMMETRIC) {  // transformation_estimation.cu:58-90
            if (t.ok) {
                const float nn[3] = {sn[0] + t.nt[0], sn[1] + t.nt[1], sn[2] + t.nt[2]};
                const float sm[3] = {vs[0] + vt[0], vs[1] + vt[1], vs[2] + vt[2]};
                r[0] = dot3(vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2], nn[0], nn[1], nn[2]);
                cross3(sm, nn, J[0]);
                J[0][3] = nn[0]; J[0][4] = nn[1]; J[0][5] = nn[2];
            }
        } else if (KIND == CPHB_EST_COLORED_ICP) {  // colored_icp.cu:150-216
            if (t.ok) {
                const float *nt = t.nt, *gt = t.gt;
                const float4 cs = cs_in;
                const float d[3] = {vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2]};
                const float dn = dot3(d[0], d[1], d[2], nt[0], nt[1], nt[2]);
                float cr[3];
                cross3(vs, nt, cr);
#pragma unroll
                for (int c = 0; c < 3; ++c) { J[0][c] = sg * cr[c]; J[0][3 + c] = sg * nt[c]; }
                r[0] = sg * dn;
                float pd[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) pd[c] = __fmaf_rn(-dn, nt[c], vs[c]) - vt[c];
                const float is = intensity(cs.x, cs.y, cs.z);
                const float it = t.it;
                const float is0 = dot3(gt[0], gt[1], gt[2], pd[0], pd[1], pd[2]) + it;
                float M[9];
#pragma unroll
                for (int p = 0; p < 3; ++p)
#pragma unroll
                    for (int q = 0; q < 3; ++q)
                        M[3 * p + q] = (p == q) ? (float)(1.0 - (double)(nt[p] * nt[p]))
                                                : (-nt[p < q ? p : q]) * nt[p < q ? q : p];
                float gm[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) gm[q] = dot3(-gt[0], -gt[1], -gt[2], M[q], M[3 + q], M[6 + q]);
                cross3(vs, gm, cr);
#pragma unroll
                for (int c = 0; c < 3; ++c) { J[1 % NROWS][c] = sp * cr[c]; J[1 % NROWS][3 + c] = sp * gm[c]; }
                r[1 % NROWS] = sp * (is - is0);
            }
        } else if (KIND == CPHB_EST_GENERALIZED_ICP) {  // generalized_icp.cu:63-105
            if (t.ok) {
                float Mx[9], Mi[9], W[9];
#pragma unroll
                for (int p = 0; p < 9; ++p) Mx[p] = t.ct[p] + Cs[p];
                inverse3x3(Mx, Mi);
                sqrt_matrix3x3(Mi, W);
                const float d[3] = {vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2]};
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float *wr = W + 3 * q;
                    J[q % NROWS][0] = __fmaf_rn(wr[2], vs[1], -(wr[1] * vs[2]));
                    J[q % NROWS][1] = __fmaf_rn(wr[2], -vs[0], wr[0] * vs[2]);
                    J[q % NROWS][2] = __fmaf_rn(wr[1], vs[0], -(wr[0] * vs[1]));
                    J[q % NROWS][3] = wr[0]; J[q % NROWS][4] = wr[1]; J[q % NROWS][5] = wr[2];
                    r[q % NROWS] = dot3(wr[0], wr[1], wr[2], d[0], d[1], d[2]);
                }
            }
        }
}
template <int KIND, int NROWS>
__device__ __forceinline__ void build_rows(const TargetAttrs &a, const float s_x, const float s_y, const float s_z,
                                           const float *sn, const float4 cs_in, const float *Cs, unsigned j,
                                           float (&J)[NROWS][6], float (&r)[NROWS]) {
    TgtVals t;
    load_tgt_orig<KIND>(a, j, t);
    build_rows_vals<KIND, NROWS>(a.sg, a.sp, s_x, s_y, s_z, sn, cs_in, Cs, t, J, r);
}

// Deliberate deviation from the reference (parity hazard 8): a row with a non-finite entry is
// dropped instead of poisoning the whole sum.  The reference's FastEigen3x3 computes x/|x| (eigenvalue.inl:28)
// which is 0/0 when an off-diagonal projection vanishes exactly; with millions of GICP rows per iteration that
// happens, and the reference then returns an all-NaN transformation.
template <int NROWS>
__device__ __forceinline__ void drop_nonfinite_rows(float (&J)[NROWS][6], float (&r)[NROWS]) {
#pragma unroll
    for (int q = 0; q < NROWS; ++q) {
        float s = r[q];
#pragma unroll
        for (int c = 0; c < 6; ++c) s += J[q][c];  // NaN/inf propagate into s
        if (!(fabsf(s) <= FLT_MAX)) {  // NaN or inf
            r[q] = 0.f;
#pragma unroll
            for (int c = 0; c < 6; ++c) J[q][c] = 0.f;
        }
    }
}

