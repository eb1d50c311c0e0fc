// icp_solve.cuh -- the device-side epilogue of one iteration: 6x6 determinant / LDLT / solve, Kabsch, convergence test, pose composition
// Part of the icp.cu translation unit (included there, in this order: icp_types, icp_solve, icp_rows); split out
// for readability only -- the arithmetic contract and the reference citations are stated in icp.cu.
#pragma once

// ===========================================================================
// finalize: 6x6 solve / Kabsch / convergence (one thread)
// ===========================================================================
__device__ float det6_partial_piv(const float *A_in) {
    float A[36];
    for (int i = 0; i < 36; ++i) A[i] = A_in[i];
    float det = 1.f;
    for (int k = 0; k < 6; ++k) {
        int p = k;
        float best = fabsf(A[6 * k + k]);
        for (int i = k + 1; i < 6; ++i)
            if (fabsf(A[6 * i + k]) > best) { best = fabsf(A[6 * i + k]); p = i; }
        if (best == 0.f) return 0.f;
        if (p != k) {
            for (int j = 0; j < 6; ++j) { float t = A[6 * k + j]; A[6 * k + j] = A[6 * p + j]; A[6 * p + j] = t; }
            det = -det;
        }
        float piv = A[6 * k + k];
        det = det * piv;
        for (int i = k + 1; i < 6; ++i) {
            float f = A[6 * i + k] / piv;
            for (int j = k + 1; j < 6; ++j) A[6 * i + j] = A[6 * i + j] - f * A[6 * k + j];
        }
    }
    return det;
}
// Eigen LDLT (diagonal pivoting, lower) + solve; eigen.cu:103 A.ldlt().solve(b)
__device__ void ldlt6_solve(const float *A_in, const float *b, float *x) {
    float A[36];
    for (int i = 0; i < 36; ++i) A[i] = A_in[i];
    int tr[6];
    for (int k = 0; k < 6; ++k) {
        int ib = k;
        float big = fabsf(A[6 * k + k]);
        for (int i = k + 1; i < 6; ++i)
            if (fabsf(A[6 * i + i]) > big) { big = fabsf(A[6 * i + i]); ib = i; }
        tr[k] = ib;
        if (ib != k) {
            for (int j = 0; j < k; ++j) { float t = A[6 * k + j]; A[6 * k + j] = A[6 * ib + j]; A[6 * ib + j] = t; }
            for (int i = ib + 1; i < 6; ++i) { float t = A[6 * i + k]; A[6 * i + k] = A[6 * i + ib]; A[6 * i + ib] = t; }
            { float t = A[6 * k + k]; A[6 * k + k] = A[6 * ib + ib]; A[6 * ib + ib] = t; }
            for (int i = k + 1; i < ib; ++i) { float t = A[6 * i + k]; A[6 * i + k] = A[6 * ib + i]; A[6 * ib + i] = t; }
        }
        float temp[6];
        if (k > 0) {
            for (int j = 0; j < k; ++j) temp[j] = A[6 * j + j] * A[6 * k + j];
            float s = 0.f;
            for (int j = 0; j < k; ++j) s = s + A[6 * k + j] * temp[j];
            A[6 * k + k] = A[6 * k + k] - s;
            for (int i = k + 1; i < 6; ++i) {
                float s2 = 0.f;
                for (int j = 0; j < k; ++j) s2 = s2 + A[6 * i + j] * temp[j];
                A[6 * i + k] = A[6 * i + k] - s2;
            }
        }
        float akk = A[6 * k + k];
        if (fabsf(akk) > 0.f)
            for (int i = k + 1; i < 6; ++i) A[6 * i + k] = A[6 * i + k] / akk;
    }
    float y[6];
    for (int i = 0; i < 6; ++i) y[i] = b[i];
    for (int k = 0; k < 6; ++k)
        if (tr[k] != k) { float t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
    for (int i = 0; i < 6; ++i) {
        float s = y[i];
        for (int j = 0; j < i; ++j) s = s - A[6 * i + j] * y[j];
        y[i] = s;
    }
    for (int i = 0; i < 6; ++i) y[i] = (fabsf(A[6 * i + i]) > FLT_MIN) ? y[i] / A[6 * i + i] : 0.f;
    for (int i = 5; i >= 0; --i) {
        float s = y[i];
        for (int j = i + 1; j < 6; ++j) s = s - A[6 * j + i] * y[j];
        y[i] = s;
    }
    for (int k = 5; k >= 0; --k)
        if (tr[k] != k) { float t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
    for (int i = 0; i < 6; ++i) x[i] = y[i];
}
__device__ void identity4(float *T) {
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
}
__device__ void se3_exp(const float *x, float *T) {  // eigen.cu:28-50
    identity4(T);
    T[3] = x[3]; T[7] = x[4]; T[11] = x[5];
    float th = sqrtf((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]);
    if (th == 0) return;
    float w0 = x[0] / th, w1 = x[1] / th, w2 = x[2] / th;
    float c = (float)cos((double)th), s = (float)sin((double)th);
    float oc = 1 - c;
    T[0] = c + w0 * w0 * oc;
    T[1] = w0 * w1 * oc - w2 * s;
    T[2] = w1 * s + w0 * w2 * oc;
    T[4] = w2 * s + w0 * w1 * oc;
    T[5] = c + w1 * w1 * oc;
    T[6] = -w0 * s + w1 * w2 * oc;
    T[8] = -w1 * s + w0 * w2 * oc;
    T[9] = w0 * s + w1 * w2 * oc;
    T[10] = c + w2 * w2 * oc;
}
__device__ bool solve_jtj(const double *S, float det_thresh, float *T) {
    float A[36], b[6], x[6];
    int p = 0;
    for (int a = 0; a < 6; ++a)
        for (int c = a; c < 6; ++c) { float v = (float)S[p++]; A[6 * a + c] = v; A[6 * c + a] = v; }
    for (int a = 0; a < 6; ++a) b[a] = -(float)S[21 + a];
    identity4(T);
    if (det_thresh > 0) {  // eigen.cu:88-100
        float det = det6_partial_piv(A);
        if (fabsf(det) < det_thresh || isnan(det) || isinf(det)) return false;
    }
    ldlt6_solve(A, b, x);
    se3_exp(x, T);
    return true;
}
// ---------------------------------------------------------------------------
// Warp-parallel versions of the 6x6 determinant / LDLT / solve above (one warp, matrices in shared memory).
// Every scalar is produced by exactly the same operations in the same order as in the sequential code
// (and the oracle); only independent rows / elements are spread over lanes, which shortens the dependent
// chain of the per-iteration epilogue from ~15 us to a few us.
// ---------------------------------------------------------------------------
struct SolveSmem {
    float A[36], L[36], b[8], y[8], x[8], temp[8];
    float T[16];
    int piv, ok;
};
__device__ float det6_warp(SolveSmem &m) {  // on m.L (copy of A), result broadcast to all lanes
    const int lane = lane_id();
    float det = 1.f;
    for (int k = 0; k < 6; ++k) {
        if (lane == 0) {
            int p = k;
            float best = fabsf(m.L[6 * k + k]);
            for (int i = k + 1; i < 6; ++i)
                if (fabsf(m.L[6 * i + k]) > best) { best = fabsf(m.L[6 * i + k]); p = i; }
            m.piv = (best == 0.f) ? -1 : p;
        }
        __syncwarp();
        const int p = m.piv;
        if (p < 0) return 0.f;
        if (p != k) {
            if (lane < 6) { float t = m.L[6 * k + lane]; m.L[6 * k + lane] = m.L[6 * p + lane]; m.L[6 * p + lane] = t; }
            det = -det;
        }
        __syncwarp();
        const float piv = m.L[6 * k + k];
        det = det * piv;
        const int nr = 5 - k;  // rows/cols below/right of the pivot
        if (lane < nr * nr) {
            const int i = k + 1 + lane / nr, j = k + 1 + lane % nr;
            const float f = m.L[6 * i + k] / piv;
            m.L[6 * i + j] = m.L[6 * i + j] - f * m.L[6 * k + j];
        }
        __syncwarp();
    }
    return det;
}
__device__ void ldlt6_solve_warp(SolveSmem &m) {  // factors m.L (copy of A) in place, solves into m.x
    const int lane = lane_id();
    int tr[6];
    for (int k = 0; k < 6; ++k) {
        if (lane == 0) {
            int ib = k;
            float big = fabsf(m.L[6 * k + k]);
            for (int i = k + 1; i < 6; ++i)
                if (fabsf(m.L[6 * i + i]) > big) { big = fabsf(m.L[6 * i + i]); ib = i; }
            m.piv = ib;
        }
        __syncwarp();
        const int ib = m.piv;
        tr[k] = ib;
        if (ib != k) {  // symmetric swap of rows/cols k and ib in the lower triangle (disjoint element sets)
            if (lane < k) { float t = m.L[6 * k + lane]; m.L[6 * k + lane] = m.L[6 * ib + lane]; m.L[6 * ib + lane] = t; }
            if (lane > ib && lane < 6) { float t = m.L[6 * lane + k]; m.L[6 * lane + k] = m.L[6 * lane + ib]; m.L[6 * lane + ib] = t; }
            if (lane == 31) { float t = m.L[6 * k + k]; m.L[6 * k + k] = m.L[6 * ib + ib]; m.L[6 * ib + ib] = t; }
            if (lane > k && lane < ib) { float t = m.L[6 * lane + k]; m.L[6 * lane + k] = m.L[6 * ib + lane]; m.L[6 * ib + lane] = t; }
        }
        __syncwarp();
        if (k > 0) {
            if (lane < k) m.temp[lane] = m.L[6 * lane + lane] * m.L[6 * k + lane];
            __syncwarp();
            if (lane == 0) {
                float sacc = 0.f;
                for (int j = 0; j < k; ++j) sacc = sacc + m.L[6 * k + j] * m.temp[j];
                m.L[6 * k + k] = m.L[6 * k + k] - sacc;
            } else if (lane > k && lane < 6) {
                float s2 = 0.f;
                for (int j = 0; j < k; ++j) s2 = s2 + m.L[6 * lane + j] * m.temp[j];
                m.L[6 * lane + k] = m.L[6 * lane + k] - s2;
            }
            __syncwarp();
        }
        const float akk = m.L[6 * k + k];
        if (fabsf(akk) > 0.f && lane > k && lane < 6) m.L[6 * lane + k] = m.L[6 * lane + k] / akk;
        __syncwarp();
    }
    if (lane == 0) {  // substitutions: short sequential chains
        float y[6];
        for (int i = 0; i < 6; ++i) y[i] = m.b[i];
        for (int k = 0; k < 6; ++k)
            if (tr[k] != k) { float t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
        for (int i = 0; i < 6; ++i) {
            float sacc = y[i];
            for (int j = 0; j < i; ++j) sacc = sacc - m.L[6 * i + j] * y[j];
            y[i] = sacc;
        }
        for (int i = 0; i < 6; ++i) y[i] = (fabsf(m.L[6 * i + i]) > FLT_MIN) ? y[i] / m.L[6 * i + i] : 0.f;
        for (int i = 5; i >= 0; --i) {
            float sacc = y[i];
            for (int j = i + 1; j < 6; ++j) sacc = sacc - m.L[6 * j + i] * y[j];
            y[i] = sacc;
        }
        for (int k = 5; k >= 0; --k)
            if (tr[k] != k) { float t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
        for (int i = 0; i < 6; ++i) m.x[i] = y[i];
    }
    __syncwarp();
}
// warp version of solve_jtj: result in m.T (all lanes may read after the call), returns success
__device__ bool solve_jtj_warp(const double *S, float det_thresh, SolveSmem &m) {
    const int lane = lane_id();
    if (lane == 0) {
        int p = 0;
        for (int a = 0; a < 6; ++a)
            for (int c = a; c < 6; ++c) { float v = (float)S[p++]; m.A[6 * a + c] = v; m.A[6 * c + a] = v; }
        for (int a = 0; a < 6; ++a) m.b[a] = -(float)S[21 + a];
        identity4(m.T);
    }
    __syncwarp();
    if (det_thresh > 0) {
        for (int e = lane; e < 36; e += 32) m.L[e] = m.A[e];
        __syncwarp();
        const float det = det6_warp(m);
        if (fabsf(det) < det_thresh || isnan(det) || isinf(det)) return false;
    }
    for (int e = lane; e < 36; e += 32) m.L[e] = m.A[e];
    __syncwarp();
    ldlt6_solve_warp(m);
    if (lane == 0) {
        float x[6];
        for (int i = 0; i < 6; ++i) x[i] = m.x[i];
        se3_exp(x, m.T);
    }
    __syncwarp();
    return true;
}

__device__ void matmul4(const float *A, const float *B, float *C) {  // registration.cu:159
    float R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            R[4 * i + j] = ((A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j]) + A[4 * i + 2] * B[8 + j]) + A[4 * i + 3] * B[12 + j];
    for (int i = 0; i < 16; ++i) C[i] = R[i];
}
// one-sided Jacobi SVD (double) -- stands in for Eigen::JacobiSVD<Matrix3f> (kabsch.cu:108-109);
// R = V diag(1,1,det(UV)) U^T is unique whatever the SVD's sign/ordering conventions.
__device__ void svd3(const double *A, double *U, double *s, double *V) {
    double B[9];
    for (int i = 0; i < 9; ++i) { B[i] = A[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double al = 0, be = 0, ga = 0;
                for (int i = 0; i < 3; ++i) {
                    al += B[3 * i + p] * B[3 * i + p];
                    be += B[3 * i + q] * B[3 * i + q];
                    ga += B[3 * i + p] * B[3 * i + q];
                }
                if (fabs(ga) <= 1e-300 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
                off += fabs(ga);
                double zeta = (be - al) / (2.0 * ga);
                double t = ((zeta >= 0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int i = 0; i < 3; ++i) {
                    double bp = B[3 * i + p], bq = B[3 * i + q];
                    B[3 * i + p] = c * bp - sn * bq;
                    B[3 * i + q] = sn * bp + c * bq;
                    double vp = V[3 * i + p], vq = V[3 * i + q];
                    V[3 * i + p] = c * vp - sn * vq;
                    V[3 * i + q] = sn * vp + c * vq;
                }
            }
        if (off == 0) break;
    }
    for (int j = 0; j < 3; ++j) {
        double nn = sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
        s[j] = nn;
        for (int i = 0; i < 3; ++i) U[3 * i + j] = (nn > 0) ? B[3 * i + j] / nn : 0.0;
    }
    for (int j = 0; j < 3; ++j)
        if (s[j] == 0) {
            int a = (j + 1) % 3, b = (j + 2) % 3;
            if (s[a] > 0 && s[b] > 0) {
                U[j] = U[3 + a] * U[6 + b] - U[6 + a] * U[3 + b];
                U[3 + j] = U[6 + a] * U[b] - U[a] * U[6 + b];
                U[6 + j] = U[a] * U[3 + b] - U[3 + a] * U[b];
            }
        }
}
// kabsch.cu:42-120 incl. the divide-by-model.size() quirk (:76-78,107)
__device__ void kabsch_from_sums(const double *S, unsigned long long n_model, float *T) {
    identity4(T);
    double C = S[29];
    float div = 1.0f / (float)n_model;
    float mc[3], tc[3];
    for (int a = 0; a < 3; ++a) { mc[a] = (float)S[a] * div; tc[a] = (float)S[3 + a] * div; }
    double H[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double h = S[6 + 3 * a + b] - (double)mc[a] * S[3 + b] - S[a] * (double)tc[b] + C * (double)mc[a] * (double)tc[b];
            H[3 * a + b] = (double)((float)h / (float)n_model);
        }
    double U[9], sv[3], V[9], UV[9];
    svd3(H, U, sv, V);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) UV[3 * i + j] = U[3 * i] * V[j] + U[3 * i + 1] * V[3 + j] + U[3 * i + 2] * V[6 + j];
    double dd = UV[0] * (UV[4] * UV[8] - UV[5] * UV[7]) - UV[1] * (UV[3] * UV[8] - UV[5] * UV[6]) +
                UV[2] * (UV[3] * UV[7] - UV[4] * UV[6]);
    double ss[3] = {1.0, 1.0, dd};
    float R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double r = 0;
            for (int k = 0; k < 3; ++k) r += V[3 * i + k] * ss[k] * U[3 * j + k];
            R[3 * i + j] = (float)r;
        }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
        T[4 * i + 3] = tc[i] - dot3(R[3 * i], R[3 * i + 1], R[3 * i + 2], mc[0], mc[1], mc[2]);
    }
}

// registration.cu:71-78,154-172 -- runs in ONE WARP (all 32 lanes call it) after the grid-wide sum.
template <int KIND>
__device__ void icp_finalize(const IcpArgs &a, IcpState *st, SolveSmem &m) {
    const int lane = lane_id();
    const double *S = st->total;
    int action = 0;  // 0 = nothing more, 1 = compute an update
    double cnt = 0.0;
    if (lane == 0) {
        cnt = S[29];
        float fit = 0.f, rmse = 0.f;
        if (cnt > 0) {
            fit = (float)cnt / (float)a.n_total;
            rmse = sqrtf((float)S[28] / (float)cnt);
        }
        const float pf = st->fitness, pr = st->rmse;
        st->fitness = fit;
        st->rmse = rmse;
        st->n_corr = (long long)cnt;
        if (a.step_mode) {
            action = 0;
        } else if (a.launch_idx > 0 && fabsf(pf - fit) < a.rel_fitness && fabsf(pr - rmse) < a.rel_rmse) {
            st->converged = 1;
            st->done = (a.corr_index && a.launch_idx < a.max_iter) ? 1 : 2;
        } else if (a.launch_idx >= a.max_iter) {
            st->done = 2;
        } else {
            action = 1;
        }
    }
    action = __shfl_sync(CPHB_FULL, action, 0);
    if (!action) return;
    const bool have_corr = __shfl_sync(CPHB_FULL, (int)(cnt > 0), 0) != 0;
    if (lane == 0) identity4(m.T);
    __syncwarp();
    if (have_corr) {
        if (KIND == CPHB_EST_POINT_TO_POINT) {
            if (lane == 0) kabsch_from_sums(S, a.n_total, m.T);
            __syncwarp();
        } else {
            bool have = true;
            if ((KIND == CPHB_EST_POINT_TO_PLANE || KIND == CPHB_EST_COLORED_ICP) && !a.tgt_nrm) have = false;
            if (KIND == CPHB_EST_SYMMETRIC && (!a.tgt_nrm || !a.src_nrm)) have = false;
            if (KIND == CPHB_EST_COLORED_ICP && (!a.has_tgt_col || !a.src_col)) have = false;
            if (KIND == CPHB_EST_GENERALIZED_ICP && (!a.tgt_cov || !a.src_cov)) have = false;
            if (have) {
                const float dt = (KIND == CPHB_EST_GENERALIZED_ICP) ? -1.f : a.det_thresh;
                const bool ok = solve_jtj_warp(S, dt, m);
                if (!ok && lane == 0) identity4(m.T);
                if (ok && KIND == CPHB_EST_SYMMETRIC && lane == 0) {  // transformation_estimation.cu:319-339
                    double R[9], R2[9];
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) R[3 * i + j] = (double)m.T[4 * i + j];
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j)
                            R2[3 * i + j] = R[3 * i] * R[j] + R[3 * i + 1] * R[3 + j] + R[3 * i + 2] * R[6 + j];
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) m.T[4 * i + j] = (float)R2[3 * i + j];
                }
                __syncwarp();
            }
        }
    }
    // transformation = update * transformation (registration.cu:159): one output element per lane
    float tn = 0.f;
    if (lane < 16) {
        const int i = lane >> 2, j = lane & 3;
        const float *A = m.T, *B = st->T;
        tn = ((A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j]) + A[4 * i + 2] * B[8 + j]) + A[4 * i + 3] * B[12 + j];
    }
    __syncwarp();
    if (lane < 16) {
        st->T[lane] = tn;
        st->U[lane] = m.T[lane];
    }
    if (lane == 0) {
        st->apply_u = 1;
        st->iterations += 1;
    }
    __syncwarp();
}

