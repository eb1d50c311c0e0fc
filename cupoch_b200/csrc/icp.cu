// icp.cu -- fused ICP iteration (one launch per iteration) + device-side 6x6
// solve / Kabsch / convergence logic + the RegistrationICP driver.
//
// Replaces, per iteration, the reference's ~13 thrust launches and >= 2 host
// synchronisations:
//   pcd.Transform(update)                      registration.cu:160, geometry_utils.cu:34-54,257-265
//   kdtree.SearchRadius(.., 1, ..)             registration.cu:47, kdtree_cuda_3d_index.cu:52-154
//   error2 / correspondence list               registration.cu:50-69
//   ComputeJTJandJTr / Kabsch sums             eigen.inl:92-145, kabsch.cu:48-104
//   SolveJacobianSystemAndObtainExtrinsicMatrix eigen.cu:75-122 (host Eigen LDLT in the reference)
//   transformation = update * transformation   registration.cu:159
//   convergence test                           registration.cu:165-170
//
// Arithmetic contract: per-row float32 arithmetic in the stated
// order with explicit FMAs (this file is compiled with -fmad=false, so only
// the __fmaf_rn/fma calls below fuse); sums of exact float*float products in
// float64, reduced in a fixed order (warp -> block -> grid), rounded once to
// float32; the 6x6 LDLT, se(3) exponential and 4x4 composition use unfused
// float32 like the reference's host code.
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>

#include "cphb_internal.cuh"
#include <vector>
#include "cphb_eigen3.cuh"

#define ICP_BLOCK 256
#define ICP_WARPS (ICP_BLOCK / 32)
#define ROW_STRIDE 9 /* doubles per staged row: J0..J5, r, d2, one */

struct IcpState {
    double total[32];
    double local[32];
    float T[16];
    float U[16];
    float fitness, rmse;
    int done;        // 0 running, 1 converged (materialise correspondences next), 2 finished
    int iterations;  // updates applied
    int converged;
    int apply_u;
    unsigned ticket;
    unsigned tile_counter;
    unsigned cert_tiles;  // tiles skipped by their certificates in the launch just finished
    int static_sched;     // next search launch may use the static tile schedule (see icp_iteration_kernel)
    long long n_corr;
    unsigned pad_local;  // host-side staging only (count of locally written correspondence pairs)
    unsigned pad_;
};

struct IcpArgs {
    IndexView ix;
    float4 *src;            // working copy, Hilbert order, w = original index
    float4 *src_nrm;        // working normals (Symmetric) or null
    float4 *src_cov;        // working covariances: 3 float4 rows per point, [3][n_pad] (GICP) or null
    const float4 *src_col;  // colors in Hilbert order (Colored) or null
    const float *tgt_xyz, *tgt_nrm, *tgt_col, *tgt_grad, *tgt_cov;
    IcpState *st;
    double *partials;     // [reduce grid][32]
    double *tile_sums;    // [n_pad/32][32]
    int2 *prev;           // [n_pad] per Hilbert position: .x last iteration's match (-1 none), .y float bits of the
                          // certificate slack (lower bound on the distance to every OTHER target point); or null
    float cert_gain;      // margin = cert_gain * displacement (0 disables the certificates)
    float cert_cap;       // matched lanes: margins above cert_cap * (point spacing in the match's leaf) are not worth
                          // the wider search (-> plain search)
    float cert_cap_r;     // unmatched lanes: same, as a distance (a fraction of max_correspondence_distance)
    float r_up;           // max_correspondence_distance rounded up (certified 'still unmatched' test)
    unsigned *dbg;        // [2][64] certified lanes / skipped tiles per launch (CPHB_DEBUG_CERT) or null
    unsigned claim_max;   // largest range of tiles one claim may take (certified regime)
    int static_sched;     // allow the atomics-free static schedule once >= 90 % of the tiles are skipped
    int32_t *corr_index;  // [n_src] matched target index per ORIGINAL source index, or null
    unsigned long long n_total;
    unsigned n_src, n_pad;
    float r2;
    float rel_fitness, rel_rmse, det_thresh, sg, sp;
    int launch_idx, max_iter;
    int tgt_cov_col_major;
    int defer_finalize;  // multi-GPU over NCCL: stop after writing st->local
    int use_p2p;         // multi-GPU over the fused peer-memory exchange
    P2pView p2p;
    int step_mode;       // debug hook: one search + sums, no solve
    int tmax;            // transposed-scan threshold (tuning hook)
};

// sums layout (32 doubles): JTJ kinds: 0..20 JTJ upper | 21..26 JTr | 27 r^2 | 28 sum d2 | 29 count
//                           P2P      : 0..2 sum s | 3..5 sum t | 6..14 sum s t^T | 28 sum d2 | 29 count
__constant__ unsigned char c_pair_jtj[32][2] = {
    {0, 0}, {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {1, 1}, {1, 2}, {1, 3}, {1, 4}, {1, 5},
    {2, 2}, {2, 3}, {2, 4}, {2, 5}, {3, 3}, {3, 4}, {3, 5}, {4, 4}, {4, 5}, {5, 5},
    {0, 6}, {1, 6}, {2, 6}, {3, 6}, {4, 6}, {5, 6}, {6, 6}, {7, 8}, {8, 8}, {8, 8}, {8, 8}};
__constant__ unsigned char c_pair_p2p[32][2] = {
    {0, 8}, {1, 8}, {2, 8}, {3, 8}, {4, 8}, {5, 8}, {0, 3}, {0, 4}, {0, 5}, {1, 3}, {1, 4},
    {1, 5}, {2, 3}, {2, 4}, {2, 5}, {8, 8}, {8, 8}, {8, 8}, {8, 8}, {8, 8}, {8, 8},
    {8, 8}, {8, 8}, {8, 8}, {8, 8}, {8, 8}, {8, 8}, {8, 8}, {7, 8}, {8, 8}, {8, 8}, {8, 8}};
__constant__ unsigned c_live_jtj = 0x3fffffffu;                          // lanes 0..29
__constant__ unsigned c_live_p2p = 0x00007fffu | (1u << 28) | (1u << 29);  // 0..14, 28, 29

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
            if (KIND == CPHB_EST_COLORED_ICP && (!a.tgt_col || !a.src_col)) have = false;
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

// ===========================================================================
// estimator rows for one correspondence (source point vs, target index j)
// ===========================================================================
struct TargetAttrs {
    const float *tgt_xyz, *tgt_nrm, *tgt_col, *tgt_grad, *tgt_cov;
    int tgt_cov_col_major;
    float sg, sp;
    bool src_nrm, src_col, src_cov;  // source attribute present
};
template <int KIND, int NROWS>
__device__ __forceinline__ void build_rows(const TargetAttrs &a, const float s_x, const float s_y, const float s_z,
                                           const float *sn, const float4 cs_in, const float *Cs, unsigned j,
                                           float (&J)[NROWS][6], float (&r)[NROWS]) {
    struct { float x, y, z; } s = {s_x, s_y, s_z};
        const float vs[3] = {s.x, s.y, s.z};
        const float vt[3] = {a.tgt_xyz[3 * (size_t)j], a.tgt_xyz[3 * (size_t)j + 1], a.tgt_xyz[3 * (size_t)j + 2]};
        if (KIND == CPHB_EST_POINT_TO_POINT) {
            J[0][0] = vs[0]; J[0][1] = vs[1]; J[0][2] = vs[2];
            J[0][3] = vt[0]; J[0][4] = vt[1]; J[0][5] = vt[2];
        } else if (KIND == CPHB_EST_POINT_TO_PLANE) {  // transformation_estimation.cu:34-56
            if (a.tgt_nrm) {
                const float nt[3] = {a.tgt_nrm[3 * (size_t)j], a.tgt_nrm[3 * (size_t)j + 1], a.tgt_nrm[3 * (size_t)j + 2]};
                r[0] = dot3(vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2], nt[0], nt[1], nt[2]);
                cross3(vs, nt, J[0]);
                J[0][3] = nt[0]; J[0][4] = nt[1]; J[0][5] = nt[2];
            }
        } else if (KIND == CPHB_EST_SYMMETRIC) {  // transformation_estimation.cu:58-90
            if (a.tgt_nrm && a.src_nrm) {
                const float nn[3] = {sn[0] + a.tgt_nrm[3 * (size_t)j], sn[1] + a.tgt_nrm[3 * (size_t)j + 1],
                                     sn[2] + a.tgt_nrm[3 * (size_t)j + 2]};
                const float sm[3] = {vs[0] + vt[0], vs[1] + vt[1], vs[2] + vt[2]};
                r[0] = dot3(vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2], nn[0], nn[1], nn[2]);
                cross3(sm, nn, J[0]);
                J[0][3] = nn[0]; J[0][4] = nn[1]; J[0][5] = nn[2];
            }
        } else if (KIND == CPHB_EST_COLORED_ICP) {  // colored_icp.cu:150-216
            if (a.tgt_nrm && a.tgt_col && a.src_col && a.tgt_grad) {
                const size_t j3 = 3 * (size_t)j;
                const float nt[3] = {a.tgt_nrm[j3], a.tgt_nrm[j3 + 1], a.tgt_nrm[j3 + 2]};
                const float gt[3] = {a.tgt_grad[j3], a.tgt_grad[j3 + 1], a.tgt_grad[j3 + 2]};
                const float4 cs = cs_in;
                const float d[3] = {vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2]};
                const float dn = dot3(d[0], d[1], d[2], nt[0], nt[1], nt[2]);
                float cr[3];
                cross3(vs, nt, cr);
#pragma unroll
                for (int c = 0; c < 3; ++c) { J[0][c] = a.sg * cr[c]; J[0][3 + c] = a.sg * nt[c]; }
                r[0] = a.sg * dn;
                float pd[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) pd[c] = __fmaf_rn(-dn, nt[c], vs[c]) - vt[c];
                const float is = intensity(cs.x, cs.y, cs.z);
                const float it = intensity(a.tgt_col[j3], a.tgt_col[j3 + 1], a.tgt_col[j3 + 2]);
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
                for (int c = 0; c < 3; ++c) { J[1 % NROWS][c] = a.sp * cr[c]; J[1 % NROWS][3 + c] = a.sp * gm[c]; }
                r[1 % NROWS] = a.sp * (is - is0);
            }
        } else if (KIND == CPHB_EST_GENERALIZED_ICP) {  // generalized_icp.cu:63-105
            if (a.tgt_cov && a.src_cov) {
                float Mx[9], Mi[9], W[9];
                const float *ct = a.tgt_cov + 9 * (size_t)j;
#pragma unroll
                for (int p = 0; p < 3; ++p)
#pragma unroll
                    for (int q = 0; q < 3; ++q)
                        Mx[3 * p + q] = ct[a.tgt_cov_col_major ? (3 * q + p) : (3 * p + q)] + Cs[3 * p + q];
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

// ===========================================================================
// the fused per-iteration kernel
//
// Persistent warps: each warp repeatedly claims a tile of 32 consecutive
// (Hilbert-ordered) source points from an atomic counter, so per-tile cost
// variation never idles a block.  Warps are fully independent (no
// __syncthreads).  Per tile: apply the previous update in place -> warm-start
// the search from last iteration's match -> exact NN search -> estimator rows
// -> 32 column sums written to tile_sums[tile][32] (one coalesced 256-B store).
// icp_reduce_kernel then adds the tile sums in a fixed order (bitwise
// reproducible whatever the tile schedule was) and its last block runs the
// solve / convergence logic.
// ===========================================================================
// pull the rows of target point j that a tile reads first into L1 (no register is tied up, nothing waits)
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
template <int KIND>
__device__ __forceinline__ void prefetch_target(const IcpArgs &a, size_t j) {
    prefetch_l1(a.tgt_xyz + 3 * j);
    if ((KIND == CPHB_EST_POINT_TO_PLANE || KIND == CPHB_EST_SYMMETRIC || KIND == CPHB_EST_COLORED_ICP) && a.tgt_nrm)
        prefetch_l1(a.tgt_nrm + 3 * j);
    if (KIND == CPHB_EST_COLORED_ICP) {
        if (a.tgt_col) prefetch_l1(a.tgt_col + 3 * j);
        if (a.tgt_grad) prefetch_l1(a.tgt_grad + 3 * j);
    }
    if (KIND == CPHB_EST_GENERALIZED_ICP && a.tgt_cov) {
        prefetch_l1(a.tgt_cov + 9 * j);
        prefetch_l1(a.tgt_cov + 9 * j + 8);
    }
}

#define ICP_SEARCH_WARPS 4
#ifndef ICP_MIN_BLOCKS
#define ICP_MIN_BLOCKS 1  // resident blocks / SM the register allocation targets
#endif
// Build-time experiments (tools/build_variant.sh; the default build has both off):
//   ICP_LOWREG  keep the update transform in shared memory and prefetch the next tile with L1 hints instead of
//               registers, so that ICP_MIN_BLOCKS=9 (56 registers, 36 warps / SM) fits without spilling
//   CPHB_PDL    programmatic dependent launch: the kernels of the loop are launched with stream serialisation
//               relaxed, run their prologue while the previous kernel drains and wait (griddepcontrol.wait)
//               before they touch anything it wrote
//   ICP_DEEP_PIPE  under the static schedule, load the point + certificate TWO tiles ahead so that the L1 prefetch
//               of the next tile's target rows can be issued at the top of the current tile instead of its end
//               (r1_icp_certified_ncu: 58 % of the stall samples of a certified launch are long-scoreboard waits
//               on exactly that gather)
#ifndef ICP_LOWREG
#define ICP_LOWREG 0
#endif
//   ICP_FAST_START  read the whole per-launch state (done, apply_u, static_sched, U) with independent loads: three
//               dependent L2 round trips at the start of every warp become one (18 % of the stall samples of a
//               certified launch sit in this prologue)
//   ICP_DUAL    two instances of the kernel per iteration, one compiled for the searching launches (more resident
//               warps: ICP_MIN_BLOCKS_SEARCH) and one for the launches that run under the static schedule (more
//               registers, deeper pipeline); the device-side regime flag decides which of the two returns at once
#ifndef ICP_DEEP_PIPE
#define ICP_DEEP_PIPE 0
#endif
#ifndef ICP_DUAL
#define ICP_DUAL 0
#endif
#ifndef ICP_MIN_BLOCKS_SEARCH
#define ICP_MIN_BLOCKS_SEARCH 8
#endif
#ifndef ICP_FAST_START
#define ICP_FAST_START 0
#endif
#ifndef CPHB_PDL
#define CPHB_PDL 0
#endif
__device__ __forceinline__ void grid_dependency_wait() {
#if CPHB_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void grid_dependency_trigger() {
#if CPHB_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// MODE 0: one kernel for every launch (the default build).  ICP_DUAL: MODE 1 does the launches of the searching
// regime and returns at once under the static schedule, MODE 2 the reverse.
template <int KIND, int TOP, int MODE = 0>
__global__ void __launch_bounds__(ICP_SEARCH_WARPS * 32, (MODE == 1) ? ICP_MIN_BLOCKS_SEARCH : ICP_MIN_BLOCKS)
        icp_iteration_kernel(const __grid_constant__ IcpArgs a) {
    __shared__ __align__(16) float4 s_tile[ICP_SEARCH_WARPS][2 * CPHB_LEAF];
    __shared__ uint64_t s_bar[ICP_SEARCH_WARPS][2];
    __shared__ double s_rows[ICP_SEARCH_WARPS][32 * ROW_STRIDE];

    IcpState *st = a.st;
    const int warp = threadIdx.x >> 5, lane = lane_id();
#if CPHB_PDL
    // everything up to the wait may overlap the tail of the reduce kernel that precedes this launch: it must
    // not read the state that kernel writes (done, apply_u, U, static_sched, tile_counter).  The working
    // arrays were last written by the search launch before it, which had completed before the reduce kernel
    // released its dependents.
    grid_dependency_trigger();
    {
        const unsigned t0 = blockIdx.x * ICP_SEARCH_WARPS + warp;
        if (t0 < a.n_pad / 32) {
            prefetch_l1(&a.src[t0 * 32 + lane]);
            if (a.prev) prefetch_l1(&a.prev[t0 * 32 + lane]);
        }
    }
    grid_dependency_wait();
#endif
#if ICP_FAST_START
    const int done = *(volatile int *)&st->done;
    const int apply_u = *(volatile int *)&st->apply_u;
    const int static_word = *(volatile int *)&st->static_sched;
    float Ur[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) Ur[k] = *(volatile float *)&st->U[k];
    if (done == 2) return;
    const bool materialize = (done == 1);
    const bool apply = a.step_mode ? true : (!materialize && apply_u != 0);
#else
    const int done = *(volatile int *)&st->done;
    if (done == 2) return;
    const bool materialize = (done == 1);
    const bool apply = a.step_mode ? true : (!materialize && *(volatile int *)&st->apply_u != 0);
#endif

    WarpSearchC w;
    warp_search_setup(w, s_tile[warp], s_bar[warp]);
    w.tmax = a.tmax;
#if ICP_LOWREG
    __shared__ float s_U[12];
    if (threadIdx.x < 12) s_U[threadIdx.x] = apply ? st->U[threadIdx.x] : ((threadIdx.x % 5 == 0) ? 1.f : 0.f);
    __syncthreads();  // the only block-wide barrier: before any warp has started its tile loop
    const float *U = s_U;
#elif ICP_FAST_START
    float U[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) U[k] = apply ? Ur[k] : ((k % 5 == 0) ? 1.f : 0.f);
#else
    float U[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) U[k] = apply ? st->U[k] : ((k % 5 == 0) ? 1.f : 0.f);
#endif
    const unsigned n_tiles = a.n_pad / 32;
    const unsigned long long init = (a.r2 > 0.f) ? init_key(a.r2) : 0ull;
    const bool write_corr = a.corr_index && (materialize || a.step_mode || a.launch_idx == a.max_iter);
    const bool use_cert = a.prev && !a.step_mode && a.cert_gain > 0.f;
    double *rows = s_rows[warp];
    const unsigned char(*pair)[2] = (KIND == CPHB_EST_POINT_TO_POINT) ? c_pair_p2p : c_pair_jtj;
    const int ca = pair[lane][0], cb = pair[lane][1];
    const unsigned live = (KIND == CPHB_EST_POINT_TO_POINT) ? c_live_p2p : c_live_jtj;
    constexpr int NROWS = (KIND == CPHB_EST_COLORED_ICP) ? 2 : (KIND == CPHB_EST_GENERALIZED_ICP) ? 3 : 1;

    // Tile schedule.  Each warp starts on a static tile (its global warp id) and then claims RANGES of
    // consecutive tiles from an atomic counter: one tile at a time while tiles need a search (cost varies 10x
    // between tiles, fine-grained claims keep every warp busy), up to 8 at a time once its tiles are skipped by
    // their certificates -- 31 k same-address atomics would otherwise serialise in one L2 slice (~1.5 ns each)
    // and bound the launch at ~50 us.
    // Software pipeline: the claim after next and the loads of the NEXT tile's point and certificate are issued
    // at the top of the current tile, and the target rows the next tile reads first (its previous match) are
    // pulled into L1 at the end of the current tile, so a certified tile never waits on a chain of L2 round trips.
    const unsigned total_warps = gridDim.x * ICP_SEARCH_WARPS;
    // once (nearly) every tile is skipped the tiles cost the same, and a static round-robin schedule needs no
    // atomics at all; tile_sums are indexed by tile, so the schedule never affects the result
#if ICP_FAST_START
    const bool static_regime = a.static_sched && !a.step_mode && static_word != 0;
#else
    const bool static_regime = a.static_sched && !a.step_mode && *(volatile int *)&st->static_sched != 0;
#endif
    if (MODE == 1 && static_regime) return;   // the other instance of the pair runs this launch
    if (MODE == 2 && !static_regime) return;
    const bool static_sched = (MODE == 1) ? false : (MODE == 2) ? true : static_regime;
    unsigned tile = blockIdx.x * ICP_SEARCH_WARPS + warp;
    unsigned range_end = tile + 1;   // current range [tile, range_end)
    unsigned pend = 0, pend_sz = 1;  // claim in flight (result in lane 0) and its size
    unsigned csize = 1;              // size of the next claim
    unsigned n_skipped = 0;
    if (!static_sched && lane == 0) pend = atomicAdd(&st->tile_counter, 1u);
#if !ICP_LOWREG
    float4 s_pf = make_float4(0.f, 0.f, 0.f, 0.f);
    int2 pv_pf = make_int2(-1, 0);
    if (tile < n_tiles) {
        s_pf = a.src[tile * 32 + lane];
        if (a.prev) pv_pf = a.prev[tile * 32 + lane];
    }
#endif
#if ICP_DEEP_PIPE && !ICP_LOWREG
    // second pipeline stage (static schedule only): data of the tile after the current one
    float4 s_pf2 = make_float4(0.f, 0.f, 0.f, 0.f);
    int2 pv_pf2 = make_int2(-1, 0);
    if (static_sched && tile + total_warps < n_tiles) {
        s_pf2 = a.src[(tile + total_warps) * 32 + lane];
        if (a.prev) pv_pf2 = a.prev[(tile + total_warps) * 32 + lane];
    }
#endif
    unsigned tn = 0;
    for (; tile < n_tiles; tile = tn) {
#if ICP_LOWREG
        float4 s = a.src[tile * 32 + lane];  // L1 hit: prefetched while the previous tile was processed
        const int2 pv = a.prev ? a.prev[tile * 32 + lane] : make_int2(-1, 0);
#else
        float4 s = s_pf;
        const int2 pv = pv_pf;
#endif
        tn = tile + 1;
        if (static_sched) {
            tn = tile + total_warps;
        } else if (tn >= range_end) {  // last tile of the range: the next one comes from the claim in flight
            tn = __shfl_sync(CPHB_FULL, pend, 0) + total_warps;
            range_end = min(tn + pend_sz, n_tiles);
            pend_sz = csize;
            if (lane == 0) pend = atomicAdd(&st->tile_counter, csize);
        }
#if ICP_DEEP_PIPE && !ICP_LOWREG
        if (static_sched) {
            // the next tile's point + certificate arrived a tile ago: its target rows can start moving now and
            // have this whole tile to arrive; the loads issued here are for the tile after next
            s_pf = s_pf2;
            pv_pf = pv_pf2;
            if (tn < n_tiles && pv_pf.x >= 0) prefetch_target<KIND>(a, (size_t)pv_pf.x);
            const unsigned tnn = tn + total_warps;
            if (tn < n_tiles && tnn < n_tiles) {
                s_pf2 = a.src[tnn * 32 + lane];
                if (a.prev) pv_pf2 = a.prev[tnn * 32 + lane];
            }
        } else
#endif
        if (tn < n_tiles) {
#if ICP_LOWREG
            if (lane < 4) prefetch_l1(reinterpret_cast<const char *>(a.src + tn * 32) + 128 * lane);
            else if (lane < 6 && a.prev) prefetch_l1(reinterpret_cast<const char *>(a.prev + tn * 32) + 128 * (lane - 4));
#else
            s_pf = a.src[tn * 32 + lane];
            if (a.prev) pv_pf = a.prev[tn * 32 + lane];
#endif
        }
        const unsigned i = tile * 32 + lane;  // position in Hilbert order (< n_pad)
        const unsigned orig = __float_as_uint(s.w);
        const bool in_range = i < a.n_src;
        const float ox = s.x, oy = s.y, oz = s.z;  // position the certificate slack refers to

        // ---- PointCloud::Transform(update) on the working copy (pointcloud.cu:293-299) ----
        float sn[3] = {0.f, 0.f, 0.f};
        float Cs[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (apply) {
            const float x = s.x, y = s.y, z = s.z;
            s.x = __fadd_rn(dot3(U[0], U[1], U[2], x, y, z), U[3]);
            s.y = __fadd_rn(dot3(U[4], U[5], U[6], x, y, z), U[7]);
            s.z = __fadd_rn(dot3(U[8], U[9], U[10], x, y, z), U[11]);
            if (!a.step_mode) a.src[i] = s;
        }
        if (KIND == CPHB_EST_SYMMETRIC && a.src_nrm) {
            const float4 n4 = a.src_nrm[i];
            if (apply) {
                sn[0] = dot3(U[0], U[1], U[2], n4.x, n4.y, n4.z);
                sn[1] = dot3(U[4], U[5], U[6], n4.x, n4.y, n4.z);
                sn[2] = dot3(U[8], U[9], U[10], n4.x, n4.y, n4.z);
                if (!a.step_mode) a.src_nrm[i] = make_float4(sn[0], sn[1], sn[2], 0.f);
            } else {
                sn[0] = n4.x; sn[1] = n4.y; sn[2] = n4.z;
            }
        }
        if (KIND == CPHB_EST_GENERALIZED_ICP && a.src_cov) {
            float C[9];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float4 c4 = a.src_cov[(size_t)r * a.n_pad + i];
                C[3 * r] = c4.x; C[3 * r + 1] = c4.y; C[3 * r + 2] = c4.z;
            }
            if (apply) {  // RotateCovariances (geometry_utils.cu:257-265): (R*C)*R^T
                float tmp[9];
                const float R[9] = {U[0], U[1], U[2], U[4], U[5], U[6], U[8], U[9], U[10]};
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        tmp[3 * r + c] = dot3(R[3 * r], R[3 * r + 1], R[3 * r + 2], C[c], C[3 + c], C[6 + c]);
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        Cs[3 * r + c] = dot3(tmp[3 * r], tmp[3 * r + 1], tmp[3 * r + 2], R[3 * c], R[3 * c + 1], R[3 * c + 2]);
                if (!a.step_mode)
#pragma unroll
                    for (int r = 0; r < 3; ++r)
                        a.src_cov[(size_t)r * a.n_pad + i] = make_float4(Cs[3 * r], Cs[3 * r + 1], Cs[3 * r + 2], 0.f);
            } else {
#pragma unroll
                for (int r = 0; r < 9; ++r) Cs[r] = C[r];
            }
        }

        // ---- SearchRadius(.., max_nn = 1) (registration.cu:47) ---------------------------
        // Certificates: prev[i].y is a lower bound (rounded down) on the distance from this point's position at
        // the time it was last searched (minus the displacements since) to every target point other than its
        // match.  If, after this iteration's displacement, the old match is still strictly closer than that
        // bound, no other point can have a smaller (d2, index) key: the search would return the same match, so
        // the lane skips it.  All roundings go against the certificate (slack down, distances up, 1e-5 relative
        // guard against the <= 3e-7 relative error of the float d2 arithmetic the keys are made of).
        w.qx = s.x; w.qy = s.y; w.qz = s.z;
        w.best = init;
        w.m1 = 0x7f800000u;
        w.m2 = 0x7f800000u;
        w.margin = 0.f;
        bool cert = false;
        float slk = 0.f;
        if (a.prev && in_range) {
            const int pj = pv.x;
            float disp = 0.f;
            if (use_cert) {
                disp = __fmul_ru(sqrt_approx(dist2(s.x, s.y, s.z, ox, oy, oz)), 1.00001f);
                slk = __fsub_rd(__int_as_float(pv.y), disp);  // NaN (never searched) stays NaN: no certificate
                w.margin = __fmul_ru(a.cert_gain, disp);
            }
            if (pj >= 0) {
                // warm start: last iteration's match is a candidate like any other (same key
                // arithmetic), so the result is unchanged; it only tightens the bounds early
                const float d2p = dist2(s.x, s.y, s.z, a.tgt_xyz[3 * (size_t)pj], a.tgt_xyz[3 * (size_t)pj + 1],
                                        a.tgt_xyz[3 * (size_t)pj + 2]);
                const unsigned long long kp = ((unsigned long long)__float_as_uint(d2p) << 32) | (unsigned)pj;
                if (kp < init) {
                    w.best = kp;
                    if (use_cert) cert = __fmul_ru(sqrt_approx(d2p), 1.00001f) < slk;
                }
                if (!cert && w.margin > 0.f) {
                    // local scale: a leaf holds 32 neighbouring points, so sqrt(largest face area / 32) is about
                    // the point spacing around the match (surface or volume sampling alike, within 2x)
                    const Box bx = a.ix.boxes[0][a.ix.inv[pj] >> 5];
                    const float ex = bx.hi.x - bx.lo.x, ey = bx.hi.y - bx.lo.y, ez = bx.hi.z - bx.lo.z;
                    const float area = fmaxf(ex * ey, fmaxf(ex * ez, ey * ez));
                    if (w.margin > a.cert_cap * sqrtf(area * (1.f / 32.f))) w.margin = 0.f;
                }
            } else if (use_cert) {
                cert = slk > a.r_up;  // every target point is still outside the radius
                if (w.margin > a.cert_cap_r) w.margin = 0.f;
            }
            if (cert) w.margin = 0.f;
        }
        w.valid = in_range && !cert;
        w.track = __any_sync(CPHB_FULL, w.valid && w.margin > 0.f);
        w.refresh();
        warp_update_bound(w);
        w.warm = __all_sync(CPHB_FULL, !w.valid || w.best < init);  // every searching lane starts from a real candidate
        if (__any_sync(CPHB_FULL, w.valid)) {
            warp_query_box(w);
            warp_nn_search<TOP>(a.ix, w);
            csize = 1;
        } else {
            csize = min(csize * 2, a.claim_max);
            ++n_skipped;
        }
        if (a.dbg) {
            const unsigned nc = __popc(__ballot_sync(CPHB_FULL, cert));
            const bool searched = __any_sync(CPHB_FULL, w.valid);
            if (lane == 0) {
                atomicAdd(&a.dbg[min(a.launch_idx, 63)], nc);
                if (!searched) atomicAdd(&a.dbg[64 + min(a.launch_idx, 63)], 1u);
            }
        }
        const bool found = in_range && (w.best != init);
        const unsigned j = (unsigned)(w.best & 0xffffffffull);
        const float d2 = __uint_as_float((unsigned)(w.best >> 32));
        if (a.prev && !a.step_mode) {
            // searched lanes: everything not evaluated lies outside the final relaxed bound, everything evaluated
            // except the winner is at least sqrt(m2) away
            // m1 is the winner's own d2 (its leaf is always scanned); if it is not -- no match, or a tie -- m1
            // itself belongs to another point
            const unsigned other = (found && w.m1 == (unsigned)(w.best >> 32)) ? w.m2 : w.m1;
            const float l2 = __uint_as_float(min(other, w.rb));
            const float fresh = __fmul_rd(sqrt_approx(l2), 0.99999f);
            a.prev[i] = make_int2(found ? (int)j : -1, __float_as_int(cert ? slk : fresh));
        }
        if (write_corr && in_range) a.corr_index[orig] = found ? (int32_t)j : -1;
        if (materialize) continue;  // fitness / rmse / T of this pose are already in the state

        // ---- rows: J (6), r; staged as doubles, one row of 9 per lane ---------------------
        float J[NROWS][6], r[NROWS];
#pragma unroll
        for (int q = 0; q < NROWS; ++q) {
            r[q] = 0.f;
#pragma unroll
            for (int c = 0; c < 6; ++c) J[q][c] = 0.f;
        }
        if (found) {
            TargetAttrs ta = {a.tgt_xyz, a.tgt_nrm, a.tgt_col, a.tgt_grad, a.tgt_cov, a.tgt_cov_col_major, a.sg, a.sp,
                              a.src_nrm != nullptr, a.src_col != nullptr, a.src_cov != nullptr};
            const float4 cs4 = (KIND == CPHB_EST_COLORED_ICP && a.src_col) ? a.src_col[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            build_rows<KIND, NROWS>(ta, s.x, s.y, s.z, sn, cs4, Cs, j, J, r);
            drop_nonfinite_rows<NROWS>(J, r);
        }
        // stage + accumulate: lane L adds column pair (ca, cb) over the 32 staged rows, in row order
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < NROWS; ++q) {
            double *my = rows + lane * ROW_STRIDE;
#pragma unroll
            for (int c = 0; c < 6; ++c) my[c] = (double)J[q][c];
            my[6] = (double)r[q];
            my[7] = (q == 0 && found) ? (double)d2 : 0.0;
            my[8] = (q == 0 && found) ? 1.0 : 0.0;
            __syncwarp();
#pragma unroll 8
            for (int t = 0; t < 32; ++t) acc = fma(rows[t * ROW_STRIDE + ca], rows[t * ROW_STRIDE + cb], acc);
            __syncwarp();
        }
        if (!((live >> lane) & 1u)) acc = 0.0;
        a.tile_sums[(size_t)tile * 32 + lane] = acc;
#if ICP_LOWREG
        if (tn < n_tiles && a.prev) {
            const int pn = a.prev[tn * 32 + lane].x;  // L1 hit (hinted at the top of this tile)
            if (pn >= 0) prefetch_target<KIND>(a, (size_t)pn);
        }
#elif ICP_DEEP_PIPE
        if (!static_sched && tn < n_tiles && pv_pf.x >= 0) prefetch_target<KIND>(a, (size_t)pv_pf.x);
#else
        if (tn < n_tiles && pv_pf.x >= 0) prefetch_target<KIND>(a, (size_t)pv_pf.x);
#endif
    }
    if (a.static_sched && lane == 0 && n_skipped) atomicAdd(&st->cert_tiles, n_skipped);
}

// Fixed-order grid sum of the tile sums, then (last block) the host-side part of the loop.
// grid = R blocks; block b owns a contiguous chunk of tiles.
#define ICP_REDUCE_BLOCK 256
template <int KIND>
__global__ void __launch_bounds__(ICP_REDUCE_BLOCK) icp_reduce_kernel(const __grid_constant__ IcpArgs a) {
    __shared__ double s_acc[ICP_REDUCE_BLOCK / 32][32];
    __shared__ unsigned s_last;
    __shared__ SolveSmem s_solve;
    IcpState *st = a.st;
    grid_dependency_wait();     // the search launch has completed: its tile sums and state are visible
    grid_dependency_trigger();  // the next search launch may start its prologue while this kernel runs
    const int done = *(volatile int *)&st->done;
    if (done == 2) return;
    if (done == 1) {  // the search launch before this one only materialised correspondences
        if (blockIdx.x == 0 && threadIdx.x == 0) { st->tile_counter = 0; st->cert_tiles = 0; st->static_sched = 0; st->done = 2; }
        return;
    }
    const unsigned n_tiles = a.n_pad / 32;
    const unsigned chunk = (n_tiles + gridDim.x - 1) / gridDim.x;
    const unsigned t0 = blockIdx.x * chunk, t1 = min(n_tiles, t0 + chunk);
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    {
        // 8 loads in flight per thread, added in tile order (x + 0.0 is exact, so the padding loads of the
        // last batch do not change the sum): the naive loop serialises one L2 round trip per tile
        double t = 0.0;
        constexpr unsigned STRIDE = ICP_REDUCE_BLOCK / 32;
        for (unsigned k = t0 + g; k < t1; k += 8 * STRIDE) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned kk = k + u * STRIDE;
                v[u] = (kk < t1) ? __ldcg(&a.tile_sums[(size_t)kk * 32 + c]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) t += v[u];
        }
        s_acc[g][c] = t;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < ICP_REDUCE_BLOCK / 32; ++k) t += s_acc[k][threadIdx.x];
        a.partials[(size_t)blockIdx.x * 32 + threadIdx.x] = t;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(&st->ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {   // all 8 warps share the grid sum (fixed order: row groups of 8, then the 8 group sums)
        double t = 0.0;
        constexpr unsigned STRIDE = ICP_REDUCE_BLOCK / 32;
        for (unsigned b = g; b < gridDim.x; b += 8 * STRIDE) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned bb = b + u * STRIDE;
                v[u] = (bb < gridDim.x) ? __ldcg(&a.partials[(size_t)bb * 32 + c]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) t += v[u];
        }
        s_acc[g][c] = t;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < ICP_REDUCE_BLOCK / 32; ++k) t += s_acc[k][threadIdx.x];
        if (a.use_p2p) t = p2p_exchange_sum(a.p2p, t);  // the collective, fused: NVLink stores + flags
        if (a.defer_finalize) st->local[threadIdx.x] = t;
        else st->total[threadIdx.x] = t;
        __syncwarp();
        if (threadIdx.x == 0) {
            st->ticket = 0;
            st->tile_counter = 0;
            st->static_sched = ((unsigned long long)st->cert_tiles * 10ull >= (unsigned long long)n_tiles * 9ull) ? 1 : 0;
            st->cert_tiles = 0;
        }
        if (!a.defer_finalize) icp_finalize<KIND>(a, st, s_solve);
        __threadfence();
    }
}

// multi-GPU: runs after the all-reduce of st->local into st->total
template <int KIND>
__global__ void icp_finalize_kernel(const __grid_constant__ IcpArgs a) {
    __shared__ SolveSmem s_solve;
    if (threadIdx.x < 32 && a.st->done != 2) icp_finalize<KIND>(a, a.st, s_solve);
}

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

// ---------------------------------------------------------------------------
// source preparation + correspondence compaction kernels
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_source_kernel(const float *__restrict__ xyz, const float *__restrict__ nrm,
                                                            const float *__restrict__ col, const float *__restrict__ cov,
                                                            int cov_col_major, const uint32_t *__restrict__ perm_all,
                                                            unsigned lo, unsigned n, unsigned n_pad, float4 *o_xyz,
                                                            float4 *o_nrm, float4 *o_col, float4 *o_cov) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const uint32_t *perm = perm_all + lo;  // this rank's block of the Hilbert order
    if (i < n) {
        size_t j = perm[i];
        o_xyz[i] = make_float4(xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2], __uint_as_float((unsigned)j));
        if (o_nrm) o_nrm[i] = make_float4(nrm[3 * j], nrm[3 * j + 1], nrm[3 * j + 2], 0.f);
        if (o_col) o_col[i] = make_float4(col[3 * j], col[3 * j + 1], col[3 * j + 2], 0.f);
        if (o_cov) {
            const float *c = cov + 9 * j;
            for (int r = 0; r < 3; ++r)
                o_cov[(size_t)r * n_pad + i] = cov_col_major ? make_float4(c[r], c[3 + r], c[6 + r], 0.f)
                                                             : make_float4(c[3 * r], c[3 * r + 1], c[3 * r + 2], 0.f);
        }
    } else {
        // padding lanes: a copy of the last real point keeps loads in bounds; w = -1, never "valid"
        size_t j = n ? perm[n - 1] : 0;
        float4 p = n ? make_float4(xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
        p.w = __uint_as_float(0xffffffffu);
        o_xyz[i] = p;
        if (o_nrm) o_nrm[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o_col) o_col[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o_cov)
            for (int r = 0; r < 3; ++r) o_cov[(size_t)r * n_pad + i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---------------------------------------------------------------------------
// Re-tiling: once the clouds are roughly aligned, re-order the working copy by the Hilbert position of
// each point's current match, so that a warp's 32 queries fall into one or two target leaves instead
// of straddling a dozen.  Pure permutation of the working arrays (w / prev travel with the point): the
// result set is unchanged, only the order in which exact products are added to the float64 sums.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) retile_key_kernel(const int2 *__restrict__ prev, const uint32_t *__restrict__ inv,
                                                         unsigned n_src, unsigned n_pad, uint32_t n_tgt_pad, uint32_t *keys,
                                                         uint32_t *vals) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    uint32_t k = n_tgt_pad + 1u;  // padding stays last
    if (i < n_src) {
        const int pj = prev[i].x;
        k = (pj >= 0) ? inv[pj] : n_tgt_pad;  // unmatched points after the matched ones (target positions < n_tgt_pad)
    }
    keys[i] = k;
    vals[i] = i;
}
__global__ void __launch_bounds__(256) retile_gather_kernel(const uint32_t *__restrict__ order, unsigned n_pad,
                                                            const float4 *__restrict__ xyz, const int2 *__restrict__ prev,
                                                            const float4 *__restrict__ nrm, const float4 *__restrict__ col,
                                                            const float4 *__restrict__ cov, float4 *o_xyz, int2 *o_prev,
                                                            float4 *o_nrm, float4 *o_col, float4 *o_cov) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const unsigned s = order[i];
    o_xyz[i] = xyz[s];
    o_prev[i] = prev[s];
    if (nrm) o_nrm[i] = nrm[s];
    if (col) o_col[i] = col[s];
    if (cov)
#pragma unroll
        for (int r = 0; r < 3; ++r) o_cov[(size_t)r * n_pad + i] = cov[(size_t)r * n_pad + s];
}

// stable compaction of (i, corr_index[i]) with corr_index[i] >= 0 (registration.cu:54-69)
#define CMP_BLOCK 1024
__global__ void __launch_bounds__(CMP_BLOCK) compact_count_kernel(const int32_t *__restrict__ ci, unsigned n,
                                                                  unsigned *block_counts) {
    __shared__ unsigned s_w[32];
    unsigned i = blockIdx.x * CMP_BLOCK + threadIdx.x;
    bool v = i < n && ci[i] >= 0;
    unsigned m = __ballot_sync(CPHB_FULL, v);
    if (lane_id() == 0) s_w[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned c = s_w[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(CPHB_FULL, c, o);
        if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
    }
}
__global__ void __launch_bounds__(1024) compact_scan_kernel(unsigned *block_counts, unsigned nb, unsigned *total) {
    __shared__ unsigned s_w[32];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < nb; base += 1024) {
        unsigned i = base + threadIdx.x;
        unsigned v = i < nb ? block_counts[i] : 0u;
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(CPHB_FULL, x, o);
            if (lane_id() >= o) x += y;
        }
        if (lane_id() == 31) s_w[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned ws = s_w[threadIdx.x], z = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned y = __shfl_up_sync(CPHB_FULL, z, o);
                if (lane_id() >= o) z += y;
            }
            s_w[threadIdx.x] = z - ws;
        }
        __syncthreads();
        unsigned excl = x - v + s_w[threadIdx.x >> 5] + s_carry;
        if (i < nb) block_counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}
__global__ void __launch_bounds__(CMP_BLOCK) compact_write_kernel(const int32_t *__restrict__ ci, unsigned n,
                                                                  const unsigned *__restrict__ block_offsets,
                                                                  int32_t *out_pairs) {
    __shared__ unsigned s_w[32];
    unsigned i = blockIdx.x * CMP_BLOCK + threadIdx.x;
    int32_t j = i < n ? ci[i] : -1;
    bool v = j >= 0;
    unsigned m = __ballot_sync(CPHB_FULL, v);
    if (lane_id() == 0) s_w[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned ws = s_w[threadIdx.x], z = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(CPHB_FULL, z, o);
            if (lane_id() >= o) z += y;
        }
        s_w[threadIdx.x] = z - ws;
    }
    __syncthreads();
    if (v) {
        unsigned pos = block_offsets[blockIdx.x] + s_w[threadIdx.x >> 5] + __popc(m & ((1u << lane_id()) - 1u));
        out_pairs[2 * (size_t)pos] = (int32_t)i;
        out_pairs[2 * (size_t)pos + 1] = j;
    }
}

// PointCloud::Transform as a standalone op (pointcloud.cu:293-299)
__global__ void __launch_bounds__(256) transform_kernel(float *p, float *nrm, float *cov, int cov_col_major, size_t n,
                                                        const float *__restrict__ Tdev) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float U[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) U[k] = Tdev[k];
    if (p) {
        float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
        p[3 * i] = __fadd_rn(dot3(U[0], U[1], U[2], x, y, z), U[3]);
        p[3 * i + 1] = __fadd_rn(dot3(U[4], U[5], U[6], x, y, z), U[7]);
        p[3 * i + 2] = __fadd_rn(dot3(U[8], U[9], U[10], x, y, z), U[11]);
    }
    if (nrm) {
        float x = nrm[3 * i], y = nrm[3 * i + 1], z = nrm[3 * i + 2];
        nrm[3 * i] = dot3(U[0], U[1], U[2], x, y, z);
        nrm[3 * i + 1] = dot3(U[4], U[5], U[6], x, y, z);
        nrm[3 * i + 2] = dot3(U[8], U[9], U[10], x, y, z);
    }
    if (cov) {
        float C[9], tmp[9], O[9];
        float *c = cov + 9 * i;
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q) C[3 * r + q] = c[cov_col_major ? 3 * q + r : 3 * r + q];
        const float R[9] = {U[0], U[1], U[2], U[4], U[5], U[6], U[8], U[9], U[10]};
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q)
                tmp[3 * r + q] = dot3(R[3 * r], R[3 * r + 1], R[3 * r + 2], C[q], C[3 + q], C[6 + q]);
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q)
                O[3 * r + q] = dot3(tmp[3 * r], tmp[3 * r + 1], tmp[3 * r + 2], R[3 * q], R[3 * q + 1], R[3 * q + 2]);
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q) c[cov_col_major ? 3 * q + r : 3 * r + q] = O[3 * r + q];
    }
}

// ===========================================================================
// host driver
// ===========================================================================

struct cphb_icp {
    cphb_index *index;
    cphb_icp_params prm;
    cphb_cloud tgt;
    unsigned n_src, n_pad;
    unsigned n_full;   // size of the source as passed (== n_src unless the library shards it)
    void *arena;
    size_t arena_bytes;
    float4 *pristine_xyz, *pristine_nrm, *pristine_cov;  // Hilbert-ordered source as given
    float4 *work_xyz, *work_nrm, *work_cov;
    float4 *src_col;
    float4 *alt_xyz, *alt_nrm, *alt_cov, *alt_col, *cur_col;  // re-tiling ping-pong buffers
    int2 *alt_prev;
    uint32_t *rt_keys, *rt_keys2, *rt_vals, *rt_order;
    IcpState *st;
    IcpState *h_st;  // pinned
    bool owns_host;
    cudaEvent_t ev0, ev1;
    double *partials;
    int32_t *corr_index;
    unsigned *cmp_counts;
    unsigned *cmp_total;
    double *tile_sums;
    int2 *prev;
    unsigned *dbg = nullptr;  // CPHB_DEBUG_CERT statistics
    cudaEvent_t *dbg_ev = nullptr;  // CPHB_DEBUG_EVENTS: 3 events per launch (before, between, after)
    unsigned grid, reduce_grid;
    unsigned grid_search;  // ICP_DUAL: grid of the search-regime instance
    cudaStream_t stream;
};

// per-thread cached pinned staging buffer + events: cudaMallocHost / cudaEventCreate cost milliseconds,
// far more than a 1M-point registration's launch loop.
struct HostCache {
    IcpState *h_st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool in_use = false;
};
static thread_local HostCache t_cache;
// cphb_registration_icp_host: the source arrays are still being uploaded on another stream while the target index is
// built; cphb_icp_create waits for this event right before it first reads them
static thread_local cudaEvent_t t_source_ready = nullptr;

static bool is_identity4(const float *T) {  // Eigen isIdentity(1e-5), registration.cu:148
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float v = T[4 * i + j];
            if (i == j) { if (fabsf(v - 1.f) > 1e-5f) return false; }
            else if (fabsf(v) > 1e-5f) return false;
        }
    return true;
}

// resident blocks / SM of the iteration kernel instance a context will launch (register-limited)
template <int KIND, int MODE>
static int iteration_occupancy(bool top3) {
    int nb = 0;
    cudaError_t e = top3 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, icp_iteration_kernel<KIND, 3, MODE>, ICP_SEARCH_WARPS * 32, 0)
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, icp_iteration_kernel<KIND, 5, MODE>, ICP_SEARCH_WARPS * 32, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return nb;
}
template <int MODE>
static int iteration_occupancy_kind(int kind, bool top3) {
    static int cache[8][2] = {};  // 0 = not queried yet
    int &c = cache[kind & 7][top3 ? 0 : 1];
    if (c == 0) {
        switch (kind) {
            case CPHB_EST_POINT_TO_POINT: c = iteration_occupancy<CPHB_EST_POINT_TO_POINT, MODE>(top3); break;
            case CPHB_EST_POINT_TO_PLANE: c = iteration_occupancy<CPHB_EST_POINT_TO_PLANE, MODE>(top3); break;
            case CPHB_EST_SYMMETRIC: c = iteration_occupancy<CPHB_EST_SYMMETRIC, MODE>(top3); break;
            case CPHB_EST_COLORED_ICP: c = iteration_occupancy<CPHB_EST_COLORED_ICP, MODE>(top3); break;
            case CPHB_EST_GENERALIZED_ICP: c = iteration_occupancy<CPHB_EST_GENERALIZED_ICP, MODE>(top3); break;
        }
        if (c <= 0) c = -1;
    }
    return c;
}

#if CPHB_PDL
// launch with stream serialisation relaxed (programmatic dependent launch): the kernel may start while its
// predecessor drains and synchronises on it with griddepcontrol.wait
template <class K>
static void launch_pdl(K kernel, unsigned grid, unsigned block, cudaStream_t s, const IcpArgs &a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a);
    ++g_cphb_launches;
    if (e != cudaSuccess) cphb_set_error("PDL launch failed: %s", cudaGetErrorString(e));
}
#endif

template <int KIND>
static void launch_iteration(const cphb_icp *icp, const IcpArgs &a, cudaStream_t s) {
#if CPHB_PDL && !ICP_DUAL
    static const bool pdl = getenv("CPHB_NO_PDL") == nullptr;
    if (pdl && !a.defer_finalize && !a.step_mode && !icp->dbg_ev) {
        if (icp->index->v.top <= 3) launch_pdl(icp_iteration_kernel<KIND, 3>, icp->grid, ICP_SEARCH_WARPS * 32, s, a);
        else launch_pdl(icp_iteration_kernel<KIND, 5>, icp->grid, ICP_SEARCH_WARPS * 32, s, a);
        launch_pdl(icp_reduce_kernel<KIND>, icp->reduce_grid, ICP_REDUCE_BLOCK, s, a);
        return;
    }
#endif
#if ICP_DUAL
    if (icp->index->v.top <= 3) {
        CPHB_LAUNCH((icp_iteration_kernel<KIND, 3, 1>), icp->grid_search, ICP_SEARCH_WARPS * 32, 0, s, a);
        CPHB_LAUNCH((icp_iteration_kernel<KIND, 3, 2>), icp->grid, ICP_SEARCH_WARPS * 32, 0, s, a);
    } else {
        CPHB_LAUNCH((icp_iteration_kernel<KIND, 5, 1>), icp->grid_search, ICP_SEARCH_WARPS * 32, 0, s, a);
        CPHB_LAUNCH((icp_iteration_kernel<KIND, 5, 2>), icp->grid, ICP_SEARCH_WARPS * 32, 0, s, a);
    }
#else
    if (icp->index->v.top <= 3) CPHB_LAUNCH((icp_iteration_kernel<KIND, 3>), icp->grid, ICP_SEARCH_WARPS * 32, 0, s, a);
    else CPHB_LAUNCH((icp_iteration_kernel<KIND, 5>), icp->grid, ICP_SEARCH_WARPS * 32, 0, s, a);
#endif
    if (icp->dbg_ev) cudaEventRecord(icp->dbg_ev[3 * a.launch_idx + 1], s);
    CPHB_LAUNCH(icp_reduce_kernel<KIND>, icp->reduce_grid, ICP_REDUCE_BLOCK, 0, s, a);
}
static void launch_iteration_kind(const cphb_icp *icp, const IcpArgs &a, cudaStream_t s) {
    switch (icp->prm.estimation) {
        case CPHB_EST_POINT_TO_POINT: launch_iteration<CPHB_EST_POINT_TO_POINT>(icp, a, s); break;
        case CPHB_EST_POINT_TO_PLANE: launch_iteration<CPHB_EST_POINT_TO_PLANE>(icp, a, s); break;
        case CPHB_EST_SYMMETRIC: launch_iteration<CPHB_EST_SYMMETRIC>(icp, a, s); break;
        case CPHB_EST_COLORED_ICP: launch_iteration<CPHB_EST_COLORED_ICP>(icp, a, s); break;
        case CPHB_EST_GENERALIZED_ICP: launch_iteration<CPHB_EST_GENERALIZED_ICP>(icp, a, s); break;
    }
}
static void launch_finalize_kind(const cphb_icp *icp, const IcpArgs &a, cudaStream_t s) {
    switch (icp->prm.estimation) {
        case CPHB_EST_POINT_TO_POINT: CPHB_LAUNCH(icp_finalize_kernel<CPHB_EST_POINT_TO_POINT>, 1, 32, 0, s, a); break;
        case CPHB_EST_POINT_TO_PLANE: CPHB_LAUNCH(icp_finalize_kernel<CPHB_EST_POINT_TO_PLANE>, 1, 32, 0, s, a); break;
        case CPHB_EST_SYMMETRIC: CPHB_LAUNCH(icp_finalize_kernel<CPHB_EST_SYMMETRIC>, 1, 32, 0, s, a); break;
        case CPHB_EST_COLORED_ICP: CPHB_LAUNCH(icp_finalize_kernel<CPHB_EST_COLORED_ICP>, 1, 32, 0, s, a); break;
        case CPHB_EST_GENERALIZED_ICP: CPHB_LAUNCH(icp_finalize_kernel<CPHB_EST_GENERALIZED_ICP>, 1, 32, 0, s, a); break;
    }
}

extern "C" int cphb_icp_create(const cphb_cloud *source, const cphb_cloud *target, const cphb_icp_params *params,
                               void *stream, cphb_icp **out) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!source || !target || !params || !out) {
        cphb_set_error("cphb_icp_create: null argument");
        return CPHB_ERR_INVALID;
    }
    if (params->estimation < CPHB_EST_POINT_TO_POINT || params->estimation > CPHB_EST_GENERALIZED_ICP) {
        cphb_set_error("cphb_icp_create: estimation %d has no fused path (use the facade's generic loop)",
                       params->estimation);
        return CPHB_ERR_UNSUPPORTED;
    }
    if (source->n > 0x7fffff00ull || target->n > 0x7fffffffull) {
        cphb_set_error("cphb_icp_create: cloud exceeds int32 indices");
        return CPHB_ERR_INVALID;
    }
    cphb_icp *icp = new cphb_icp();
    memset(icp, 0, sizeof(*icp));
    icp->prm = *params;
    icp->tgt = *target;
    icp->stream = s;
    int rc = cphb_index_create(target->points, target->n, s, &icp->index);
    if (rc) { delete icp; return rc; }
    const unsigned n_full = (unsigned)source->n;
    unsigned lo = 0, n = n_full;
    if (params->shard_world > 1) {
        if (params->shard_rank < 0 || params->shard_rank >= params->shard_world) {
            cphb_set_error("cphb_icp_create: shard_rank %d outside [0,%d)", params->shard_rank, params->shard_world);
            cphb_index_destroy(icp->index);
            delete icp;
            return CPHB_ERR_INVALID;
        }
        lo = (unsigned)(((unsigned long long)n_full * params->shard_rank) / params->shard_world);
        unsigned hi = (unsigned)(((unsigned long long)n_full * (params->shard_rank + 1)) / params->shard_world);
        n = hi - lo;
    }
    icp->n_full = n_full;
    const unsigned n_pad = (unsigned)cphb_align(n ? n : 1, ICP_BLOCK);
    icp->n_src = n;
    icp->n_pad = n_pad;
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const unsigned n_tiles = n_pad / 32;
        unsigned want = (n_tiles + ICP_SEARCH_WARPS - 1) / ICP_SEARCH_WARPS;
        // every block must be resident: warps start on a static tile, and a block waiting for an SM slot
        // would hold its four tiles back until the dynamic queue has drained
        unsigned per_sm = 9u;
#if ICP_DUAL
        const int occ = iteration_occupancy_kind<2>(params->estimation, icp->index->v.top <= 3);
        {
            unsigned ps = 9u;
            const int occ_s = iteration_occupancy_kind<1>(params->estimation, icp->index->v.top <= 3);
            if (occ_s > 0 && (unsigned)occ_s < ps) ps = (unsigned)occ_s;
            const unsigned cap_s = (unsigned)sms * ps;
            icp->grid_search = want < cap_s ? want : cap_s;
        }
#else
        const int occ = iteration_occupancy_kind<0>(params->estimation, icp->index->v.top <= 3);
#endif
        if (occ > 0 && (unsigned)occ < per_sm) per_sm = (unsigned)occ;
        if (const char *e = getenv("CPHB_ICP_BLOCKS_PER_SM")) {  // tuning hook
            int v = atoi(e);
            if (v >= 1 && v <= 32) per_sm = (unsigned)v;
        }
        unsigned cap = (unsigned)sms * per_sm;
        icp->grid = want < cap ? want : cap;
        unsigned rg = (n_tiles + 63) / 64;
        icp->reduce_grid = rg < 1 ? 1 : (rg > (unsigned)sms ? (unsigned)sms : rg);
    }
    const bool want_nrm = params->estimation == CPHB_EST_SYMMETRIC && source->normals;
    const bool want_col = params->estimation == CPHB_EST_COLORED_ICP && source->colors;
    const bool want_cov = params->estimation == CPHB_EST_GENERALIZED_ICP && source->covariances;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = cphb_align(off + bytes, 256); return o; };
    size_t o_pxyz = take(sizeof(float4) * n_pad), o_wxyz = take(sizeof(float4) * n_pad);
    size_t o_pnrm = want_nrm ? take(sizeof(float4) * n_pad) : 0, o_wnrm = want_nrm ? take(sizeof(float4) * n_pad) : 0;
    size_t o_pcov = want_cov ? take(sizeof(float4) * 3 * n_pad) : 0, o_wcov = want_cov ? take(sizeof(float4) * 3 * n_pad) : 0;
    size_t o_col = want_col ? take(sizeof(float4) * n_pad) : 0;
    size_t o_st = take(sizeof(IcpState));
    size_t o_part = take(sizeof(double) * 32 * icp->reduce_grid);
    size_t o_ts = take(sizeof(double) * n_pad);
    size_t o_prev = take(sizeof(int2) * n_pad);
    size_t o_axyz = take(sizeof(float4) * n_pad), o_aprev = take(sizeof(int2) * n_pad);
    size_t o_anrm = want_nrm ? take(sizeof(float4) * n_pad) : 0, o_acov = want_cov ? take(sizeof(float4) * 3 * n_pad) : 0;
    size_t o_acol = want_col ? take(sizeof(float4) * n_pad) : 0, o_ccol = want_col ? take(sizeof(float4) * n_pad) : 0;
    size_t o_rt = take(sizeof(uint32_t) * 4 * n_pad);
    const unsigned nf_pad = (unsigned)cphb_align(n_full ? n_full : 1, ICP_BLOCK);
    size_t o_ci = take(sizeof(int32_t) * nf_pad);
    unsigned cmp_blocks = (nf_pad + CMP_BLOCK - 1) / CMP_BLOCK;
    size_t o_cc = take(sizeof(unsigned) * (cmp_blocks + 1));
    size_t o_ct = take(16);
    size_t o_perm = take(sizeof(uint32_t) * nf_pad);
    icp->arena_bytes = off;
    rc = cphb_alloc_async(&icp->arena, off, s);
    if (rc) { cphb_index_destroy(icp->index); delete icp; return rc; }
    char *b = (char *)icp->arena;
    icp->pristine_xyz = (float4 *)(b + o_pxyz);
    icp->work_xyz = (float4 *)(b + o_wxyz);
    icp->pristine_nrm = want_nrm ? (float4 *)(b + o_pnrm) : nullptr;
    icp->work_nrm = want_nrm ? (float4 *)(b + o_wnrm) : nullptr;
    icp->pristine_cov = want_cov ? (float4 *)(b + o_pcov) : nullptr;
    icp->work_cov = want_cov ? (float4 *)(b + o_wcov) : nullptr;
    icp->src_col = want_col ? (float4 *)(b + o_col) : nullptr;
    icp->st = (IcpState *)(b + o_st);
    icp->partials = (double *)(b + o_part);
    icp->tile_sums = (double *)(b + o_ts);
    icp->prev = (int2 *)(b + o_prev);
    icp->alt_xyz = (float4 *)(b + o_axyz);
    icp->alt_prev = (int2 *)(b + o_aprev);
    icp->alt_nrm = want_nrm ? (float4 *)(b + o_anrm) : nullptr;
    icp->alt_cov = want_cov ? (float4 *)(b + o_acov) : nullptr;
    icp->alt_col = want_col ? (float4 *)(b + o_acol) : nullptr;
    icp->cur_col = want_col ? (float4 *)(b + o_ccol) : nullptr;
    icp->rt_keys = (uint32_t *)(b + o_rt);
    icp->rt_keys2 = icp->rt_keys + n_pad;
    icp->rt_vals = icp->rt_keys + 2 * (size_t)n_pad;
    icp->rt_order = icp->rt_keys + 3 * (size_t)n_pad;
    icp->corr_index = (int32_t *)(b + o_ci);
    icp->cmp_counts = (unsigned *)(b + o_cc);
    icp->cmp_total = (unsigned *)(b + o_ct);
    uint32_t *perm = (uint32_t *)(b + o_perm);
    if (!t_cache.in_use) {
        if (!t_cache.h_st) {
            CPHB_CUDA(cudaMallocHost((void **)&t_cache.h_st, sizeof(IcpState)));
            CPHB_CUDA(cudaEventCreate(&t_cache.ev0));
            CPHB_CUDA(cudaEventCreate(&t_cache.ev1));
        }
        t_cache.in_use = true;
        icp->owns_host = false;
        icp->h_st = t_cache.h_st;
        icp->ev0 = t_cache.ev0;
        icp->ev1 = t_cache.ev1;
    } else {  // a second live context on this thread gets its own
        icp->owns_host = true;
        CPHB_CUDA(cudaMallocHost((void **)&icp->h_st, sizeof(IcpState)));
        CPHB_CUDA(cudaEventCreate(&icp->ev0));
        CPHB_CUDA(cudaEventCreate(&icp->ev1));
    }
    if (t_source_ready) {
        cudaStreamWaitEvent(s, t_source_ready, 0);
        t_source_ready = nullptr;
    }
    if (n_full) {
        rc = cphb_hilbert_order(source->points, n_full, perm, nullptr, 0, s);
        if (rc) { cphb_icp_destroy(icp); return rc; }
    }
    CPHB_LAUNCH(gather_source_kernel, n_pad / 256, 256, 0, s, source->points, source->normals, source->colors,
                source->covariances, source->cov_col_major, perm, lo, n, n_pad, icp->pristine_xyz, icp->pristine_nrm,
                icp->src_col, icp->pristine_cov);
    CPHB_CHECK_LAUNCH();
    *out = icp;
    return CPHB_OK;
}

extern "C" void cphb_icp_destroy(cphb_icp *icp) {
    if (!icp) return;
    if (icp->arena) cudaFreeAsync(icp->arena, icp->stream);
    if (icp->dbg) cudaFree(icp->dbg);
    if (icp->owns_host) {
        if (icp->h_st) cudaFreeHost(icp->h_st);
        if (icp->ev0) cudaEventDestroy(icp->ev0);
        if (icp->ev1) cudaEventDestroy(icp->ev1);
    } else if (icp->h_st == t_cache.h_st) {
        t_cache.in_use = false;
    }
    cphb_index_destroy(icp->index);
    delete icp;
}

static void fill_args(const cphb_icp *icp, IcpArgs &a) {
    memset(&a, 0, sizeof(a));
    a.ix = icp->index->v;
    a.src = icp->work_xyz;
    a.src_nrm = icp->work_nrm;
    a.src_cov = icp->work_cov;
    a.src_col = icp->src_col;
    a.tgt_xyz = icp->tgt.points;
    a.tgt_nrm = icp->tgt.normals;
    a.tgt_col = icp->tgt.colors;
    a.tgt_grad = icp->tgt.color_gradient;
    a.tgt_cov = icp->tgt.covariances;
    a.tgt_cov_col_major = icp->tgt.cov_col_major;
    a.st = icp->st;
    a.partials = icp->partials;
    a.tile_sums = icp->tile_sums;
    a.prev = icp->prev;
    a.n_src = icp->n_src;
    a.n_pad = icp->n_pad;
    a.n_total = icp->n_src;
    float r = icp->prm.max_correspondence_distance;
    a.r2 = (r > 0.f) ? r * r : 0.f;  // registration.cu:40-42: r <= 0 -> no correspondences
    a.rel_fitness = icp->prm.relative_fitness;
    a.rel_rmse = icp->prm.relative_rmse;
    a.det_thresh = icp->prm.det_thresh;
    a.max_iter = icp->prm.max_iteration > 0 ? icp->prm.max_iteration : 0;
    float lg = icp->prm.lambda_geometric;
    if (lg < 0.f || lg > 1.0f) lg = 0.968f;  // colored_icp.cu:48-52
    a.sg = (float)sqrt((double)lg);
    float lp = (float)(1.0 - (double)lg);
    a.sp = (float)sqrt((double)lp);
    a.cert_gain = 4.f;   // measured on config 2 (profiles/r1_cert_sweep.txt): 2..8 within 6 %
    a.cert_cap = 0.5f;
    if (const char *e = getenv("CPHB_CERT_GAIN")) a.cert_gain = (float)atof(e);  // tuning hooks; 0 = no certificates
    if (const char *e = getenv("CPHB_CERT_CAP")) a.cert_cap = (float)atof(e);
    a.cert_cap_r = 0.25f * (r > 0.f ? r : 0.f);
    a.r_up = (float)(sqrt((double)a.r2) * 1.00002);
    a.dbg = icp->dbg;
    a.claim_max = 8u;
    if (const char *e = getenv("CPHB_CLAIM_MAX")) { int v = atoi(e); if (v >= 1 && v <= 1024) a.claim_max = (unsigned)v; }
    a.static_sched = 1;  // 43 vs 49 us per certified launch on config 2 (profiles/r1_cert_events.txt)
    if (const char *e = getenv("CPHB_STATIC_SCHED")) a.static_sched = atoi(e) != 0;
    a.tmax = CPHB_TRANSPOSE_MAX;
    if (const char *e = getenv("CPHB_TRANSPOSE_MAX")) {  // tuning hook
        int v = atoi(e);
        if (v >= 0 && v <= 32) a.tmax = v;
    }
}

static int reset_working_copy(cphb_icp *icp, cudaStream_t s) {
    CPHB_CUDA(cudaMemsetAsync(icp->prev, 0xff, sizeof(int2) * icp->n_pad, s));  // no warm start at launch 0
    CPHB_CUDA(cudaMemcpyAsync(icp->work_xyz, icp->pristine_xyz, sizeof(float4) * icp->n_pad, cudaMemcpyDeviceToDevice, s));
    if (icp->work_nrm)
        CPHB_CUDA(cudaMemcpyAsync(icp->work_nrm, icp->pristine_nrm, sizeof(float4) * icp->n_pad, cudaMemcpyDeviceToDevice, s));
    if (icp->work_cov)
        CPHB_CUDA(cudaMemcpyAsync(icp->work_cov, icp->pristine_cov, sizeof(float4) * 3 * icp->n_pad, cudaMemcpyDeviceToDevice, s));
    return CPHB_OK;
}

static int compact_correspondences(cphb_icp *icp, int32_t *corr_out, cudaStream_t s) {
    unsigned n = icp->n_full;
    unsigned nb = (n + CMP_BLOCK - 1) / CMP_BLOCK;
    if (nb == 0) nb = 1;
    CPHB_LAUNCH(compact_count_kernel, nb, CMP_BLOCK, 0, s, icp->corr_index, n, icp->cmp_counts);
    CPHB_LAUNCH(compact_scan_kernel, 1, 1024, 0, s, icp->cmp_counts, nb, icp->cmp_total);
    CPHB_LAUNCH(compact_write_kernel, nb, CMP_BLOCK, 0, s, icp->corr_index, n, icp->cmp_counts, corr_out);
    CPHB_CHECK_LAUNCH();
    return CPHB_OK;
}

// permute the working arrays referenced by `a` into the alt buffers and swap
static int retile(cphb_icp *icp, IcpArgs &a, cudaStream_t s) {
    const unsigned n_pad = icp->n_pad, grid = n_pad / 256;
    const uint32_t n_tgt_pad = (uint32_t)icp->index->v.n_leaves * CPHB_LEAF;
    CPHB_LAUNCH(retile_key_kernel, grid, 256, 0, s, a.prev, icp->index->v.inv, icp->n_src, n_pad, n_tgt_pad, icp->rt_keys,
                icp->rt_vals);
    CPHB_CHECK_LAUNCH();
    // keys are target positions (< n_tgt_pad) or the two sentinels right above them: the stable radix sort only
    // has to look at the bits that can differ (21 for 1 M points: 3 onesweep passes instead of 4)
    int bits = 1;
    while (bits < 32 && ((uint64_t)1 << bits) < (uint64_t)n_tgt_pad + 2u) ++bits;
    int rc = cphb_sort_pairs_u32(icp->rt_keys, icp->rt_keys2, icp->rt_vals, icp->rt_order, n_pad, bits, s);
    if (rc) return rc;
    float4 *o_xyz = (a.src == icp->work_xyz) ? icp->alt_xyz : icp->work_xyz;
    int2 *o_prev = (a.prev == icp->prev) ? icp->alt_prev : icp->prev;
    float4 *o_nrm = a.src_nrm ? ((a.src_nrm == icp->work_nrm) ? icp->alt_nrm : icp->work_nrm) : nullptr;
    float4 *o_cov = a.src_cov ? ((a.src_cov == icp->work_cov) ? icp->alt_cov : icp->work_cov) : nullptr;
    float4 *o_col = a.src_col ? ((a.src_col == icp->alt_col) ? icp->cur_col : icp->alt_col) : nullptr;
    CPHB_LAUNCH(retile_gather_kernel, grid, 256, 0, s, icp->rt_order, n_pad, a.src, a.prev, a.src_nrm, a.src_col, a.src_cov,
                o_xyz, o_prev, o_nrm, o_col, o_cov);
    CPHB_CHECK_LAUNCH();
    a.src = o_xyz;
    a.prev = o_prev;
    a.src_nrm = o_nrm;
    a.src_cov = o_cov;
    a.src_col = o_col;
    return CPHB_OK;
}

extern "C" int cphb_icp_run(cphb_icp *icp, const float h_init[16], cphb_comm *comm, cphb_icp_result *h_result,
                            int32_t *corr_out, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!icp || !h_init || !h_result) {
        cphb_set_error("cphb_icp_run: null argument");
        return CPHB_ERR_INVALID;
    }
    int rc = reset_working_copy(icp, s);
    if (rc) return rc;
    IcpState *h = icp->h_st;
    memset(h, 0, sizeof(*h));
    memcpy(h->T, h_init, 64);
    memcpy(h->U, h_init, 64);
    h->apply_u = is_identity4(h_init) ? 0 : 1;
    CPHB_CUDA(cudaMemcpyAsync(icp->st, h, sizeof(IcpState), cudaMemcpyHostToDevice, s));
    static const bool dbg_cert = getenv("CPHB_DEBUG_CERT") != nullptr;
    if (dbg_cert) {
        if (!icp->dbg) CPHB_CUDA(cudaMalloc(&icp->dbg, sizeof(unsigned) * 128));
        CPHB_CUDA(cudaMemsetAsync(icp->dbg, 0, sizeof(unsigned) * 128, s));
    }
    IcpArgs a;
    fill_args(icp, a);
    a.corr_index = corr_out ? icp->corr_index : nullptr;
    if (corr_out && icp->n_full != icp->n_src)  // points owned by other ranks have no entry here
        CPHB_CUDA(cudaMemsetAsync(icp->corr_index, 0xff, sizeof(int32_t) * icp->n_full, s));
    unsigned long long n_total = icp->n_src;
    const bool lib_sharded = icp->prm.shard_world > 1;
    if (lib_sharded) n_total = icp->n_full;
    void *nccl_comm = nullptr;
    if (comm && comm->world > 1 && lib_sharded) {
        if (comm->kind == CPHB_COMM_NCCL) {
            nccl_comm = comm->nccl;
            a.defer_finalize = 1;
        } else {
            a.use_p2p = 1;
            a.p2p = comm->view;
        }
    } else if (comm && comm->world > 1) {
        // caller-sharded source: global source size by one tiny exchange before the loop
        double *tmp = icp->partials;  // scratch
        double hn = (double)icp->n_src;
        CPHB_CUDA(cudaMemcpyAsync(tmp, &hn, 8, cudaMemcpyHostToDevice, s));
        rc = cphb_comm_allreduce_f64(comm, tmp, 1, s);
        if (rc) return rc;
        CPHB_CUDA(cudaMemcpyAsync(&hn, tmp, 8, cudaMemcpyDeviceToHost, s));
        CPHB_CUDA(cudaStreamSynchronize(s));
        n_total = (unsigned long long)hn;
        if (comm->kind == CPHB_COMM_NCCL) {
            nccl_comm = comm->nccl;
            a.defer_finalize = 1;
        } else {
            a.use_p2p = 1;
            a.p2p = comm->view;
        }
    }
    a.n_total = n_total;
    // launches 0..max_iter: search (+ update).  If the convergence test stops the loop at launch
    // j < max_iter, launch j+1 re-runs that search only to materialise its correspondences;
    // launches after "done" exit at their first instruction.
    const unsigned long long launches0 = g_cphb_launches;
    // after these launches; later ones are skipped by their certificates, whatever the tile order
    unsigned long long retile_mask = (1ull << 1) | (1ull << 4);
    if (const char *e = getenv("CPHB_RETILE_MASK")) retile_mask = strtoull(e, nullptr, 0);  // tuning hook
    static const bool dbg_events = getenv("CPHB_DEBUG_EVENTS") != nullptr;
    std::vector<cudaEvent_t> evs;
    if (dbg_events) {
        evs.resize(3 * (size_t)(a.max_iter + 1));
        for (auto &e : evs) cudaEventCreate(&e);
        icp->dbg_ev = evs.data();
    }
    CPHB_CUDA(cudaEventRecord(icp->ev0, s));
    for (int it = 0; it <= a.max_iter; ++it) {
        a.launch_idx = it;
        if (dbg_events) cudaEventRecord(evs[3 * it], s);
        launch_iteration_kind(icp, a, s);
        if (dbg_events) cudaEventRecord(evs[3 * it + 2], s);
        if (nccl_comm) {
            rc = cphb_nccl_allreduce_f64(nccl_comm, icp->st->local, icp->st->total, 32, s);
            if (rc) return rc;
            launch_finalize_kind(icp, a, s);
        }
        if (!(icp->prm.flags & CPHB_ICP_NO_RETILE) && it < a.max_iter && it < 64 && ((retile_mask >> it) & 1ull) && icp->n_src >= 4096) {
            rc = retile(icp, a, s);
            if (rc) return rc;
        }
    }
    CPHB_CUDA(cudaEventRecord(icp->ev1, s));
    const int loop_launches = (int)(g_cphb_launches - launches0);
    CPHB_CHECK_LAUNCH();
    if (corr_out) {
        rc = compact_correspondences(icp, corr_out, s);
        if (rc) return rc;
    }
    unsigned h_local = 0;
    if (corr_out) CPHB_CUDA(cudaMemcpyAsync(&h->pad_local, icp->cmp_total, 4, cudaMemcpyDeviceToHost, s));
    CPHB_CUDA(cudaMemcpyAsync(h, icp->st, offsetof(IcpState, pad_local), cudaMemcpyDeviceToHost, s));
    CPHB_CUDA(cudaStreamSynchronize(s));
    h_local = h->pad_local;
    memcpy(h_result->transformation, h->T, 64);
    h_result->fitness = h->fitness;
    h_result->inlier_rmse = h->rmse;
    h_result->n_correspondences = h->n_corr;
    h_result->n_local_correspondences = corr_out ? (long long)h_local : 0;
    h_result->iterations = h->iterations;
    h_result->converged = h->converged;
    h_result->loop_ms = 0.f;
    cudaEventElapsedTime(&h_result->loop_ms, icp->ev0, icp->ev1);
    h_result->loop_launches = loop_launches;
    if (dbg_events) {
        fprintf(stderr, "[cphb] per launch us (search kernel / reduce+solve / until next launch):\n");
        for (int it = 0; it <= a.max_iter; ++it) {
            float t_it = 0.f, t_red = 0.f, t_gap = 0.f;
            cudaEventElapsedTime(&t_it, evs[3 * it], evs[3 * it + 1]);
            cudaEventElapsedTime(&t_red, evs[3 * it + 1], evs[3 * it + 2]);
            if (it < a.max_iter) cudaEventElapsedTime(&t_gap, evs[3 * it + 2], evs[3 * it + 3]);
            fprintf(stderr, " %d:%.1f/%.1f/%.1f", it, 1e3f * t_it, 1e3f * t_red, 1e3f * t_gap);
        }
        fprintf(stderr, "\n");
        for (auto &e : evs) cudaEventDestroy(e);
        icp->dbg_ev = nullptr;
    }
    if (dbg_cert && icp->dbg) {
        unsigned hd[128];
        CPHB_CUDA(cudaMemcpy(hd, icp->dbg, sizeof(hd), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[cphb] certificates (n_src %u, tiles %u): launch: certified lanes / skipped tiles\n", icp->n_src, icp->n_pad / 32);
        for (int it = 0; it <= a.max_iter && it < 64; ++it) fprintf(stderr, " %d:%u/%u", it, hd[it], hd[64 + it]);
        fprintf(stderr, "\n");
    }
    return CPHB_OK;
}

extern "C" int cphb_icp_step(cphb_icp *icp, const float h_T[16], double h_sums[32], int32_t *corr_index, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!icp || !h_T || !h_sums) {
        cphb_set_error("cphb_icp_step: null argument");
        return CPHB_ERR_INVALID;
    }
    int rc = reset_working_copy(icp, s);
    if (rc) return rc;
    IcpState *h = icp->h_st;
    memset(h, 0, sizeof(*h));
    memcpy(h->T, h_T, 64);
    memcpy(h->U, h_T, 64);
    h->apply_u = 1;
    CPHB_CUDA(cudaMemcpyAsync(icp->st, h, sizeof(IcpState), cudaMemcpyHostToDevice, s));
    IcpArgs a;
    fill_args(icp, a);
    a.step_mode = 1;
    a.corr_index = icp->corr_index;
    CPHB_CUDA(cudaMemsetAsync(icp->corr_index, 0xff, sizeof(int32_t) * icp->n_full, s));
    launch_iteration_kind(icp, a, s);
    CPHB_CHECK_LAUNCH();
    if (corr_index)
        CPHB_CUDA(cudaMemcpyAsync(corr_index, icp->corr_index, sizeof(int32_t) * icp->n_full, cudaMemcpyDeviceToDevice, s));
    CPHB_CUDA(cudaMemcpyAsync(h, icp->st, sizeof(IcpState), cudaMemcpyDeviceToHost, s));
    CPHB_CUDA(cudaStreamSynchronize(s));
    memcpy(h_sums, h->total, sizeof(double) * 32);
    return CPHB_OK;
}

extern "C" int cphb_registration_icp(const cphb_cloud *source, const cphb_cloud *target, const float h_init[16],
                                     const cphb_icp_params *params, cphb_comm *comm, cphb_icp_result *h_result,
                                     int32_t *corr_out, void *stream) {
    static const bool dbg = getenv("CPHB_DEBUG_TIMING") != nullptr;
    struct timespec t0, t1, t2, t3;
    if (dbg) clock_gettime(CLOCK_MONOTONIC, &t0);
    cphb_icp *icp = nullptr;
    int rc = cphb_icp_create(source, target, params, stream, &icp);
    if (rc) return rc;
    if (dbg) clock_gettime(CLOCK_MONOTONIC, &t1);
    rc = cphb_icp_run(icp, h_init, comm, h_result, corr_out, stream);
    if (dbg) clock_gettime(CLOCK_MONOTONIC, &t2);
    cphb_icp_destroy(icp);
    if (dbg) {
        clock_gettime(CLOCK_MONOTONIC, &t3);
        auto ms = [](const timespec &a, const timespec &b) { return (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6; };
        fprintf(stderr, "[cphb] registration_icp host ms: create %.3f run %.3f destroy %.3f (loop device %.3f)\n", ms(t0, t1),
                ms(t1, t2), ms(t2, t3), h_result->loop_ms);
    }
    return rc;
}

// One-shot registration from HOST buffers (pinned or pageable): the uploads run on a separate stream and overlap the
// target index build and the source ordering -- target points first (the index needs nothing else), then the
// source, then the target attributes that only the first iteration reads.  The clouds' pointers are host pointers;
// h_corr_out (optional, host, 2 * source.n int32) receives the (i, j) pairs.  Everything is complete on return.
struct HostUploadCache {
    cudaStream_t copy = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};
static thread_local HostUploadCache t_up;

extern "C" int cphb_registration_icp_host(const cphb_cloud *h_source, const cphb_cloud *h_target, const float h_init[16],
                                          const cphb_icp_params *params, cphb_comm *comm, cphb_icp_result *h_result,
                                          int32_t *h_corr_out, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!h_source || !h_target || !params || !h_result || !h_init) {
        cphb_set_error("cphb_registration_icp_host: null argument");
        return CPHB_ERR_INVALID;
    }
    if (!t_up.copy) {
        CPHB_CUDA(cudaStreamCreateWithFlags(&t_up.copy, cudaStreamNonBlocking));
        for (auto &e : t_up.ev) CPHB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const size_t ns = h_source->n, nt = h_target->n;
    // device arena: [tgt points | src points, normals, colors, covariances | tgt normals, colors, covariances, gradient | pairs]
    size_t off = 0;
    auto take = [&](const void *hp, size_t bytes) { size_t o = off; if (hp) off = cphb_align(off + bytes, 256); return hp ? o : (size_t)-1; };
    const size_t o_tp = take(h_target->points, 12 * nt);
    const size_t o_sp = take(h_source->points, 12 * ns), o_sn = take(h_source->normals, 12 * ns), o_sc = take(h_source->colors, 12 * ns),
                 o_sv = take(h_source->covariances, 36 * ns);
    const size_t o_tn = take(h_target->normals, 12 * nt), o_tc = take(h_target->colors, 12 * nt),
                 o_tv = take(h_target->covariances, 36 * nt), o_tg = take(h_target->color_gradient, 12 * nt);
    const size_t o_pairs = take(h_corr_out, 8 * ns);
    char *base = nullptr;
    int rc = cphb_alloc_async((void **)&base, off ? off : 256, s);
    if (rc) return rc;
    auto dp = [&](size_t o) { return o == (size_t)-1 ? (float *)nullptr : (float *)(base + o); };
    cudaStream_t c = t_up.copy;
    CPHB_CUDA(cudaEventRecord(t_up.ev[0], s));            // the arena exists from here on in stream order
    CPHB_CUDA(cudaStreamWaitEvent(c, t_up.ev[0], 0));
    auto up = [&](size_t o, const void *hp, size_t bytes) {
        if (hp && bytes) cudaMemcpyAsync(base + o, hp, bytes, cudaMemcpyHostToDevice, c);
    };
    up(o_tp, h_target->points, 12 * nt);
    CPHB_CUDA(cudaEventRecord(t_up.ev[1], c));            // target points: all the index build needs
    up(o_sp, h_source->points, 12 * ns);
    up(o_sn, h_source->normals, 12 * ns);
    up(o_sc, h_source->colors, 12 * ns);
    up(o_sv, h_source->covariances, 36 * ns);
    CPHB_CUDA(cudaEventRecord(t_up.ev[2], c));            // source: first read by the Hilbert ordering
    up(o_tn, h_target->normals, 12 * nt);
    up(o_tc, h_target->colors, 12 * nt);
    up(o_tv, h_target->covariances, 36 * nt);
    up(o_tg, h_target->color_gradient, 12 * nt);
    CPHB_CUDA(cudaEventRecord(t_up.ev[3], c));            // target attributes: first read by the first iteration
    cphb_cloud ds = *h_source, dt = *h_target;
    ds.points = dp(o_sp); ds.normals = dp(o_sn); ds.colors = dp(o_sc); ds.covariances = dp(o_sv); ds.color_gradient = nullptr;
    dt.points = dp(o_tp); dt.normals = dp(o_tn); dt.colors = dp(o_tc); dt.covariances = dp(o_tv); dt.color_gradient = dp(o_tg);
    CPHB_CUDA(cudaStreamWaitEvent(s, t_up.ev[1], 0));
    t_source_ready = t_up.ev[2];
    cphb_icp *icp = nullptr;
    rc = cphb_icp_create(&ds, &dt, params, stream, &icp);
    t_source_ready = nullptr;
    if (!rc) {
        cudaStreamWaitEvent(s, t_up.ev[2], 0);            // (already waited inside create; harmless if it returned early)
        cudaStreamWaitEvent(s, t_up.ev[3], 0);
        int32_t *d_pairs = (int32_t *)dp(o_pairs);
        rc = cphb_icp_run(icp, h_init, comm, h_result, d_pairs, stream);
        if (!rc && h_corr_out && h_result->n_local_correspondences > 0) {
            cudaError_t e = cudaMemcpyAsync(h_corr_out, d_pairs, sizeof(int32_t) * 2 * (size_t)h_result->n_local_correspondences,
                                            cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { cphb_set_error("cphb_registration_icp_host: %s", cudaGetErrorString(e)); rc = CPHB_ERR_CUDA; }
        }
        cphb_icp_destroy(icp);
    }
    cudaStreamSynchronize(c);                              // no upload may outlive the host buffers or the arena
    cphb_free_async(base, s);
    cudaStreamSynchronize(s);
    return rc;
}

extern "C" int cphb_evaluate_registration(const cphb_cloud *source, const cphb_cloud *target,
                                          float max_correspondence_distance, const float h_T[16],
                                          cphb_icp_result *h_result, int32_t *corr_out, void *stream) {
    cphb_icp_params p;
    memset(&p, 0, sizeof(p));
    p.estimation = CPHB_EST_POINT_TO_POINT;
    p.max_correspondence_distance = max_correspondence_distance;
    p.max_iteration = 0;
    cphb_cloud src = *source, tgt = *target;
    src.normals = src.colors = src.covariances = nullptr;
    tgt.normals = tgt.colors = tgt.covariances = tgt.color_gradient = nullptr;
    return cphb_registration_icp(&src, &tgt, h_T, &p, nullptr, h_result, corr_out, stream);
}

extern "C" int cphb_transform(float *points, float *normals, float *covariances, int cov_col_major, size_t n,
                              const float h_T[16], void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return CPHB_OK;
    float *Td = nullptr;
    int rc = cphb_alloc_async((void **)&Td, 64, s);
    if (rc) return rc;
    CPHB_CUDA(cudaMemcpyAsync(Td, h_T, 64, cudaMemcpyHostToDevice, s));
    CPHB_LAUNCH(transform_kernel, (unsigned)((n + 255) / 256), 256, 0, s, points, normals, covariances, cov_col_major, n, Td);
    CPHB_CHECK_LAUNCH();
    cphb_free_async(Td, s);
    return CPHB_OK;
}

// ---------------------------------------------------------------------------
// standalone estimator entry points
// ---------------------------------------------------------------------------
static int run_estimate(int estimation, const cphb_cloud *source, const cphb_cloud *target, const int32_t *corr,
                        size_t n_corr, const cphb_icp_params *params, double h_S[32], float h_T[16], cudaStream_t s) {
    if (!source || !target || (n_corr && !corr)) {
        cphb_set_error("estimate: null argument");
        return CPHB_ERR_INVALID;
    }
    if (estimation < CPHB_EST_POINT_TO_POINT || estimation > CPHB_EST_GENERALIZED_ICP) {
        cphb_set_error("estimate: estimation %d unsupported", estimation);
        return CPHB_ERR_UNSUPPORTED;
    }
    for (int i = 0; i < 32; ++i) h_S[i] = 0.0;
    if (h_T) for (int i = 0; i < 16; ++i) h_T[i] = (i % 5 == 0) ? 1.f : 0.f;
    if (n_corr == 0) return CPHB_OK;
    EstArgs a;
    memset(&a, 0, sizeof(a));
    float lg = params ? params->lambda_geometric : 0.968f;
    if (lg < 0.f || lg > 1.0f) lg = 0.968f;
    a.ta.tgt_xyz = target->points; a.ta.tgt_nrm = target->normals; a.ta.tgt_col = target->colors;
    a.ta.tgt_grad = target->color_gradient; a.ta.tgt_cov = target->covariances;
    a.ta.tgt_cov_col_major = target->cov_col_major;
    a.ta.sg = (float)sqrt((double)lg);
    a.ta.sp = (float)sqrt((double)(float)(1.0 - (double)lg));
    a.ta.src_nrm = source->normals != nullptr; a.ta.src_col = source->colors != nullptr;
    a.ta.src_cov = source->covariances != nullptr;
    a.src_xyz = source->points; a.src_nrm = source->normals; a.src_col = source->colors;
    a.src_cov = source->covariances; a.src_cov_col_major = source->cov_col_major;
    a.corr = corr;
    a.n_corr = (unsigned)n_corr;
    unsigned grid = (unsigned)((n_corr + ICP_BLOCK - 1) / ICP_BLOCK);
    char *buf = nullptr;
    size_t bytes = cphb_align(sizeof(double) * 32 * grid, 256) + 256 + 256 + 64;
    int rc = cphb_alloc_async((void **)&buf, bytes, s);
    if (rc) return rc;
    a.partials = (double *)buf;
    a.total = (double *)(buf + cphb_align(sizeof(double) * 32 * grid, 256));
    a.ticket = (unsigned *)((char *)a.total + 256);
    float *Td = (float *)((char *)a.total + 512);
    CPHB_CUDA(cudaMemsetAsync(a.ticket, 0, 4, s));
    bool have = true;
    switch (estimation) {
        case CPHB_EST_POINT_TO_POINT: CPHB_LAUNCH(estimate_kernel<CPHB_EST_POINT_TO_POINT>, grid, ICP_BLOCK, 0, s, a); break;
        case CPHB_EST_POINT_TO_PLANE: have = target->normals; CPHB_LAUNCH(estimate_kernel<CPHB_EST_POINT_TO_PLANE>, grid, ICP_BLOCK, 0, s, a); break;
        case CPHB_EST_SYMMETRIC: have = target->normals && source->normals; CPHB_LAUNCH(estimate_kernel<CPHB_EST_SYMMETRIC>, grid, ICP_BLOCK, 0, s, a); break;
        case CPHB_EST_COLORED_ICP: have = target->normals && target->colors && source->colors; CPHB_LAUNCH(estimate_kernel<CPHB_EST_COLORED_ICP>, grid, ICP_BLOCK, 0, s, a); break;
        case CPHB_EST_GENERALIZED_ICP: have = target->covariances && source->covariances; CPHB_LAUNCH(estimate_kernel<CPHB_EST_GENERALIZED_ICP>, grid, ICP_BLOCK, 0, s, a); break;
    }
    if (h_T) {
        float dt = params ? params->det_thresh : 1e-6f;
        unsigned long long nm = source->n;
        switch (estimation) {
            case CPHB_EST_POINT_TO_POINT: CPHB_LAUNCH(estimate_solve_kernel<CPHB_EST_POINT_TO_POINT>, 1, 32, 0, s, a.total, nm, dt, have, Td); break;
            case CPHB_EST_POINT_TO_PLANE: CPHB_LAUNCH(estimate_solve_kernel<CPHB_EST_POINT_TO_PLANE>, 1, 32, 0, s, a.total, nm, dt, have, Td); break;
            case CPHB_EST_SYMMETRIC: CPHB_LAUNCH(estimate_solve_kernel<CPHB_EST_SYMMETRIC>, 1, 32, 0, s, a.total, nm, dt, have, Td); break;
            case CPHB_EST_COLORED_ICP: CPHB_LAUNCH(estimate_solve_kernel<CPHB_EST_COLORED_ICP>, 1, 32, 0, s, a.total, nm, dt, have, Td); break;
            case CPHB_EST_GENERALIZED_ICP: CPHB_LAUNCH(estimate_solve_kernel<CPHB_EST_GENERALIZED_ICP>, 1, 32, 0, s, a.total, nm, dt, have, Td); break;
        }
        CPHB_CUDA(cudaMemcpyAsync(h_T, Td, 64, cudaMemcpyDeviceToHost, s));
    }
    CPHB_CHECK_LAUNCH();
    CPHB_CUDA(cudaMemcpyAsync(h_S, a.total, sizeof(double) * 32, cudaMemcpyDeviceToHost, s));
    CPHB_CUDA(cudaStreamSynchronize(s));
    cphb_free_async(buf, s);
    return CPHB_OK;
}

extern "C" int cphb_compute_transformation(int estimation, const cphb_cloud *source, const cphb_cloud *target,
                                           const int32_t *corr, size_t n_corr, const cphb_icp_params *params,
                                           float h_T[16], void *stream) {
    double S[32];
    return run_estimate(estimation, source, target, corr, n_corr, params, S, h_T, (cudaStream_t)stream);
}

extern "C" int cphb_compute_rmse(int estimation, const cphb_cloud *source, const cphb_cloud *target, const int32_t *corr,
                                 size_t n_corr, const cphb_icp_params *params, float *h_rmse, void *stream) {
    double S[32];
    *h_rmse = 0.f;
    int rc = run_estimate(estimation, source, target, corr, n_corr, params, S, nullptr, (cudaStream_t)stream);
    if (rc || n_corr == 0) return rc;
    const float C = (float)n_corr;
    switch (estimation) {
        case CPHB_EST_POINT_TO_POINT: *h_rmse = sqrtf((float)S[28] / C); break;                      // transformation_estimation.cu:92-116
        case CPHB_EST_POINT_TO_PLANE: *h_rmse = target->normals ? sqrtf((float)S[27] / C) : 0.f; break;  // :118-166
        case CPHB_EST_SYMMETRIC: *h_rmse = (target->normals && source->normals) ? sqrtf((float)S[28] / C) : 0.f; break;  // :224-287
        case CPHB_EST_GENERALIZED_ICP: *h_rmse = sqrtf((float)S[28] / C); break;                     // generalized_icp.cu:134-151
        case CPHB_EST_COLORED_ICP: *h_rmse = (float)S[27]; break;  // colored_icp.cu:303-327 returns the plain sum (quirk)
    }
    return CPHB_OK;
}

// registration::Kabsch(model, target[, corres]) (kabsch.h:30-49); corr == NULL pairs i<->i
__global__ void __launch_bounds__(256) iota_pairs_kernel(int32_t *p, unsigned n) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { p[2 * (size_t)i] = (int32_t)i; p[2 * (size_t)i + 1] = (int32_t)i; }
}
extern "C" int cphb_kabsch(const float *model, size_t n_model, const float *target, const int32_t *corr, size_t n_corr,
                           float h_T[16], void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    cphb_cloud src, tgt;
    memset(&src, 0, sizeof(src));
    memset(&tgt, 0, sizeof(tgt));
    src.points = model; src.n = n_model;
    tgt.points = target; tgt.n = n_model;
    int32_t *tmp = nullptr;
    if (!corr) {
        n_corr = n_model;
        int rc = cphb_alloc_async((void **)&tmp, sizeof(int32_t) * 2 * (n_model ? n_model : 1), s);
        if (rc) return rc;
        if (n_model) CPHB_LAUNCH(iota_pairs_kernel, (unsigned)((n_model + 255) / 256), 256, 0, s, tmp, (unsigned)n_model);
        corr = tmp;
    }
    int rc = cphb_compute_transformation(CPHB_EST_POINT_TO_POINT, &src, &tgt, corr, n_corr, nullptr, h_T, stream);
    cphb_free_async(tmp, s);
    return rc;
}
