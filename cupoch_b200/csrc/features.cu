// features.cu -- per-point neighbourhood features that sit immediately before
// the ICP loop: EstimateNormals (estimate_normals.cu:38-127), the GICP
// covariance initialisation (generalized_icp.cu:18-61) and the Colored-ICP
// intensity-gradient fit (colored_icp.cu:73-148).  Each is one search
// (search.cu) + one kernel; the reference uses a search + reduce_by_key over
// n*k cumulant tuples + a transform.
#include <float.h>
#include <math.h>

#include "cphb_eigen3.cuh"
#include "cphb_internal.cuh"

// compute_cumulant_functor (geometry_functor.h:35-55) + ComputeNormal (estimate_normals.cu:38-54)
__global__ void __launch_bounds__(128) normals_kernel(const float *__restrict__ pts, size_t n,
                                                      const int32_t *__restrict__ nbr, int k, float *out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double cum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int j = 0; j < k; ++j) {
        int id = nbr[i * k + j];
        if (id < 0) continue;
        const float x = pts[3 * (size_t)id], y = pts[3 * (size_t)id + 1], z = pts[3 * (size_t)id + 2];
        cum[0] += x; cum[1] += y; cum[2] += z;
        cum[3] = fma((double)x, (double)x, cum[3]); cum[4] = fma((double)x, (double)y, cum[4]);
        cum[5] = fma((double)x, (double)z, cum[5]); cum[6] = fma((double)y, (double)y, cum[6]);
        cum[7] = fma((double)y, (double)z, cum[7]); cum[8] = fma((double)z, (double)z, cum[8]);
        ++cnt;
    }
    float nx = 0.f, ny = 0.f, nz = 1.f;
    if (cnt >= 3) {
        float c[9];
#pragma unroll
        for (int a = 0; a < 9; ++a) c[a] = (float)cum[a] / (float)cnt;
        float cov[9];
        cov[0] = __fmaf_rn(-c[0], c[0], c[3]);
        cov[4] = __fmaf_rn(-c[1], c[1], c[6]);
        cov[8] = __fmaf_rn(-c[2], c[2], c[8]);
        cov[1] = cov[3] = __fmaf_rn(-c[0], c[1], c[4]);
        cov[2] = cov[6] = __fmaf_rn(-c[0], c[2], c[5]);
        cov[5] = cov[7] = __fmaf_rn(-c[1], c[2], c[7]);
        float e[3], V[9];
        fast_eigen3x3(cov, e, V);
        int mi = 0;
        if (e[1] < e[mi]) mi = 1;
        if (e[2] < e[mi]) mi = 2;
        nx = V[mi]; ny = V[3 + mi]; nz = V[6 + mi];
        float nn = sqrtf(dot3(nx, ny, nz, nx, ny, nz));
        if (nn == 0.0f || nn != nn) { nx = 0.f; ny = 0.f; nz = 1.f; }
    }
    out[3 * i] = nx; out[3 * i + 1] = ny; out[3 * i + 2] = nz;
}

__global__ void __launch_bounds__(256) fill_normals_kernel(float *out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { out[3 * i] = 0.f; out[3 * i + 1] = 0.f; out[3 * i + 2] = 1.f; }
}

extern "C" int cphb_estimate_normals(const float *points, size_t n, int knn, float radius, int max_nn,
                                     float *out_normals, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return CPHB_OK;
    if (!points || !out_normals) {
        cphb_set_error("cphb_estimate_normals: null argument");
        return CPHB_ERR_INVALID;
    }
    const int k = (knn > 0) ? knn : max_nn;
    if (k <= 0) {  // estimate_normals.cu:102-106
        CPHB_LAUNCH(fill_normals_kernel, (unsigned)((n + 255) / 256), 256, 0, s, out_normals, n);
        CPHB_CHECK_LAUNCH();
        return CPHB_OK;
    }
    cphb_index *ix = nullptr;
    int rc = cphb_index_create(points, n, stream, &ix);  // the reference also builds a fresh tree (:86-87)
    if (rc) return rc;
    int32_t *idx = nullptr;
    float *d2 = nullptr;
    rc = cphb_alloc_async((void **)&idx, sizeof(int32_t) * n * k, s);
    if (!rc) rc = cphb_alloc_async((void **)&d2, sizeof(float) * n * k, s);
    if (!rc) {
        if (knn > 0) rc = cphb_search_knn(ix, points, n, k, idx, d2, nullptr, stream);
        else rc = cphb_search_radius(ix, points, n, radius, k, idx, d2, nullptr, stream);
    }
    if (!rc) {
        CPHB_LAUNCH(normals_kernel, (unsigned)((n + 127) / 128), 128, 0, s, points, n, idx, k, out_normals);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { cphb_set_error("normals_kernel: %s", cudaGetErrorString(e)); rc = CPHB_ERR_CUDA; }
    }
    cphb_free_async(idx, s);
    cphb_free_async(d2, s);
    cudaStreamSynchronize(s);
    cphb_index_destroy(ix);
    return rc;
}

// Multi-GPU building block: normals of the points [first, first + count) only, neighbours taken
// from the whole cloud.  Every rank indexes the full cloud and estimates its own block; the blocks are then
// all-gathered (cupoch_b200.distributed.estimate_normals).  Same kernels as cphb_estimate_normals: rows of the
// neighbour table index the full point array, so a block of query rows needs nothing new.
extern "C" int cphb_estimate_normals_range(const float *points, size_t n, int knn, float radius, int max_nn, size_t first,
                                           size_t count, float *out_normals, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (count == 0) return CPHB_OK;
    if (!points || !out_normals || first > n || count > n - first) {
        cphb_set_error("cphb_estimate_normals_range: invalid argument");
        return CPHB_ERR_INVALID;
    }
    const int k = (knn > 0) ? knn : max_nn;
    if (k <= 0) {
        CPHB_LAUNCH(fill_normals_kernel, (unsigned)((count + 255) / 256), 256, 0, s, out_normals, count);
        CPHB_CHECK_LAUNCH();
        return CPHB_OK;
    }
    cphb_index *ix = nullptr;
    int rc = cphb_index_create(points, n, stream, &ix);
    if (rc) return rc;
    int32_t *idx = nullptr;
    float *d2 = nullptr;
    rc = cphb_alloc_async((void **)&idx, sizeof(int32_t) * count * k, s);
    if (!rc) rc = cphb_alloc_async((void **)&d2, sizeof(float) * count * k, s);
    if (!rc) {
        const float *q = points + 3 * first;
        if (knn > 0) rc = cphb_search_knn(ix, q, count, k, idx, d2, nullptr, stream);
        else rc = cphb_search_radius(ix, q, count, radius, k, idx, d2, nullptr, stream);
    }
    if (!rc) {
        CPHB_LAUNCH(normals_kernel, (unsigned)((count + 127) / 128), 128, 0, s, points, count, idx, k, out_normals);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { cphb_set_error("normals_kernel: %s", cudaGetErrorString(e)); rc = CPHB_ERR_CUDA; }
    }
    cphb_free_async(idx, s);
    cphb_free_async(d2, s);
    cudaStreamSynchronize(s);
    cphb_index_destroy(ix);
    return rc;
}

// GetRotationFromE1ToX + Rx*diag(eps,1,1)*Rx^T (generalized_icp.cu:18-30,53-60)
__global__ void __launch_bounds__(256) cov_from_normals_kernel(const float *__restrict__ nrm, size_t n, float eps,
                                                               float *out, int col_major) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x[3] = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const float c = x[0];
    if (!(c < -0.99f)) {
        const float v[3] = {0.f, -x[2], x[1]};
        const float sv[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
        float ss[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                ss[3 * a + b] = dot3(sv[3 * a], sv[3 * a + 1], sv[3 * a + 2], sv[b], sv[3 + b], sv[6 + b]);
        const float factor = 1 / (1 + c);
#pragma unroll
        for (int a = 0; a < 9; ++a) R[a] = __fmaf_rn(ss[a], factor, R[a] + sv[a]);
    }
    const float cd[3] = {eps, 1.f, 1.f};
    float tmp[9], C[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) tmp[3 * a + b] = R[3 * a + b] * cd[b];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
            C[3 * a + b] = dot3(tmp[3 * a], tmp[3 * a + 1], tmp[3 * a + 2], R[3 * b], R[3 * b + 1], R[3 * b + 2]);
    float *o = out + 9 * i;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) o[col_major ? 3 * b + a : 3 * a + b] = C[3 * a + b];
}

extern "C" int cphb_covariances_from_normals(const float *normals, size_t n, float epsilon, float *out_cov,
                                             int cov_col_major, void *stream) {
    if (n == 0) return CPHB_OK;
    if (!normals || !out_cov) {
        cphb_set_error("cphb_covariances_from_normals: null argument");
        return CPHB_ERR_INVALID;
    }
    CPHB_LAUNCH(cov_from_normals_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, normals, n, epsilon, out_cov,
                cov_col_major);
    CPHB_CHECK_LAUNCH();
    return CPHB_OK;
}

// compute_color_gradient_functor (colored_icp.cu:73-118): slot 0 of the neighbour row is skipped
__global__ void __launch_bounds__(128) color_gradient_kernel(const float *__restrict__ pts, const float *__restrict__ nrm,
                                                             const float *__restrict__ col, size_t n,
                                                             const int32_t *__restrict__ nbr, int k, float *out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float vt[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    const float nt[3] = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
    const float it = intensity(col[3 * i], col[3 * i + 1], col[3 * i + 2]);
    float AtA[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Atb[3] = {0, 0, 0};
    int nn = 0;
    for (int j = 1; j < k; ++j) {
        const int a = nbr[i * k + j];
        if (a < 0) continue;
        const float va[3] = {pts[3 * (size_t)a], pts[3 * (size_t)a + 1], pts[3 * (size_t)a + 2]};
        const float s = dot3(va[0] - vt[0], va[1] - vt[1], va[2] - vt[2], nt[0], nt[1], nt[2]);
        float vtmp[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) vtmp[c] = __fmaf_rn(-s, nt[c], va[c]) - vt[c];
        const float di = intensity(col[3 * (size_t)a], col[3 * (size_t)a + 1], col[3 * (size_t)a + 2]) - it;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) AtA[3 * r + c] = __fmaf_rn(vtmp[r], vtmp[c], AtA[3 * r + c]);
            Atb[r] = __fmaf_rn(di, vtmp[r], Atb[r]);
        }
        ++nn;
    }
    float o[3] = {0.f, 0.f, 0.f};
    if (nn >= 4) {
        const float w = (float)((nn - 1) * (nn - 1));
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float wn = w * nt[r];
#pragma unroll
            for (int c = 0; c < 3; ++c) AtA[3 * r + c] = __fmaf_rn(wn, nt[c], AtA[3 * r + c]);
        }
        AtA[0] += 1.0e-6f; AtA[4] += 1.0e-6f; AtA[8] += 1.0e-6f;
        float inv[9];
        inverse3x3(AtA, inv);
#pragma unroll
        for (int r = 0; r < 3; ++r) o[r] = dot3(inv[3 * r], inv[3 * r + 1], inv[3 * r + 2], Atb[0], Atb[1], Atb[2]);
    }
    out[3 * i] = o[0]; out[3 * i + 1] = o[1]; out[3 * i + 2] = o[2];
}

extern "C" int cphb_color_gradient(const float *points, const float *normals, const float *colors, size_t n,
                                   float radius, int max_nn, float *out_gradient, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return CPHB_OK;
    if (!points || !normals || !colors || !out_gradient || max_nn < 1 || max_nn > 100) {
        cphb_set_error("cphb_color_gradient: invalid argument");
        return CPHB_ERR_INVALID;
    }
    cphb_index *ix = nullptr;
    int rc = cphb_index_create(points, n, stream, &ix);
    if (rc) return rc;
    int32_t *idx = nullptr;
    float *d2 = nullptr;
    rc = cphb_alloc_async((void **)&idx, sizeof(int32_t) * n * max_nn, s);
    if (!rc) rc = cphb_alloc_async((void **)&d2, sizeof(float) * n * max_nn, s);
    if (!rc) rc = cphb_search_radius(ix, points, n, radius, max_nn, idx, d2, nullptr, stream);
    if (!rc) {
        CPHB_LAUNCH(color_gradient_kernel, (unsigned)((n + 127) / 128), 128, 0, s, points, normals, colors, n, idx, max_nn,
                    out_gradient);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { cphb_set_error("color_gradient_kernel: %s", cudaGetErrorString(e)); rc = CPHB_ERR_CUDA; }
    }
    cphb_free_async(idx, s);
    cphb_free_async(d2, s);
    cudaStreamSynchronize(s);
    cphb_index_destroy(ix);
    return rc;
}
