// features.cu -- per-point neighbourhood features that sit immediately before
// the ICP loop: EstimateNormals (estimate_normals.cu:38-127), the GICP
// covariance initialisation (generalized_icp.cu:18-61) and the Colored-ICP
// intensity-gradient fit (colored_icp.cu:73-148).  EstimateNormals is ONE
// kernel (k-best search + cumulants + eigen-solve); the gradient fit is one
// search (search.cu) + one kernel; the reference uses a search + reduce_by_key
// over n*k cumulant tuples + a transform.
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include "cphb_eigen3.cuh"
#include "cphb_internal.cuh"
#include "cphb_searchk.cuh"

// compute_cumulant_functor (geometry_functor.h:35-55): one neighbour into the 9 cumulants (float64 sums of exact products)
__device__ __forceinline__ void cumulant_add(double (&cum)[9], float x, float y, float z) {
    cum[0] += x; cum[1] += y; cum[2] += z;
    cum[3] = fma((double)x, (double)x, cum[3]); cum[4] = fma((double)x, (double)y, cum[4]);
    cum[5] = fma((double)x, (double)z, cum[5]); cum[6] = fma((double)y, (double)y, cum[6]);
    cum[7] = fma((double)y, (double)z, cum[7]); cum[8] = fma((double)z, (double)z, cum[8]);
}
// ComputeNormal (estimate_normals.cu:38-54): covariance E[xx^T] - E[x]E[x]^T in float32, eigenvector of the smallest
// eigenvalue, (0,0,1) with fewer than 3 neighbours or a zero / NaN vector
__device__ __forceinline__ void normal_from_cumulants(const double (&cum)[9], int cnt, float &nx, float &ny, float &nz) {
    nx = 0.f; ny = 0.f; nz = 1.f;
    if (cnt < 3) return;
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
    const float nn = sqrtf(dot3(nx, ny, nz, nx, ny, nz));
    if (nn == 0.0f || nn != nn) { nx = 0.f; ny = 0.f; nz = 1.f; }
}

__global__ void __launch_bounds__(256) fill_normals_kernel(float *out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { out[3 * i] = 0.f; out[3 * i + 1] = 0.f; out[3 * i + 2] = 1.f; }
}

// two-pass form (search result [n][k] in HBM, then this kernel): kept as the A/B partner of the fused kernel below
// (CPHB_NORMALS_UNFUSED=1) -- the two must agree bit for bit (tests/test_gpu_geometry.py)
__global__ void __launch_bounds__(128) normals_kernel(const float *__restrict__ pts, size_t n,
                                                      const int32_t *__restrict__ nbr, int k, float *out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double cum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int j = 0; j < k; ++j) {
        int id = nbr[i * k + j];
        if (id < 0) continue;
        cumulant_add(cum, pts[3 * (size_t)id], pts[3 * (size_t)id + 1], pts[3 * (size_t)id + 2]);
        ++cnt;
    }
    float nx, ny, nz;
    normal_from_cumulants(cum, cnt, nx, ny, nz);
    out[3 * i] = nx; out[3 * i + 1] = ny; out[3 * i + 2] = nz;
}

// EstimateNormals in ONE kernel (estimate_normals.cu:82-127 + geometry_functor.h:35-55): the warp's
// k-best search (lists in shared memory) is followed, in the same kernel, by the 9 cumulants over each lane's list in
// list order and the 3x3 eigen-solve -- the [n][k] index / distance table (8 k bytes per point written and read back,
// 2.4 GB at 10 M x k = 30) never exists.
//   self_order: the queries are the indexed cloud itself, taken in INDEX order (lane = one point of a leaf, so a warp's
//               queries are exactly one leaf and need no ordering pass); original position = ix.pts[i].w
//   otherwise : query i is pts[first + perm[i]] (perm = Hilbert order of the block, or NULL), output row perm[i]
template <int TOP>
__global__ void normals_fused_kernel(IndexView ix, const float *__restrict__ pts, size_t first, size_t count,
                                     const uint32_t *__restrict__ perm, int self_order, float r2, int k, float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int nwarp = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    float4 *tiles = (float4 *)smem;
    uint64_t *bars = (uint64_t *)(tiles + (size_t)nwarp * 2 * CPHB_LEAF);
    unsigned long long *lists = (unsigned long long *)(bars + 2 * nwarp);
    WarpSearchK w;
    warp_search_setup(w, tiles + (size_t)warp * 2 * CPHB_LEAF, bars + 2 * warp);
    w.k = k;
    w.list = lists + (size_t)warp * k * 32 + lane_id();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t pos = 0;
    w.qx = w.qy = w.qz = 0.f;
    if (self_order) {
        w.valid = i < (size_t)ix.n_leaves * CPHB_LEAF;
        if (w.valid) {
            const float4 q = ix.pts[i];
            w.valid = __float_as_uint(q.w) != 0xffffffffu;  // padded tail of the index
            pos = __float_as_uint(q.w);
            w.qx = q.x; w.qy = q.y; w.qz = q.z;
        }
    } else {
        w.valid = i < count;
        if (w.valid) {
            pos = perm ? perm[i] : i;
            const float *q = pts + 3 * (first + pos);
            w.qx = q[0]; w.qy = q[1]; w.qz = q[2];
        }
    }
    const unsigned long long init = (r2 > 0.f) ? init_key(r2) : 0ull;
    for (int j = 0; j < k; ++j) w.list[j * 32] = init;
    w.worst = init;
    w.bound = (unsigned)(init >> 32);
    warp_query_box(w);
    if (__any_sync(CPHB_FULL, w.valid)) warp_nn_search<TOP>(ix, w);
    if (!w.valid) return;
    double cum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int j = 0; j < k; ++j) {
        const unsigned long long key = w.list[j * 32];
        if (!(key < init)) break;  // ascending: the unfilled slots are at the end
        const size_t id = (unsigned)(key & 0xffffffffull);
        cumulant_add(cum, pts[3 * id], pts[3 * id + 1], pts[3 * id + 2]);
        ++cnt;
    }
    float nx, ny, nz;
    normal_from_cumulants(cum, cnt, nx, ny, nz);
    out[3 * pos] = nx; out[3 * pos + 1] = ny; out[3 * pos + 2] = nz;
}

// launch of the fused kernel for the queries [first, first + count) of `points` against the index of the whole cloud
static int launch_normals_fused(const cphb_index *ix, const float *points, size_t n, size_t first, size_t count, int knn,
                                float radius, int k, float *out, cudaStream_t s) {
    const float r2 = (knn > 0) ? INFINITY : radius * radius;  // kdtree_flann.inl:120
    const bool self_order = (first == 0 && count == n);
    uint32_t *perm = nullptr;
    int rc = CPHB_OK;
    if (!self_order && count > 32) {
        rc = cphb_alloc_async((void **)&perm, sizeof(uint32_t) * count, s);
        if (rc) return rc;
        rc = cphb_hilbert_order_n(points + 3 * first, count, perm, ix->bounds, 1, count < n ? count : n, s);
        if (rc) { cphb_free_async(perm, s); return rc; }
    }
    const size_t per_warp = 2 * CPHB_LEAF * sizeof(float4) + 2 * sizeof(uint64_t) + (size_t)k * 32 * sizeof(unsigned long long);
    int warps = (int)((96 * 1024) / per_warp);
    if (warps > 8) warps = 8;
    if (warps < 1) warps = 1;
    const size_t smem = per_warp * warps;
    const unsigned block = warps * 32;
    const size_t lanes = self_order ? (size_t)ix->v.n_leaves * CPHB_LEAF : count;
    const unsigned grid = (unsigned)((lanes + block - 1) / block);
    if (ix->v.top <= 3) {
        CPHB_CUDA(cudaFuncSetAttribute(normals_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CPHB_LAUNCH(normals_fused_kernel<3>, grid, block, smem, s, ix->v, points, first, count, perm, self_order ? 1 : 0, r2, k, out);
    } else {
        CPHB_CUDA(cudaFuncSetAttribute(normals_fused_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CPHB_LAUNCH(normals_fused_kernel<5>, grid, block, smem, s, ix->v, points, first, count, perm, self_order ? 1 : 0, r2, k, out);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cphb_set_error("normals_fused_kernel: %s", cudaGetErrorString(e)); rc = CPHB_ERR_CUDA; }
    cphb_free_async(perm, s);
    return rc;
}

static int normals_two_pass(cphb_index *ix, const float *points, size_t first, size_t count, int knn, float radius, int k,
                            float *out, cudaStream_t s) {
    int32_t *idx = nullptr;
    float *d2 = nullptr;
    int rc = cphb_alloc_async((void **)&idx, sizeof(int32_t) * count * k, s);
    if (!rc) rc = cphb_alloc_async((void **)&d2, sizeof(float) * count * k, s);
    if (!rc) {
        const float *q = points + 3 * first;
        if (knn > 0) rc = cphb_search_knn(ix, q, count, k, idx, d2, nullptr, s);
        else rc = cphb_search_radius(ix, q, count, radius, k, idx, d2, nullptr, s);
    }
    if (!rc) {
        CPHB_LAUNCH(normals_kernel, (unsigned)((count + 127) / 128), 128, 0, s, points, count, idx, k, out);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { cphb_set_error("normals_kernel: %s", cudaGetErrorString(e)); rc = CPHB_ERR_CUDA; }
    }
    cphb_free_async(idx, s);
    cphb_free_async(d2, s);
    return rc;
}

// k limits as the search entry points have them: SearchKNN rejects knn > NUM_MAX_NN = 100 (kdtree_flann.cu:52-54),
// SearchRadius is bounded by what one warp's lists can hold in shared memory (search.cu)
static int normals_impl(const float *points, size_t n, int knn, float radius, int max_nn, size_t first, size_t count,
                        float *out_normals, cudaStream_t s) {
    const int k = (knn > 0) ? knn : max_nn;
    if (k <= 0) {  // estimate_normals.cu:102-106
        CPHB_LAUNCH(fill_normals_kernel, (unsigned)((count + 255) / 256), 256, 0, s, out_normals, count);
        CPHB_CHECK_LAUNCH();
        return CPHB_OK;
    }
    if ((knn > 0 && k > 100) || k > 256) {
        cphb_set_error("search: k=%d outside [0,%d]", k, knn > 0 ? 100 : 256);
        return CPHB_ERR_INVALID;
    }
    cphb_index *ix = nullptr;
    int rc = cphb_index_create(points, n, s, &ix);  // the reference also builds a fresh tree (:86-87)
    if (rc) return rc;
    const bool unfused = getenv("CPHB_NORMALS_UNFUSED") != nullptr;  // A/B hook (read per call): identical results
    if (unfused) rc = normals_two_pass(ix, points, first, count, knn, radius, k, out_normals, s);
    else rc = launch_normals_fused(ix, points, n, first, count, knn, radius, k, out_normals, s);
    cudaStreamSynchronize(s);
    cphb_index_destroy(ix);
    return rc;
}

extern "C" int cphb_estimate_normals(const float *points, size_t n, int knn, float radius, int max_nn,
                                     float *out_normals, void *stream) {
    if (n == 0) return CPHB_OK;
    if (!points || !out_normals) {
        cphb_set_error("cphb_estimate_normals: null argument");
        return CPHB_ERR_INVALID;
    }
    return normals_impl(points, n, knn, radius, max_nn, 0, n, out_normals, (cudaStream_t)stream);
}

// Multi-GPU building block: normals of the points [first, first + count) only, neighbours taken
// from the whole cloud.  Every rank indexes the full cloud and estimates its own block; the blocks are then
// all-gathered (cupoch_b200.distributed.estimate_normals).  Same kernel as cphb_estimate_normals (its queries in the
// block's own Hilbert order instead of index order): identical per-point arithmetic, identical results.
extern "C" int cphb_estimate_normals_range(const float *points, size_t n, int knn, float radius, int max_nn, size_t first,
                                           size_t count, float *out_normals, void *stream) {
    if (count == 0) return CPHB_OK;
    if (!points || !out_normals || first > n || count > n - first) {
        cphb_set_error("cphb_estimate_normals_range: invalid argument");
        return CPHB_ERR_INVALID;
    }
    return normals_impl(points, n, knn, radius, max_nn, first, count, out_normals, (cudaStream_t)stream);
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
