// icp_aux.cuh -- working-copy kernels of the loop: source gather, re-tiling, correspondence compaction, PointCloud::Transform
// Part of the icp.cu translation unit (included there); split out for readability only.
#pragma once

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

// Private copies of the target attributes in INDEX order (icp_types.cuh): row p belongs to the point ix.pts[p].
// Padding rows (w == -1) are zero.  intensity(): colored_icp.cu:176-181, evaluated once here instead of once per row.
__global__ void __launch_bounds__(256) gather_target_kernel(const float4 *__restrict__ pts, size_t n_pad,
                                                            const float *__restrict__ nrm, const float *__restrict__ col,
                                                            const float *__restrict__ grad, const float *__restrict__ cov,
                                                            int cov_col_major, float4 *o_nrm, float4 *o_grad, float4 *o_cov) {
    size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= n_pad) return;
    const unsigned j = __float_as_uint(pts[p].w);
    const bool real = j != 0xffffffffu;
    const size_t j3 = 3 * (size_t)j;
    if (o_nrm) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (real) {
            v.x = nrm[j3]; v.y = nrm[j3 + 1]; v.z = nrm[j3 + 2];
            if (col) v.w = intensity(col[j3], col[j3 + 1], col[j3 + 2]);
        }
        o_nrm[p] = v;
    }
    if (o_grad) o_grad[p] = real ? make_float4(grad[j3], grad[j3 + 1], grad[j3 + 2], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (o_cov) {
        const float *c = cov + 9 * (size_t)j;
#pragma unroll
        for (int r = 0; r < 3; ++r)
            o_cov[3 * p + r] = !real ? make_float4(0.f, 0.f, 0.f, 0.f)
                                     : cov_col_major ? make_float4(c[r], c[3 + r], c[6 + r], 0.f)
                                                     : make_float4(c[3 * r], c[3 * r + 1], c[3 * r + 2], 0.f);
    }
}

// ---------------------------------------------------------------------------
// Re-tiling: once the clouds are roughly aligned, re-order the working copy by the Hilbert position of
// each point's current match, so that a warp's 32 queries fall into one or two target leaves instead
// of straddling a dozen.  Pure permutation of the working arrays (w / prev travel with the point): the
// result set is unchanged, only the order in which exact products are added to the float64 sums.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) retile_key_kernel(const int2 *__restrict__ prev, unsigned n_src, unsigned n_pad,
                                                         uint32_t n_tgt_pad, uint32_t *keys, uint32_t *vals) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    uint32_t k = (n_tgt_pad >> 5) + 1u;  // padding stays last
    if (i < n_src) {
        const int pp = prev[i].x;  // index position of the match
        // by LEAF of the match (the order inside a leaf does not matter for locality, and the stable sort keeps it
        // deterministic): 5 key bits less = one radix pass less at 1 M points.  Unmatched points after the matched ones
        k = (pp >= 0) ? ((uint32_t)pp >> 5) : (n_tgt_pad >> 5);
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

