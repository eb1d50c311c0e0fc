// icp_kernels.cuh -- the two kernels of one ICP iteration: icp_iteration_kernel (transform -> certificate / exact 1-NN search ->
// estimator rows -> per-tile column sums) and icp_reduce_kernel (fixed-order sum, multi-GPU exchange, solve), plus the build-time
// experiment switches.  Part of the icp.cu translation unit (included there); split out for readability only.
#pragma once
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
//   ICP_LANE_ACC  under the static schedule every lane keeps the products of ITS OWN rows in 30 float64 registers over all
//               the tiles of its warp and the warp reduces them once, at its end: no shared-memory staging, no
//               per-tile 32-step DFMA chain, W instead of n_tiles rows for the reduce kernel.  The static schedule
//               makes the order (and so every bit) reproducible; it differs from the per-tile order in the last
//               bits of the float64 sums only
#ifndef ICP_DEEP_PIPE
#define ICP_DEEP_PIPE 0
#endif
#ifndef ICP_LANE_ACC
#define ICP_LANE_ACC 0
#endif
// closed forms of c_pair_jtj / c_pair_p2p (icp_types.cuh) for compile-time unrolling
__host__ __device__ constexpr int pair_jtj_a(int p) {
    return p < 6 ? 0 : p < 11 ? 1 : p < 15 ? 2 : p < 18 ? 3 : p < 20 ? 4 : p < 21 ? 5 : p < 27 ? p - 21 : p == 27 ? 6 : p == 28 ? 7 : 8;
}
__host__ __device__ constexpr int pair_jtj_b(int p) {
    return p < 6 ? p : p < 11 ? p - 5 : p < 15 ? p - 9 : p < 18 ? p - 12 : p < 20 ? p - 14 : p < 21 ? 5 : p < 28 ? 6 : 8;
}
__host__ __device__ constexpr int pair_p2p_a(int p) { return p < 6 ? p : p < 15 ? (p - 6) / 3 : p == 28 ? 7 : 8; }
__host__ __device__ constexpr int pair_p2p_b(int p) { return p < 6 ? 8 : p < 15 ? 3 + (p - 6) % 3 : 8; }
__host__ __device__ constexpr bool pair_live(bool p2p, int p) { return p2p ? (p < 15 || p == 28 || p == 29) : p < 30; }
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
#if ICP_LANE_ACC
    double lacc[32];
#pragma unroll
    for (int p = 0; p < 32; ++p) lacc[p] = 0.0;
#endif
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
#if ICP_LANE_ACC
        if (static_sched) {
            // this lane's own rows, every product into its own float64 accumulator (exact products, one rounding per
            // addition); the warp-wide reduction happens once, after the last tile
            constexpr bool P2P = (KIND == CPHB_EST_POINT_TO_POINT);
#pragma unroll
            for (int q = 0; q < NROWS; ++q) {
                const double v[9] = {(double)J[q][0], (double)J[q][1], (double)J[q][2], (double)J[q][3], (double)J[q][4],
                                     (double)J[q][5], (double)r[q], (q == 0 && found) ? (double)d2 : 0.0,
                                     (q == 0 && found) ? 1.0 : 0.0};
#pragma unroll
                for (int p = 0; p < 32; ++p)
                    if (pair_live(P2P, p))
                        lacc[p] = fma(v[P2P ? pair_p2p_a(p) : pair_jtj_a(p)], v[P2P ? pair_p2p_b(p) : pair_jtj_b(p)], lacc[p]);
            }
#if ICP_DEEP_PIPE
            continue;  // (the deep pipeline issued this tile's prefetches at its top)
#else
            if (tn < n_tiles && pv_pf.x >= 0) prefetch_target<KIND>(a, (size_t)pv_pf.x);
            continue;
#endif
        }
#endif
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
#if ICP_LANE_ACC
    {
        const unsigned gw = blockIdx.x * ICP_SEARCH_WARPS + warp;
        if (static_sched && !materialize && gw < n_tiles) {
            constexpr bool P2P = (KIND == CPHB_EST_POINT_TO_POINT);
            double out = 0.0;
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                if (!pair_live(P2P, p)) continue;
                double t = lacc[p];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(CPHB_FULL, t, o);  // same bits on every lane
                if (lane == p) out = t;
            }
            a.tile_sums[(size_t)gw * 32 + lane] = out;  // one row per warp instead of one per tile
        }
        // rows the reduce kernel has to add: one per warp that owned a tile, or one per tile
        if (gw == 0 && lane == 0 && !materialize) st->sum_rows = static_sched ? min(total_warps, n_tiles) : n_tiles;
    }
#endif
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
#if ICP_LANE_ACC
    const unsigned n_rows = *(volatile unsigned *)&st->sum_rows;
#else
    const unsigned n_rows = n_tiles;
#endif
    const unsigned chunk = (n_rows + gridDim.x - 1) / gridDim.x;
    const unsigned t0 = blockIdx.x * chunk, t1 = min(n_rows, t0 + chunk);
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

