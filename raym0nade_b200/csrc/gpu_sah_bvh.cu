// gpu_sah_bvh.cu — the secondary-ray tree built on the device by a top-down SWEEP SAH.
//
// gpu_bvh.cu builds that tree bottom-up (Morton sort + PLOC) in ~10 ms per million triangles, but its trees cost ~14 % more
// node visits per ray than the host's binned-SAH tree, which is why a host thread used to refine it in the background for
// ~0.4 s ("tree_builder" 2).  This builder closes the gap on the device: the exact sweep SAH, every level of the tree as a
// handful of maps and scans over all n triangles (sah_sweep.h has the scheme and the per-element code, shared with a host
// mirror that tests it without a GPU).  Its trees need ~2 % FEWER visits than the host's, so nothing has to be refined,
// cached or shared between ranks any more: every upload builds the final tree.
//
//   k_tri_boxes      per triangle: box, the three centre keys, the triangle's node of the binary tree
//   radix sort x 3   the three lists (cub::DeviceRadixSort - library plumbing)
//   per level        cub::DeviceScan::InclusiveScan of box unions (6 n items produced on the fly by an input iterator, reduced
//                    to their surface by an output iterator: the boxes themselves never touch memory) -> k_candidates
//                    (cost of every split, minimum per node: warp / CTA pre-reduced 64-bit atomicMin) -> k_decide (one thread
//                    per node) -> k_mark -> cub::DeviceScan::ExclusiveSum of 3 n flags -> k_scatter (stable partition of
//                    the three lists); one 8-byte read-back (nodes and slots handed out) tells the host when to stop
//   k_refit          boxes of the inner nodes, deepest level first
//   collapse         the 4-wide form (gpu_bvh.cu: rm_gpu_collapse_wide)
//
// There is no reference counterpart (the reference's tree, src/bvh.cpp:18-54, is what primary rays traverse);
// tests/test_gpu_trace.py::test_secondary_ray_tree_finds_the_reference_hits runs on this tree.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/iterator/transform_output_iterator.h>

#include "rm_context.cuh"
#include "gpu_bvh.h"
#include "sah_sweep.h"

namespace {

using namespace rm_sah;

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) k_tri_boxes(const float *__restrict__ pos, int n, float3 fallback, Tree T, float4 *__restrict__ tbox, float *__restrict__ keys,
                                                      int *__restrict__ ids) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float *p = pos + size_t(t) * 9;
    float lo[3], hi[3];
    bool finite = true;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        lo[a] = fminf(fminf(p[a], p[3 + a]), p[6 + a]);
        hi[a] = fmaxf(fmaxf(p[a], p[3 + a]), p[6 + a]);
        finite = finite && isfinite(p[a]) && isfinite(p[3 + a]) && isfinite(p[6 + a]);
    }
    if (!finite) { lo[0] = hi[0] = fallback.x; lo[1] = hi[1] = fallback.y; lo[2] = hi[2] = fallback.z; }      // can never be hit; must not derail the order
    const float4 l = make_float4(lo[0], lo[1], lo[2], 0.0f), h = make_float4(hi[0], hi[1], hi[2], 0.0f);
    T.lo[t] = l;
    T.hi[t] = h;
    tbox[2 * size_t(t)] = l;
    tbox[2 * size_t(t) + 1] = h;
    T.left[t] = ~t;
    T.right[t] = -1;
    T.count[t] = 1;
#pragma unroll
    for (int a = 0; a < 3; a++) keys[size_t(a) * n + t] = centre_key(l, h, a);
    ids[t] = t;
}

__global__ void k_start(int n, int *aL, int *aR, int *aB, unsigned long long *best, int *counters) {
    aL[0] = 0;
    aR[0] = n;
    aB[0] = n;              // the root is the first inner node
    best[0] = kNoSplit;
    counters[0] = n + 1;    // next_node
    counters[1] = 0;        // next_slot
    counters[2] = 0;        // RM_SAH_SCAN=check: areas that differ between the two scans
}

struct ItemOp {
    Level V;
    __device__ SweepItem operator()(int idx) const { return sweep_item(V, idx); }
};
struct AreaOp {
    __device__ float operator()(const SweepItem &b) const { return sweep_area(b); }
};
struct LeftOp {
    Level V;
    const uint8_t *side;
    __device__ int operator()(int c) const { return sweep_goes_left(V, side, c); }
};

__device__ __forceinline__ unsigned long long warp_min(unsigned long long k) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, d);
        k = o < k ? o : k;
    }
    return k;
}

// ---- the segmented scan of box unions, written out: cub::DeviceScan moves 32-byte items through its decoupled look-back at
// ~10 G items/s (0.6 ms per level and million triangles - three quarters of the whole build); as three plain kernels - fold
// every tile of kTile items to one aggregate, scan the few thousand aggregates in one CTA, scan every tile again from its
// carry - the same scan runs at the speed of its gathers.  (SweepUnion is exact min / max: both give the same bits.)
constexpr int kPerThread = 8, kTile = kBlock * kPerThread;

__device__ __forceinline__ SweepItem scan_identity() {
    SweepItem e;
    e.lx = e.ly = e.lz = CUDART_INF_F;
    e.hx = e.hy = e.hz = -CUDART_INF_F;
    e.flag = 0;
    e.pad = 0;
    return e;
}
__device__ __forceinline__ SweepItem shfl_item(const SweepItem &v, int src_lane) {
    SweepItem r;
    r.lx = __shfl_sync(0xffffffffu, v.lx, src_lane); r.ly = __shfl_sync(0xffffffffu, v.ly, src_lane); r.lz = __shfl_sync(0xffffffffu, v.lz, src_lane);
    r.hx = __shfl_sync(0xffffffffu, v.hx, src_lane); r.hy = __shfl_sync(0xffffffffu, v.hy, src_lane); r.hz = __shfl_sync(0xffffffffu, v.hz, src_lane);
    r.flag = __shfl_sync(0xffffffffu, v.flag, src_lane);
    r.pad = 0;
    return r;
}

// inclusive scan of one item per thread across the CTA, in thread order; returns the EXCLUSIVE prefix of the calling thread and
// leaves the CTA's total in s_warp[kBlock / 32]
__device__ __forceinline__ SweepItem cta_exclusive(const SweepItem &mine, SweepItem *s_warp) {
    const SweepUnion op;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SweepItem inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const SweepItem o = shfl_item(inc, (lane - d) & 31);
        if (lane >= d) inc = op(o, inc);
    }
    if (lane == 31) s_warp[warp] = inc;
    SweepItem excl = shfl_item(inc, (lane - 1) & 31);
    if (lane == 0) excl = scan_identity();
    __syncthreads();
    if (threadIdx.x == 0) {
        SweepItem acc = scan_identity();
        for (int w = 0; w < kBlock / 32; w++) { const SweepItem t = s_warp[w]; s_warp[w] = acc; acc = op(acc, t); }
        s_warp[kBlock / 32] = acc;
    }
    __syncthreads();
    return op(s_warp[warp], excl);
}

// grid = (tiles per row, 6 rows): a tile never straddles two rows, and nobody divides by n
__global__ void __launch_bounds__(kBlock) k_union_fold(Level V, SweepItem *__restrict__ tile_agg) {
    __shared__ SweepItem s_warp[kBlock / 32 + 1];
    const SweepUnion op;
    const int row = blockIdx.y, base = blockIdx.x * kTile + threadIdx.x * kPerThread;
    const int *list = list_of(V, row >> 1);
    const bool rev = row & 1;
    SweepItem acc = scan_identity();
#pragma unroll
    for (int k = 0; k < kPerThread; k++)
        if (base + k < V.n) acc = op(acc, sweep_item_at(V, list, rev, base + k));
    cta_exclusive(acc, s_warp);
    if (threadIdx.x == 0) tile_agg[row * gridDim.x + blockIdx.x] = s_warp[kBlock / 32];
}

// one CTA: tile_agg[t] becomes the exclusive prefix of the aggregates (the carry into tile t)
__global__ void __launch_bounds__(kBlock) k_union_carries(SweepItem *tile_agg, int tiles) {
    __shared__ SweepItem s_warp[kBlock / 32 + 1];
    const SweepUnion op;
    const int per = (tiles + kBlock - 1) / kBlock, first = threadIdx.x * per, last = min(first + per, tiles);
    SweepItem acc = scan_identity();
    for (int t = first; t < last; t++) acc = op(acc, tile_agg[t]);
    SweepItem carry = cta_exclusive(acc, s_warp);
    for (int t = first; t < last; t++) { const SweepItem mine = tile_agg[t]; tile_agg[t] = carry; carry = op(carry, mine); }
}

__global__ void __launch_bounds__(kBlock) k_union_scan(Level V, const SweepItem *__restrict__ tile_carry, float *__restrict__ areas) {
    __shared__ SweepItem s_warp[kBlock / 32 + 1];
    const SweepUnion op;
    const int row = blockIdx.y, base = blockIdx.x * kTile + threadIdx.x * kPerThread;
    const int *list = list_of(V, row >> 1);
    const bool rev = row & 1;
    SweepItem out[kPerThread];
    SweepItem acc = scan_identity();
#pragma unroll
    for (int k = 0; k < kPerThread; k++) {
        if (base + k < V.n) acc = op(acc, sweep_item_at(V, list, rev, base + k));
        out[k] = acc;
    }
    const SweepItem carry = op(tile_carry[row * gridDim.x + blockIdx.x], cta_exclusive(acc, s_warp));
    float *dst = areas + size_t(row) * V.n;
#pragma unroll
    for (int k = 0; k < kPerThread; k++)
        if (base + k < V.n) dst[base + k] = sweep_area(op(carry, out[k]));
}

__global__ void __launch_bounds__(kBlock) k_compare(const float *__restrict__ a, const float *__restrict__ b, int total, int *mismatches) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total && __float_as_uint(a[i]) != __float_as_uint(b[i])) atomicAdd(mismatches, 1);
}

// The positions of a node are contiguous, so most warps - near the top of the tree most CTAs - hold candidates of one node only:
// those are reduced in registers / shared memory first and cost one atomic.
__global__ void __launch_bounds__(kBlock) k_candidates(Level V, const float *__restrict__ areas, unsigned long long *best) {
    __shared__ unsigned long long s_key[kBlock / 32];
    __shared__ int s_nd[kBlock / 32];
    const long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
    unsigned long long key = kNoSplit;
    int nd = -1;
    if (c < 3ll * V.n) nd = sweep_candidate(V, areas, int(c), &key);
    const unsigned has = __ballot_sync(0xffffffffu, nd >= 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int w_nd = -1;
    unsigned long long w_key = kNoSplit;
    if (has) {
        const int leader = __ffs(has) - 1;
        const int nd0 = __shfl_sync(0xffffffffu, nd, leader);
        if (__all_sync(0xffffffffu, nd < 0 || nd == nd0)) {
            w_key = warp_min(nd >= 0 ? key : kNoSplit);
            w_nd = nd0;
        } else if (nd >= 0) atomicMin(best + nd, key);
    }
    if (lane == 0) { s_nd[warp] = w_nd; s_key[warp] = w_key; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int cur = -1;
        unsigned long long k = kNoSplit;
        for (int w = 0; w < kBlock / 32; w++) {
            if (s_nd[w] < 0) continue;
            if (s_nd[w] != cur) {
                if (cur >= 0) atomicMin(best + cur, k);
                cur = s_nd[w];
                k = s_key[w];
            } else k = s_key[w] < k ? s_key[w] : k;
        }
        if (cur >= 0) atomicMin(best + cur, k);
    }
}

__global__ void __launch_bounds__(kBlock) k_decide(Level V, const int *__restrict__ aB, const unsigned long long *__restrict__ best, int n_slots, int level, int depth_cap, Tree T,
                                                   Split S, NextLevel X) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_slots) sweep_decide(V, aB, s, best[s], level, depth_cap, T, S, X);
}

__global__ void __launch_bounds__(kBlock) k_mark(Level V, Split S, uint8_t *side) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V.n) sweep_mark(V, S, i, side);
}

struct OutLists { int *list[3]; };

__global__ void __launch_bounds__(kBlock) k_scatter(Level V, Split S, const uint8_t *__restrict__ side, const int *__restrict__ zeros, OutLists O, int *out_nodeid) {
    const long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (c < 3ll * V.n) sweep_scatter(V, S, side, zeros, int(c), O.list, out_nodeid);
}

__global__ void __launch_bounds__(kBlock) k_refit(Tree T, int first, int last) {
    const int b = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (b < last) sweep_refit(T, b);
}

} // namespace

int rm_gpu_build_wide_sah(RmContext *ctx, const float *d_pos, int n, int depth_cap, const float fallback_point[3], int *levels_out, int *nodes_out) {
    if (n <= 0) return rm_fail(RM_ERR_INVALID, "rm_gpu_build_wide_sah: no triangles");
    if (n > (1 << 28)) return rm_fail(RM_ERR_INVALID, "rm_gpu_build_wide_sah: more than 2^28 triangles");      // 6 n scan items are indexed by int
    cudaStream_t st = ctx->stream;
    int rc;
    DevBuf *B = ctx->b_build;
    const size_t N2 = size_t(2) * n;
    if ((rc = B[4].alloc(N2 * 16)) || (rc = B[5].alloc(N2 * 16)) || (rc = B[6].alloc(N2 * 4)) || (rc = B[7].alloc(N2 * 4)) || (rc = B[8].alloc(N2 * 4))) return rc;
    Tree T{B[4].as<float4>(), B[5].as<float4>(), B[6].as<int>(), B[7].as<int>(), B[8].as<int>()};

    // one arena for the builder's own arrays, in 4-byte words, every array 256-byte aligned
    size_t words = 0;
    auto take = [&](size_t w) { const size_t at = words; words += (w + 63) & ~size_t(63); return at; };
    const size_t nn = size_t(n);
    const size_t o_keys = take(3 * nn), o_keys_tmp = take(nn), o_ids = take(nn);
    const size_t o_list[2] = {take(3 * nn), take(3 * nn)}, o_nodeid[2] = {take(nn), take(nn)};
    size_t o_aL[2], o_aR[2], o_aB[2], o_best[2];
    for (int k = 0; k < 2; k++) { o_aL[k] = take(nn); o_aR[k] = take(nn); o_aB[k] = take(nn); o_best[k] = take(2 * nn); }
    const size_t o_axis = take(nn), o_M = take(nn), o_cl = take(nn), o_cr = take(nn);
    const int total = 6 * n, row_tiles = (n + kTile - 1) / kTile, tiles = 6 * row_tiles;
    // RM_SAH_SCAN=cub: the box-union scan through cub::DeviceScan (the first form of this builder); =check: both, compared bit for bit
    const char *senv = getenv("RM_SAH_SCAN");
    const bool scan_cub = senv && !strcmp(senv, "cub"), scan_check = senv && !strcmp(senv, "check");
    const size_t o_side = take((nn + 3) / 4), o_areas = take(6 * nn), o_zeros = take(3 * nn), o_counters = take(64);
    const size_t o_tbox = take(8 * nn);
    const size_t o_tiles = take(size_t(tiles) * 8), o_areas2 = take(scan_check ? 6 * nn : 0);
    if ((rc = B[0].alloc(words * 4))) return rc;
    int *W = B[0].as<int>();
    float *keys = reinterpret_cast<float *>(W + o_keys), *keys_tmp = reinterpret_cast<float *>(W + o_keys_tmp), *areas = reinterpret_cast<float *>(W + o_areas);
    int *ids = W + o_ids, *zeros = W + o_zeros, *counters = W + o_counters;
    uint8_t *side = reinterpret_cast<uint8_t *>(W + o_side);
    Split S{W + o_axis, W + o_M, W + o_cl, W + o_cr};

    Level V0{};
    V0.n = n; V0.tlo = T.lo; V0.thi = T.hi; V0.tbox = reinterpret_cast<const float4 *>(W + o_tbox);
    for (int a = 0; a < 3; a++) V0.list[a] = W + o_list[0] + size_t(a) * n;
    V0.nodeid = W + o_nodeid[0]; V0.aL = W + o_aL[0]; V0.aR = W + o_aR[0];
    auto items_in = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), ItemOp{V0});
    auto areas_out = thrust::make_transform_output_iterator(areas, AreaOp());
    auto left_in = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), LeftOp{V0, side});
    size_t temp_sort = 0, temp_scan = 0, temp_sum = 0;
    RM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_sort, keys, keys_tmp, ids, W + o_list[0], n, 0, 32, st));
    RM_CUDA(cub::DeviceScan::InclusiveScan(nullptr, temp_scan, items_in, areas_out, SweepUnion(), total, st));
    RM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp_sum, left_in, zeros, 3 * n, st));
    size_t temp_bytes = std::max(temp_sort, std::max(temp_scan, temp_sum));
    if ((rc = B[15].alloc(temp_bytes))) return rc;

    const int grid_n = (n + kBlock - 1) / kBlock, grid_3n = int((3ll * n + kBlock - 1) / kBlock);
    k_tri_boxes<<<grid_n, kBlock, 0, st>>>(d_pos, n, make_float3(fallback_point[0], fallback_point[1], fallback_point[2]), T, reinterpret_cast<float4 *>(W + o_tbox), keys, ids);
    for (int a = 0; a < 3; a++)
        RM_CUDA(cub::DeviceRadixSort::SortPairs(B[15].p, temp_bytes, keys + size_t(a) * n, keys_tmp, ids, W + o_list[0] + size_t(a) * n, n, 0, 32, st));
    ctx->launches += 1;

    int root = 0;
    if (n >= 2) {
        root = n;
        RM_CUDA(cudaMemsetAsync(W + o_nodeid[0], 0, nn * 4, st));          // every position belongs to slot 0, the root
        k_start<<<1, 1, 0, st>>>(n, W + o_aL[0], W + o_aR[0], W + o_aB[0], reinterpret_cast<unsigned long long *>(W + o_best[0]), counters);
        std::vector<std::pair<int, int>> ranges;          // inner nodes handed out per level
        ranges.push_back({n, n + 1});
        int n_slots = 1, next_node = n + 1;
        // RM_TIMING=2: device time per step, summed over the levels (events on the build's stream)
        const char *tenv = getenv("RM_TIMING");
        const bool timing = tenv && atoi(tenv) >= 2;
        std::vector<cudaEvent_t> ev;
        auto tick = [&]() { if (timing) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev.push_back(e); } };
        for (int level = 0, cur = 0; n_slots > 0; level++, cur ^= 1) {
            if (level > 4096) return rm_fail(RM_ERR_STATE, "rm_gpu_build_wide_sah: the build did not end");
            Level V = V0;
            for (int a = 0; a < 3; a++) V.list[a] = W + o_list[cur] + size_t(a) * n;
            V.nodeid = W + o_nodeid[cur]; V.aL = W + o_aL[cur]; V.aR = W + o_aR[cur];
            unsigned long long *best = reinterpret_cast<unsigned long long *>(W + o_best[cur]);
            NextLevel X{W + o_aL[cur ^ 1], W + o_aR[cur ^ 1], W + o_aB[cur ^ 1], reinterpret_cast<unsigned long long *>(W + o_best[cur ^ 1]), counters, counters + 1};
            OutLists O;
            for (int a = 0; a < 3; a++) O.list[a] = W + o_list[cur ^ 1] + size_t(a) * n;
            auto in = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), ItemOp{V});
            tick();
            if (scan_cub) RM_CUDA(cub::DeviceScan::InclusiveScan(B[15].p, temp_bytes, in, areas_out, SweepUnion(), total, st));
            else {
                SweepItem *tile_agg = reinterpret_cast<SweepItem *>(W + o_tiles);
                k_union_fold<<<dim3(row_tiles, 6), kBlock, 0, st>>>(V, tile_agg);
                k_union_carries<<<1, kBlock, 0, st>>>(tile_agg, tiles);
                k_union_scan<<<dim3(row_tiles, 6), kBlock, 0, st>>>(V, tile_agg, areas);
            }
            if (scan_check) {
                float *areas2 = reinterpret_cast<float *>(W + o_areas2);
                RM_CUDA(cub::DeviceScan::InclusiveScan(B[15].p, temp_bytes, in, thrust::make_transform_output_iterator(areas2, AreaOp()), SweepUnion(), total, st));
                k_compare<<<(total + kBlock - 1) / kBlock, kBlock, 0, st>>>(areas, areas2, total, counters + 2);
            }
            tick();
            k_candidates<<<grid_3n, kBlock, 0, st>>>(V, areas, best);
            tick();
            RM_CUDA(cudaMemsetAsync(counters + 1, 0, 4, st));
            k_decide<<<(n_slots + kBlock - 1) / kBlock, kBlock, 0, st>>>(V, W + o_aB[cur], best, n_slots, level, depth_cap, T, S, X);
            tick();
            k_mark<<<grid_n, kBlock, 0, st>>>(V, S, side);
            tick();
            auto lin = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), LeftOp{V, side});
            RM_CUDA(cub::DeviceScan::ExclusiveSum(B[15].p, temp_bytes, lin, zeros, 3 * n, st));
            tick();
            k_scatter<<<grid_3n, kBlock, 0, st>>>(V, S, side, zeros, O, W + o_nodeid[cur ^ 1]);
            tick();
            int h[3] = {0, 0, 0};
            RM_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, st));
            RM_CUDA(cudaStreamSynchronize(st));
            ctx->launches += 4;
            if (h[0] < next_node || h[0] > 2 * n - 1 || h[1] < 0 || h[1] > n) return rm_fail(RM_ERR_STATE, "rm_gpu_build_wide_sah: level %d handed out %d nodes, %d slots", level, h[0], h[1]);
            if (h[2]) return rm_fail(RM_ERR_STATE, "rm_gpu_build_wide_sah: level %d: %d areas differ between the tile scan and cub's", level, h[2]);
            ranges.push_back({next_node, h[0]});
            next_node = h[0];
            n_slots = h[1];
        }
        if (timing) {
            static const char *names[6] = {"box-union scan", "k_candidates", "k_decide", "k_mark", "flag scan", "k_scatter"};
            double sum[6] = {0, 0, 0, 0, 0, 0};
            for (size_t l = 0; l + 7 <= ev.size(); l += 7)
                for (int k = 0; k < 6; k++) { float ms = 0; cudaEventElapsedTime(&ms, ev[l + k], ev[l + k + 1]); sum[k] += ms; }
            for (int k = 0; k < 6; k++) fprintf(stderr, "rm_gpu_build_wide_sah: %-16s %7.2f ms over %zu levels\n", names[k], sum[k], ev.size() / 7);
            for (cudaEvent_t e : ev) cudaEventDestroy(e);
        }
        if (next_node != 2 * n - 1) return rm_fail(RM_ERR_STATE, "rm_gpu_build_wide_sah: %d of %d inner nodes", next_node - n, n - 1);
        for (int l = int(ranges.size()) - 1; l >= 0; l--) {
            const int cnt = ranges[l].second - ranges[l].first;
            if (cnt <= 0) continue;
            k_refit<<<(cnt + kBlock - 1) / kBlock, kBlock, 0, st>>>(T, ranges[l].first, ranges[l].second);
            ctx->launches++;
        }
        RM_CUDA(cudaGetLastError());
    }
    RmBinTree N{T.lo, T.hi, T.left, T.right, T.count};
    return rm_gpu_collapse_wide(ctx, N, root, n, levels_out, nodes_out);
}
