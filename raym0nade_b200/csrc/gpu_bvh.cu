// gpu_bvh.cu — the secondary-ray tree built on the device ("device-side BVH ... GPU build").
//
// rm_scene_upload used to build this tree on the host (fast_bvh.cpp: binned SAH, ~0.3 s per million triangles, per rank) and
// hide the cost behind a geometry-hash cache.  Here the same product - the 4-wide, 8-bit quantised tree of wide_bvh.h that
// bounce and shadow rays traverse - comes out of a handful of kernels in a few milliseconds, so every upload can afford it:
//
//   1. k_tri_keys      per triangle: box, centroid, 63-bit Morton code of the centroid inside the scene bounds
//   2. radix sort      triangles along the Morton curve (cub::DeviceRadixSort - library plumbing, not a hot path)
//   3. PLOC            parallel locally-ordered clustering (Meister & Bittner 2018): every cluster looks kPlocRadius
//                      neighbours left and right along the curve for the partner whose union has the smallest surface;
//                      mutual nearest neighbours merge into a new node; the survivors are compacted (cub::DeviceSelect,
//                      order preserved) and the round repeats until one cluster - the root - is left.  Bottom-up
//                      agglomeration by surface area: close to a SAH sweep build in quality, embarrassingly parallel.
//   4. k_collapse      top-down, one launch per level: a wide node takes its binary node's two children and keeps opening
//                      the child with the largest surface until it holds four (subtrees of <= 3 triangles are leaves),
//                      quantises their boxes on its own 8-bit grid (wide_quantise, the routine the host builder uses,
//                      checked conservative in double precision) and reserves consecutive records for its inner children.
//
// There is no reference counterpart (the reference's tree, src/bvh.cpp:18-54, is what primary rays traverse, built by the
// host as the north star keeps it); any valid hierarchy over the same triangles returns the same closest accepted hit under
// the reference's triangle test - tests/test_gpu_trace.py::test_secondary_ray_tree_finds_the_reference_hits runs on this tree.
#include <algorithm>
#include <cmath>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "rm_context.cuh"
#include "wide_bvh.h"
#include "gpu_bvh.h"

namespace {

constexpr int kPlocRadius = 16;
constexpr int kPlocBlock = 256;

using BNodes = RmBinTree;   // the binary tree under construction: nodes [0, n) are the triangles in Morton order, [n, 2n-1) merges

__device__ __forceinline__ unsigned long long spread21(unsigned v) {      // 21 bits -> every third bit of 63
    unsigned long long x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_tri_keys(const float *__restrict__ pos, int n, float3 lo, float3 scale, unsigned long long *keys, int *ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = pos + size_t(i) * 9;
    float c[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        c[a] = (p[a] + p[3 + a] + p[6 + a]) * (1.0f / 3.0f);
        if (!isfinite(c[a])) c[a] = 0.0f;            // a non-finite triangle can never be hit; it must not derail the order
    }
    const float fx = fminf(fmaxf((c[0] - lo.x) * scale.x, 0.0f), 2097151.0f), fy = fminf(fmaxf((c[1] - lo.y) * scale.y, 0.0f), 2097151.0f),
                fz = fminf(fmaxf((c[2] - lo.z) * scale.z, 0.0f), 2097151.0f);
    keys[i] = spread21(unsigned(fx)) | spread21(unsigned(fy)) << 1 | spread21(unsigned(fz)) << 2;
    ids[i] = i;
}

__global__ void k_tri_nodes(const float *__restrict__ pos, const int *__restrict__ ids, int n, float3 fallback, BNodes N, int *cluster) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int t = ids[k];
    const float *p = pos + size_t(t) * 9;
    float lo[3], hi[3];
    bool finite = true;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        lo[a] = fminf(fminf(p[a], p[3 + a]), p[6 + a]);
        hi[a] = fmaxf(fmaxf(p[a], p[3 + a]), p[6 + a]);
        finite = finite && isfinite(p[a]) && isfinite(p[3 + a]) && isfinite(p[6 + a]);
    }
    if (!finite) { lo[0] = hi[0] = fallback.x; lo[1] = hi[1] = fallback.y; lo[2] = hi[2] = fallback.z; }
    N.lo[k] = make_float4(lo[0], lo[1], lo[2], 0.0f);
    N.hi[k] = make_float4(hi[0], hi[1], hi[2], 0.0f);
    N.left[k] = ~t;
    N.right[k] = -1;
    N.count[k] = 1;
    cluster[k] = k;
}

__device__ __forceinline__ float union_area(float4 alo, float4 ahi, float4 blo, float4 bhi) {
    const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y), dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}

// nearest neighbour of every cluster within kPlocRadius positions along the curve; the boxes of a block's clusters and of
// the halo on either side are staged in shared memory.  Ties go to the smaller position, which makes "mutual nearest
// neighbours" well defined: the pair with the globally smallest union is always mutual, so every round merges something.
__global__ void __launch_bounds__(kPlocBlock) k_ploc_nn(const int *__restrict__ cluster, int m, BNodes N, int *__restrict__ nn) {
    __shared__ float4 s_lo[kPlocBlock + 2 * kPlocRadius], s_hi[kPlocBlock + 2 * kPlocRadius];
    const int b0 = blockIdx.x * kPlocBlock;
    for (int k = threadIdx.x; k < kPlocBlock + 2 * kPlocRadius; k += kPlocBlock) {
        const int j = b0 - kPlocRadius + k;
        if (j >= 0 && j < m) { const int c = cluster[j]; s_lo[k] = N.lo[c]; s_hi[k] = N.hi[c]; }
    }
    __syncthreads();
    const int i = b0 + threadIdx.x;
    if (i >= m) return;
    const int me = threadIdx.x + kPlocRadius;
    const float4 mlo = s_lo[me], mhi = s_hi[me];
    float best = CUDART_INF_F;
    int best_j = -1;
    const int j0 = max(i - kPlocRadius, 0), j1 = min(i + kPlocRadius, m - 1);
    for (int j = j0; j <= j1; j++) {
        if (j == i) continue;
        const int k = j - b0 + kPlocRadius;
        const float a = union_area(mlo, mhi, s_lo[k], s_hi[k]);
        if (a < best || best_j < 0) { best = a; best_j = j; }
    }
    nn[i] = best_j;
}

__global__ void __launch_bounds__(kPlocBlock) k_ploc_merge(const int *__restrict__ cluster, const int *__restrict__ nn, int m, BNodes N, int *next_node,
                                                          int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int j = nn[i], c = cluster[i];
    if (j < 0 || nn[j] != i) { out[i] = c; return; }
    if (i > j) { out[i] = -1; return; }               // the pair lives on at the smaller position
    const int d = cluster[j];
    const int id = atomicAdd(next_node, 1);
    const float4 alo = N.lo[c], ahi = N.hi[c], blo = N.lo[d], bhi = N.hi[d];
    N.lo[id] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
    N.hi[id] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
    N.left[id] = c;
    N.right[id] = d;
    N.count[id] = N.count[c] + N.count[d];
    out[i] = id;
}

struct IsCluster { __device__ bool operator()(int v) const { return v >= 0; } };

// ---- collapse
struct LevelItem { int bin, slot; };

__device__ __forceinline__ float node_area(const BNodes &N, int b) {
    const float4 lo = N.lo[b], hi = N.hi[b];
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// the <= 3 triangles of a small subtree, left to right
__device__ int small_subtree(const BNodes &N, int b, int tri[3]) {
    if (N.right[b] < 0) { tri[0] = ~N.left[b]; return 1; }
    int n = 0;
    const int kids[2] = {N.left[b], N.right[b]};
    for (int s = 0; s < 2; s++) {
        const int k = kids[s];
        if (N.right[k] < 0) { if (n < 3) tri[n++] = ~N.left[k]; }
        else {                                      // two triangles (count 2): both children are triangles
            if (n < 3) tri[n++] = ~N.left[N.left[k]];
            if (n < 3) tri[n++] = ~N.left[N.right[k]];
        }
    }
    return n;
}

__global__ void __launch_bounds__(128) k_collapse(BNodes N, const LevelItem *__restrict__ in, const int *__restrict__ in_count, LevelItem *out, int *out_count,
                                                  RmWideNode *wide, int *next_wide, int *order, int *next_tri, int level, int *levels) {
    const int n_in = *in_count;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_in; e += gridDim.x * blockDim.x) {
        if (e == 0) atomicMax(levels, level + 1);
        const LevelItem it = in[e];
        int ch[4], n = 0;
        if (N.count[it.bin] <= 3) ch[n++] = it.bin;           // only the root of a scene of <= 3 triangles
        else {
            ch[n++] = N.left[it.bin];
            ch[n++] = N.right[it.bin];
            while (n < 4) {
                int open = -1;
                float open_area = -1.0f;
                for (int k = 0; k < n; k++)
                    if (N.count[ch[k]] > 3) {
                        const float a = node_area(N, ch[k]);
                        if (open < 0 || a > open_area) { open = k; open_area = a; }
                    }
                if (open < 0) break;
                const int b = ch[open];
                ch[open] = N.left[b];
                ch[n++] = N.right[b];
            }
        }
        float lo[4][3], hi[4][3];
        int n_inner = 0, n_tris = 0;
        for (int k = 0; k < n; k++) {
            const float4 l = N.lo[ch[k]], h = N.hi[ch[k]];
            lo[k][0] = l.x; lo[k][1] = l.y; lo[k][2] = l.z; hi[k][0] = h.x; hi[k][1] = h.y; hi[k][2] = h.z;
            if (N.count[ch[k]] > 3) n_inner++; else n_tris += N.count[ch[k]];
        }
        RmWideNode w;
        wide_quantise(lo, hi, n, w);
        const int child_base = n_inner ? atomicAdd(next_wide, n_inner) : 0;
        const int tri_base = n_tris ? atomicAdd(next_tri, n_tris) : 0;
        const int q = n_inner ? atomicAdd(out_count, n_inner) : 0;
        w.child_base = child_base;
        w.tri_base = tri_base;
        w._pad = 0;
        int k_inner = 0, tri_off = 0;
        for (int k = 0; k < 4; k++) w.meta[k] = 0;
        for (int k = 0; k < n; k++) {
            if (N.count[ch[k]] > 3) {
                w.meta[k] = uint8_t(0x80 | k_inner);
                out[q + k_inner] = LevelItem{ch[k], child_base + k_inner};
                k_inner++;
            } else {
                int tri[3];
                const int cnt = small_subtree(N, ch[k], tri);
                w.meta[k] = uint8_t((tri_off << 2) | cnt);
                for (int t = 0; t < cnt; t++) order[tri_base + tri_off + t] = tri[t];
                tri_off += cnt;
            }
        }
        // one 64-byte record: four 16-byte stores
        const uint4 *src = reinterpret_cast<const uint4 *>(&w);
        uint4 *dst = reinterpret_cast<uint4 *>(wide + it.slot);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
    }
}

__global__ void k_collapse_start(int root, LevelItem *q, int *counts) {
    q[0] = LevelItem{root, 0};
    counts[0] = 1;          // level-0 queue length
    counts[1] = 0;
    counts[2] = 1;          // next_wide: record 0 is the root
    counts[3] = 0;          // next_tri
    counts[4] = 0;          // levels
}

} // namespace

// positions: device, [n][9].  Fills ctx->b_nodes_wide / b_facemap_wide (device) and reports levels.  Asynchronous on the
// context's stream except for one 4-byte read-back per clustering round.
int rm_gpu_build_wide(RmContext *ctx, const float *d_pos, int n, const float scene_lo[3], const float scene_hi[3], int *levels_out, int *nodes_out) {
    if (n <= 0) return rm_fail(RM_ERR_INVALID, "rm_gpu_build_wide: no triangles");
    cudaStream_t st = ctx->stream;
    int rc;
    const size_t N2 = size_t(2) * n;
    DevBuf *B = ctx->b_build;
    // [0] keys a, [1] keys b, [2] ids a, [3] ids b, [4] lo, [5] hi, [6] left, [7] right, [8] count, [9..11] clusters, [12] nn,
    // [13] level queues (2), [14] counters, [15] cub temp
    if ((rc = B[0].alloc(size_t(n) * 8)) || (rc = B[1].alloc(size_t(n) * 8)) || (rc = B[2].alloc(size_t(n) * 4)) || (rc = B[3].alloc(size_t(n) * 4)) ||
        (rc = B[4].alloc(N2 * 16)) || (rc = B[5].alloc(N2 * 16)) || (rc = B[6].alloc(N2 * 4)) || (rc = B[7].alloc(N2 * 4)) || (rc = B[8].alloc(N2 * 4)) ||
        (rc = B[9].alloc(size_t(n) * 4)) || (rc = B[10].alloc(size_t(n) * 4)) || (rc = B[11].alloc(size_t(n) * 4)) || (rc = B[12].alloc(size_t(n) * 4)) ||
        (rc = B[13].alloc(size_t(n + 1) * 2 * sizeof(LevelItem))) || (rc = B[14].alloc(64)))
        return rc;
    if ((rc = ctx->b_nodes_wide.alloc(size_t(n + 1) * sizeof(RmWideNode))) || (rc = ctx->b_facemap_wide.alloc(size_t(n) * 4))) return rc;
    size_t temp_sort = 0, temp_sel = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_sort, B[0].as<unsigned long long>(), B[1].as<unsigned long long>(), B[2].as<int>(), B[3].as<int>(), n, 0, 63, st);
    cub::DeviceSelect::If(nullptr, temp_sel, B[10].as<int>(), B[11].as<int>(), B[14].as<int>() + 8, n, IsCluster(), st);
    size_t temp_bytes = std::max(temp_sort, temp_sel);
    if ((rc = B[15].alloc(temp_bytes))) return rc;

    float3 lo = make_float3(scene_lo[0], scene_lo[1], scene_lo[2]);
    float3 scale;
    {
        const float ex = scene_hi[0] - scene_lo[0], ey = scene_hi[1] - scene_lo[1], ez = scene_hi[2] - scene_lo[2];
        scale = make_float3(ex > 0.0f ? 2097151.0f / ex : 0.0f, ey > 0.0f ? 2097151.0f / ey : 0.0f, ez > 0.0f ? 2097151.0f / ez : 0.0f);
        if (!std::isfinite(scale.x)) scale.x = 0.0f;
        if (!std::isfinite(scale.y)) scale.y = 0.0f;
        if (!std::isfinite(scale.z)) scale.z = 0.0f;
    }
    const int grid_n = (n + 255) / 256;
    k_tri_keys<<<grid_n, 256, 0, st>>>(d_pos, n, lo, scale, B[0].as<unsigned long long>(), B[2].as<int>());
    RM_CUDA(cub::DeviceRadixSort::SortPairs(B[15].p, temp_bytes, B[0].as<unsigned long long>(), B[1].as<unsigned long long>(), B[2].as<int>(), B[3].as<int>(), n, 0, 63, st));
    BNodes N{B[4].as<float4>(), B[5].as<float4>(), B[6].as<int>(), B[7].as<int>(), B[8].as<int>()};
    int *cl_a = B[9].as<int>(), *cl_tmp = B[10].as<int>(), *cl_b = B[11].as<int>(), *nn = B[12].as<int>();
    int *counters = B[14].as<int>();          // [0..4] collapse, [8] selected count, [9] next_node
    k_tri_nodes<<<grid_n, 256, 0, st>>>(d_pos, B[3].as<int>(), n, lo, N, cl_a);
    RM_CUDA(cudaMemcpyAsync(counters + 9, &n, 4, cudaMemcpyHostToDevice, st));          // (pageable source: copied before the call returns)
    ctx->launches += 2;
    int m = n, rounds = 0;
    while (m > 1) {
        const int grid = (m + kPlocBlock - 1) / kPlocBlock;
        k_ploc_nn<<<grid, kPlocBlock, 0, st>>>(cl_a, m, N, nn);
        k_ploc_merge<<<grid, kPlocBlock, 0, st>>>(cl_a, nn, m, N, counters + 9, cl_tmp);
        RM_CUDA(cub::DeviceSelect::If(B[15].p, temp_bytes, cl_tmp, cl_b, counters + 8, m, IsCluster(), st));
        int m_new = 0;
        RM_CUDA(cudaMemcpyAsync(&m_new, counters + 8, 4, cudaMemcpyDeviceToHost, st));
        RM_CUDA(cudaStreamSynchronize(st));
        ctx->launches += 2;
        if (m_new <= 0 || m_new >= m) return rm_fail(RM_ERR_STATE, "rm_gpu_build_wide: clustering round %d went from %d to %d clusters", rounds, m, m_new);
        m = m_new;
        std::swap(cl_a, cl_b);
        if (++rounds > 4096) return rm_fail(RM_ERR_STATE, "rm_gpu_build_wide: clustering did not converge");
    }
    int root = 0;
    RM_CUDA(cudaMemcpyAsync(&root, cl_a, 4, cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaStreamSynchronize(st));

    return rm_gpu_collapse_wide(ctx, N, root, n, levels_out, nodes_out);
}

// The binary tree N (root `root`, n triangles) collapsed into ctx->b_nodes_wide / b_facemap_wide, one launch per level; the
// queues hold at most n items each.  Shared by the PLOC builder above and the sweep-SAH builder (gpu_sah_bvh.cu).
int rm_gpu_collapse_wide(RmContext *ctx, const RmBinTree &N, int root, int n, int *levels_out, int *nodes_out) {
    cudaStream_t st = ctx->stream;
    int rc;
    DevBuf *B = ctx->b_build;
    if ((rc = B[13].alloc(size_t(n + 1) * 2 * sizeof(LevelItem))) || (rc = B[14].alloc(64))) return rc;
    if ((rc = ctx->b_nodes_wide.alloc(size_t(n + 1) * sizeof(RmWideNode))) || (rc = ctx->b_facemap_wide.alloc(size_t(n) * 4))) return rc;
    int *counters = B[14].as<int>();
    LevelItem *q[2] = {B[13].as<LevelItem>(), B[13].as<LevelItem>() + (n + 1)};
    k_collapse_start<<<1, 1, 0, st>>>(root, q[0], counters);
    const int cgrid = std::max(1, std::min(ctx->sm_count * 8, (n + 127) / 128));
    int h[5] = {0, 0, 0, 0, 0};
    for (int level = 0, cur = 0; level < 96; level++, cur ^= 1) {
        k_collapse<<<cgrid, 128, 0, st>>>(N, q[cur], counters + cur, q[cur ^ 1], counters + (cur ^ 1), ctx->b_nodes_wide.as<RmWideNode>(), counters + 2,
                                          ctx->b_facemap_wide.as<int>(), counters + 3, level, counters + 4);
        RM_CUDA(cudaMemsetAsync(counters + cur, 0, 4, st));          // this level's queue is consumed: it becomes the next output queue
        ctx->launches++;
        if ((level & 7) == 7) {
            RM_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, st));
            RM_CUDA(cudaStreamSynchronize(st));
            if (h[cur ^ 1] == 0) break;
        }
    }
    RM_CUDA(cudaGetLastError());
    RM_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaStreamSynchronize(st));
    if (h[0] != 0 || h[1] != 0) return rm_fail(RM_ERR_STATE, "rm_gpu_collapse_wide: the tree is deeper than 96 wide levels");
    if (h[3] != n) return rm_fail(RM_ERR_STATE, "rm_gpu_collapse_wide: the collapse placed %d of %d triangles", h[3], n);
    if (levels_out) *levels_out = h[4];
    if (nodes_out) *nodes_out = h[2];
    return RM_OK;
}
